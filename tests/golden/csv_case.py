"""The fixed arrays a fake ``ActorCriticSolver.train()`` returns when the CSV writers of main.py are compared byte for
byte (tests/golden/make_golden_csv.py runs the REFERENCE's main.py on them, tests/test_main_csv.py runs ours)."""
import numpy as np

CONFIG = {
    "eqn_config": {"eqn_name": "LQR", "discount": 1.0, "p": 1.0, "q": 1.0, "beta": 1.0, "R": 1.0, "dim": 3, "control_dim": 3,
                   "total_time_critic": 0.1, "total_time_actor": 0.1, "num_time_interval_critic": 5, "num_time_interval_actor": 5},
    "net_config": {"num_hiddens_critic": [8], "num_hiddens_actor": [8], "lr_values_critic": [1e-3], "lr_boundaries_critic": [],
                   "lr_values_actor": [1e-3], "lr_boundaries_actor": [], "num_iterations": 3, "batch_size": 8, "valid_size": 6,
                   "logging_frequency": 1, "dtype": "float64", "verbose": False},
    "train_config": {"sample_type": "normal", "scheme": "adaptive", "TD_type": "TD1", "train": "actor-critic"},
}


def fake_train_result():
    """the 7-tuple of ActorCriticSolver.train() (solver.py:71): history rows [step, 7 floats, elapsed] + a sentinel row"""
    rng = np.random.RandomState(20260218)
    n, d, m = 6, 3, 3
    hist = []
    for step in range(4):
        hist.append([step] + list(np.exp(rng.normal(0, 3, 7)) * rng.choice([-1, 1], 7)) + [step * 37 + 0.6])
    hist.append([0, 0.0, 0.123456789, 0.0, 0.0, 0.0, 0.0, 0.0, 111.9])       # solver.py:64 sentinel row
    x = rng.normal(0, 0.5, (n, d))
    y, true_y = rng.normal(0, 1, (n, 1)), rng.normal(0, 1, (n, 1))
    z, true_z = rng.normal(0, 1, (n, m)), rng.normal(0, 1, (n, m))
    grad_y = rng.normal(0, 1, (n, d))
    return np.array(hist), x, y, true_y, z, true_z, grad_y
