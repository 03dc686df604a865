#!/usr/bin/env python
"""Writes the 12 experiment configs (same file names, keys and hyper-parameters as the reference's
configs/*.json, so `main.py --config_path=configs/lqr_d5.json` is a drop-in) plus the four
BASELINE.json variants.  Generated rather than hand-kept so that the table below is the single
source.  Extra optional keys understood by this implementation only: train_config.sampler
("device" | "host")."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))

EQN = {
    "lqr": dict(eqn_name="LQR", discount=1.0, p=1.0, q=1.0, beta=1.0),
    "vdp": dict(eqn_name="VDP", discount=1.0, a=1.0, epsilon=0.1, q=1.0),
    "ekn": dict(eqn_name="EKN", discount=0, a2=1.2, a3=0.2),
    "lqr_var": dict(eqn_name="LQR_var", discount=1.0, q=1.0, beta=1.0),
}
# name: (family, dim, control_dim, N, T, hidden, batch, iterations, lr boundaries, lr values, extra eqn keys)
TABLE = {
    "lqr_d5": ("lqr", 5, 5, 50, 0.2, [200, 200], 1024, 40000, [20000, 30000], [1e-3, 1e-4, 1e-5], {}),
    "lqr_d10": ("lqr", 10, 10, 100, 0.2, [200] * 3, 2048, 40000, [20000, 30000], [1e-3, 1e-4, 1e-5], {}),
    "lqr_d20": ("lqr", 20, 20, 100, 0.2, [200] * 3, 2048, 50000, [30000, 40000], [1e-3, 1e-4, 1e-5], {}),
    "vdp_d4": ("vdp", 4, 2, 50, 0.1, [50, 50], 512, 15000, [10000], [1e-3, 1e-4], {}),
    "vdp_d10": ("vdp", 10, 5, 100, 0.2, [200] * 3, 2048, 40000, [20000, 30000], [1e-3, 1e-4, 1e-5], {}),
    "vdp_d20": ("vdp", 20, 10, 100, 0.2, [200] * 3, 2048, 50000, [30000, 40000], [1e-3, 1e-4, 1e-5], {}),
    "ekn_d5": ("ekn", 5, 5, 50, 0.2, [200, 200], 1024, 40000, [20000, 30000], [1e-3, 1e-4, 1e-5], {}),
    "ekn_d10": ("ekn", 10, 10, 100, 0.2, [200] * 3, 2048, 40000, [20000, 30000], [1e-3, 1e-4, 1e-5], {}),
    "ekn_d20": ("ekn", 20, 20, 100, 0.2, [200] * 3, 2048, 50000, [30000, 40000], [1e-3, 1e-4, 1e-5], {}),
    "lqr_var_d5": ("lqr_var", 5, 5, 50, 0.2, [200, 200], 1024, 40000, [20000, 30000], [1e-3, 1e-4, 1e-5], {"epsilon": 0.1}),
    "lqr_var_d10": ("lqr_var", 10, 10, 100, 0.2, [200] * 3, 2048, 40000, [20000, 30000], [1e-3, 1e-4, 1e-5], {"epsilon": 0.1}),
    "lqr_var_d20": ("lqr_var", 20, 20, 100, 0.2, [200] * 3, 2048, 50000, [30000, 40000], [1e-3, 1e-4, 1e-5], {"epsilon": 0.01}),
}
# BASELINE.json:configs variants (field overrides of the shipped files)
VARIANTS = {
    "bench_lqr_d5_naive_normal_td1": ("lqr_d5", dict(scheme="naive", sample_type="normal", TD_type="TD1")),
    "bench_vdp_d10_adaptive_bounded_td2": ("vdp_d10", dict(scheme="adaptive", sample_type="bounded", TD_type="TD2")),
    "bench_ekn_d20_adaptive_normal_td1": ("ekn_d20", dict()),
    "bench_lqr_var_d20_adaptive_normal_td1": ("lqr_var_d20", dict()),
}


def make(name, train_over=None):
    fam, d, m, N, T, hid, B, iters, bnd, vals, extra = TABLE[name]
    e = dict(EQN[fam])
    e.update(extra)
    e.update(dim=d, control_dim=m, total_time_critic=T, total_time_actor=T, num_time_interval_critic=N,
             num_time_interval_actor=N, R=1.0)
    net = dict(num_hiddens_critic=hid, num_hiddens_actor=hid, lr_values_critic=vals, lr_boundaries_critic=bnd,
               lr_values_actor=vals, lr_boundaries_actor=bnd, num_iterations=iters, batch_size=B, valid_size=B,
               logging_frequency=100, dtype="float64", verbose=True)
    train = dict(sample_type="normal", scheme="adaptive", TD_type="TD1", train="actor-critic")
    train.update(train_over or {})
    return {"eqn_config": e, "net_config": net, "train_config": train}


if __name__ == "__main__":
    for name in TABLE:
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(make(name), f, indent=1, sort_keys=True)
    for name, (base, over) in VARIANTS.items():
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(make(base, over), f, indent=1, sort_keys=True)
