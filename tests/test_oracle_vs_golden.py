"""Pin the oracle (oracle/ref_*.py) against golden vectors produced by the reference's own
source files under the TF shim (tests/golden/make_golden.py).  CPU only."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_equation as RE
from oracle import ref_solver as RS

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if "samplers" not in p)
TOL = dict(rtol=1e-10, atol=1e-12)


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = json.loads(str(z["config_json"]))
    return z, cfg


def setup(z, cfg):
    eqn = RE.make_ref_equation(cfg["eqn_config"])
    thetas = {k: torch.tensor(z["theta_" + k]) for k in ("actor", "critic", "critic_grad")}
    inputs = tuple(torch.tensor(z[k]) for k in ("x0", "dw", "xb"))
    return eqn, thetas, inputs


def test_cases_exist():
    assert len(CASES) >= 8


@pytest.mark.parametrize("name", CASES)
def test_networks_and_closed_forms(name):
    z, cfg = load(name)
    eqn, th, (x0, dw, xb) = setup(z, cfg)
    nets = RS.RefNets(cfg, th["actor"], th["critic"], th["critic_grad"])
    for k in ("actor", "critic", "critic_grad"):
        np.testing.assert_allclose(nets(k, x0).numpy(), z["net_" + k], **TOL)
    np.testing.assert_allclose(eqn.u_true(x0).numpy(), z["u_true"], **TOL)
    np.testing.assert_allclose(eqn.V_true(x0).numpy(), z["V_true"], **TOL)
    np.testing.assert_allclose(eqn.V_grad_true(x0).numpy(), z["V_grad_true"], **TOL)
    np.testing.assert_allclose(eqn.Z(xb).numpy(), z["Z_tf"], **TOL)
    np.testing.assert_allclose(eqn.w(x0, nets("actor", x0)).numpy(), z["w_tf"], **TOL)
    # SDE coefficients (equation.py:169-176,229-238,267-276,304-311) and CriticModel.control (solver.py:153-157)
    u = nets("actor", x0)
    np.testing.assert_allclose(torch.diag_embed(eqn.sigma_diag(x0, u) * torch.ones_like(x0)).numpy(), z["sigma"], **TOL)
    np.testing.assert_allclose(eqn.drift(x0, u).numpy(), z["drift"], **TOL)
    np.testing.assert_allclose(eqn.diffusion(x0, u, dw[:, :, 0]).numpy(), z["diffusion"], **TOL)
    np.testing.assert_allclose(u.numpy(), z["control_nn"], **TOL)
    np.testing.assert_allclose(eqn.u_true(x0).numpy(), z["control_cheat"], **TOL)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("cheat", [False, True])
def test_propagate(name, cheat):
    z, cfg = load(name)
    eqn, th, (x0, dw, xb) = setup(z, cfg)
    nets = RS.RefNets(cfg, th["actor"], th["critic"], th["critic_grad"])
    ec = cfg["eqn_config"]
    control = (lambda x: eqn.u_true(x)) if cheat else (lambda x: nets("actor", x))
    x, dt, coef = RS._propagate(eqn, cfg, x0, dw, control, ec["total_time_critic"], ec["num_time_interval_critic"])
    tag = "cheat" if cheat else "nn"
    # schedule: exit pattern and step sizes bit-exact
    assert np.array_equal(coef.numpy(), z[f"prop_{tag}_coef"])
    np.testing.assert_allclose(dt.numpy(), z[f"prop_{tag}_dt"], rtol=1e-13, atol=0)
    np.testing.assert_allclose(x.numpy(), z[f"prop_{tag}_x"], **TOL)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("cheat", [False, True])
def test_critic(name, cheat):
    z, cfg = load(name)
    eqn, th, inputs = setup(z, cfg)
    tag = "cheat" if cheat else "nn"
    loss, gV, gG, delta, delta_b, _ = RS.grad_critic(eqn, cfg, th, inputs, cheat)
    np.testing.assert_allclose(delta.numpy(), z[f"critic_{tag}_delta"], **TOL)
    np.testing.assert_allclose(delta_b.numpy(), z[f"critic_{tag}_delta_bdry"], **TOL)
    np.testing.assert_allclose(float(loss), float(z[f"critic_{tag}_loss"]), rtol=1e-11)
    np.testing.assert_allclose(gV.numpy(), z[f"critic_{tag}_grad_V"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(gG.numpy(), z[f"critic_{tag}_grad_G"], rtol=1e-8, atol=1e-11)
    if cfg["train_config"]["TD_type"] == "TD2":
        assert not gG.any()       # NN_value_grad unused under LSTD (solver.py:177)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("cheat_v", [False, True])
def test_actor(name, cheat_v):
    z, cfg = load(name)
    eqn, th, inputs = setup(z, cfg)
    tag = "cheatV" if cheat_v else "nn"
    loss, gA, y, _ = RS.grad_actor(eqn, cfg, th, inputs, cheat_v, False)
    np.testing.assert_allclose(y.numpy(), z[f"actor_{tag}_y"], **TOL)
    np.testing.assert_allclose(float(loss), float(z[f"actor_{tag}_loss"]), rtol=1e-11)
    np.testing.assert_allclose(gA.numpy(), z[f"actor_{tag}_grad"], rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("name", CASES)
def test_true_loss_and_errors(name):
    z, cfg = load(name)
    eqn, th, inputs = setup(z, cfg)
    nets = RS.RefNets(cfg, th["actor"], th["critic"], th["critic_grad"])
    with torch.no_grad():
        loss, _, _ = RS.loss_actor(eqn, cfg, nets, inputs, True, True)
        np.testing.assert_allclose(float(loss), float(z["actor_true_loss"]), rtol=1e-11)
        y, _ = RS.actor_cost(eqn, cfg, nets, inputs, False, False)
        err_cost = torch.mean(y - nets("critic", inputs[0]))
        np.testing.assert_allclose(float(err_cost), float(z["err_cost"]), rtol=1e-9, atol=1e-12)
    s = RS.RefSolver(cfg, eqn)
    s.thetas = th
    e = s.errors(inputs[0])
    for k in ("err_value", "err_control", "err_value_grad", "err_value_infty"):
        np.testing.assert_allclose(e[k], float(z[k]), rtol=1e-10)


@pytest.mark.parametrize("name", CASES)
def test_three_adam_iterations(name):
    """Keras Adam + PiecewiseConstantDecay semantics (solver.py:16-21,99-107)."""
    z, cfg = load(name)
    eqn, th, inputs = setup(z, cfg)
    s = RS.RefSolver(cfg, eqn)
    s.thetas = {k: v.clone() for k, v in th.items()}
    s.cheat_control_in_critic = False
    s.cheat_value_in_actor = False
    for _ in range(3):
        s.train_step_critic(inputs)
        s.train_step_actor(inputs)
    for k in ("actor", "critic", "critic_grad"):
        np.testing.assert_allclose(s.thetas[k].numpy(), z["theta_after3_" + k], rtol=1e-7, atol=1e-10)


def test_sampler_streams():
    """equation.py:13-44 under np.random.seed: scipy multivariate_normal.rvs consumes the same
    Mersenne-Twister stream as standard_normal."""
    z = np.load(os.path.join(GOLDEN, "samplers_seed7.npz"))
    eqn = RE.make_ref_equation({"eqn_name": "LQR", "dim": 5, "control_dim": 5, "discount": 1.0, "R": 1.0,
                                "p": 1.0, "q": 1.0, "beta": 1.0})
    for fn in ("sample_normal", "sample_bounded", "sample0"):
        np.random.seed(7)
        x0, dw, xb = getattr(eqn, fn)(6, 4)
        np.testing.assert_allclose(x0, z[fn + "_x0"], rtol=1e-14)
        np.testing.assert_allclose(dw, z[fn + "_dw"], rtol=1e-14)
        np.testing.assert_allclose(xb, z[fn + "_xb"], rtol=1e-14)
