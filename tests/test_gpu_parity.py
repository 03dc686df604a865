"""CUDA path vs the oracle, through the C ABI (libdeeppde_b200.so), on the golden fixtures that the
reference's own source produced (tests/golden/make_golden.py) and on seeded oracle runs.

Tolerances (stated, BASELINE.json north_star):
  float64 exact path : values 1e-9 relative, gradients 1e-7 relative to the gradient's max-norm;
                       coef / exit index bit-exact, dt 1e-12 relative.
  float32 exact path : values 2e-4 (relative to max(1,|ref|)), gradients 2e-3 of the gradient's
                       max-norm; schedule compared bit-exactly against the float32 CPU build of the
                       very same per-path code (tests/harness) under the true control.
"""
import ctypes as C
import glob
import json
import os
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from deeppde_actorcritic_b200 import _cabi
from deeppde_actorcritic_b200.engine import Engine
from oracle import ref_equation as RE
from oracle import ref_solver as RS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if "samplers" not in p)
DTYPES = ["float64", "float32"]
VTOL = {"float64": dict(rtol=1e-9, atol=1e-11), "float32": dict(rtol=2e-4, atol=2e-4)}
GTOL = {"float64": 1e-7, "float32": 2e-3}


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z, json.loads(str(z["config_json"]))


def engine_for(cfg, dtype):
    return Engine(cfg["eqn_config"], cfg["net_config"], cfg["train_config"], dtype=dtype)


def dev(eng, z, keys):
    return [eng.tensor(z[k]) for k in keys]


def npy(t):
    return t.detach().cpu().double().numpy()


def assert_grad_close(got, ref, tol, what):
    scale = np.abs(ref).max()
    if scale == 0:
        assert not np.any(got), what
        return
    err = np.abs(got - ref).max() / scale
    assert err < tol, f"{what}: max-norm relative error {err:.3e} >= {tol}"


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", CASES)
def test_networks_and_closed_forms(name, dtype):
    z, cfg = load(name)
    eng = engine_for(cfg, dtype)
    x0, xb = dev(eng, z, ["x0", "xb"])
    for k in ("actor", "critic", "critic_grad"):
        out = eng.mlp_forward(k, eng.tensor(z["theta_" + k]), x0)
        np.testing.assert_allclose(npy(out), z["net_" + k], **VTOL[dtype])
    np.testing.assert_allclose(npy(eng.closed_form(_cabi.CF_U_TRUE, x0)), z["u_true"], **VTOL[dtype])
    np.testing.assert_allclose(npy(eng.closed_form(_cabi.CF_V_TRUE, x0)), z["V_true"], **VTOL[dtype])
    np.testing.assert_allclose(npy(eng.closed_form(_cabi.CF_V_GRAD_TRUE, x0)), z["V_grad_true"], **VTOL[dtype])
    np.testing.assert_allclose(npy(eng.closed_form(_cabi.CF_Z, xb)), z["Z_tf"], **VTOL[dtype])
    u = eng.mlp_forward("actor", eng.tensor(z["theta_actor"]), x0)
    np.testing.assert_allclose(npy(eng.closed_form(_cabi.CF_W, x0, u)), z["w_tf"], **VTOL[dtype])
    # SDE coefficients through the C ABI (equation.py:169-176,229-238,267-276,304-311)
    np.testing.assert_allclose(npy(eng.closed_form(_cabi.CF_SIGMA, x0, u)), z["sigma"], **VTOL[dtype])
    np.testing.assert_allclose(npy(eng.closed_form(_cabi.CF_DRIFT, x0, u)), z["drift"], **VTOL[dtype])
    np.testing.assert_allclose(npy(eng.diffusion(x0, u, eng.tensor(z["dw"][:, :, 0]))), z["diffusion"], **VTOL[dtype])


@pytest.mark.parametrize("name", CASES)
def test_equation_and_control_api(name):
    """the reference's callable surface: Equation.sigma / drift / diffusion (equation.py:132-142 and subclasses) and
    CriticModel.control (solver.py:153-157), same names and argument order, NumPy in / device tensors out"""
    from deeppde_actorcritic_b200 import equation, munchify
    from deeppde_actorcritic_b200.solver import ActorCriticSolver
    z, cfg = load(name)
    config = munchify(cfg)
    bsde = getattr(equation, config.eqn_config.eqn_name)(config.eqn_config)
    s = ActorCriticSolver(config, bsde, compute_dtype="float64", seed=0)
    s.model_actor.NN_control.theta.copy_(s.engine.tensor(z["theta_actor"]))
    x0, dw0 = z["x0"], z["dw"][:, :, 0]
    n = x0.shape[0]
    u = s.model_critic.control(x0, False, s.model_actor)
    np.testing.assert_allclose(npy(u), z["control_nn"], **VTOL["float64"])
    np.testing.assert_allclose(npy(s.model_critic.control(x0, True, s.model_actor)), z["control_cheat"], **VTOL["float64"])
    np.testing.assert_allclose(npy(bsde.sigma(x0, u, n)), z["sigma"], **VTOL["float64"])
    np.testing.assert_allclose(npy(bsde.drift(x0, u)), z["drift"], **VTOL["float64"])
    np.testing.assert_allclose(npy(bsde.diffusion(x0, u, dw0, n)), z["diffusion"], **VTOL["float64"])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cheat", [False, True])
@pytest.mark.parametrize("name", CASES)
def test_propagate(name, cheat, dtype):
    z, cfg = load(name)
    eng = engine_for(cfg, dtype)
    x0, dw = dev(eng, z, ["x0", "dw"])
    ec = cfg["eqn_config"]
    r = eng.critic_step(eng.tensor(z["theta_actor"]), None, None, x0, dw, None, ec["num_time_interval_critic"],
                        ec["total_time_critic"], cheat_control=cheat, propagate_only=True,
                        want=("x_smp", "dt", "coef", "exit_index"))
    tag = "cheat" if cheat else "nn"
    coef_ref, dt_ref, x_ref = z[f"prop_{tag}_coef"], z[f"prop_{tag}_dt"], z[f"prop_{tag}_x"]
    coef = npy(r["coef"])
    if dtype == "float64":
        assert np.array_equal(coef, coef_ref)                                   # exit pattern bit-exact
        assert np.array_equal(npy(r["exit_index"]), coef_ref.sum(1))
        np.testing.assert_allclose(npy(r["dt"]), dt_ref, rtol=1e-12, atol=0)
        np.testing.assert_allclose(npy(r["x_smp"]), x_ref, **VTOL[dtype])
    else:
        # float32 vs the float64 reference: a path may flip its exit step only if it passed within
        # rounding distance of the boundary; none of the fixture paths does
        same = (coef == coef_ref).all(1)
        assert same.mean() >= 0.95
        np.testing.assert_allclose(npy(r["dt"])[same], dt_ref[same], rtol=5e-3, atol=1e-7)
        np.testing.assert_allclose(npy(r["x_smp"])[same], x_ref[same], rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cheat", [False, True])
@pytest.mark.parametrize("name", CASES)
def test_critic(name, cheat, dtype):
    z, cfg = load(name)
    eng = engine_for(cfg, dtype)
    x0, dw, xb = dev(eng, z, ["x0", "dw", "xb"])
    thA, thV, thG = dev(eng, z, ["theta_actor", "theta_critic", "theta_critic_grad"])
    ec = cfg["eqn_config"]
    r = eng.critic_step(thA, thV, thG, x0, dw, xb, ec["num_time_interval_critic"], ec["total_time_critic"],
                        cheat_control=cheat, need_grad=True, want=("delta", "delta_bdry", "coef"))
    tag = "cheat" if cheat else "nn"
    same = (npy(r["coef"]) == z[f"prop_{tag}_coef"]).all(1)
    assert same.all() or dtype == "float32"
    np.testing.assert_allclose(npy(r["delta"])[same], z[f"critic_{tag}_delta"][same], **VTOL[dtype])
    np.testing.assert_allclose(npy(r["delta_bdry"]), z[f"critic_{tag}_delta_bdry"], **VTOL[dtype])
    if same.all():
        loss = npy(r["loss"]).sum()
        np.testing.assert_allclose(loss, float(z[f"critic_{tag}_loss"]), rtol=1e-9 if dtype == "float64" else 1e-3)
        assert_grad_close(npy(r["grad_V"]), z[f"critic_{tag}_grad_V"], GTOL[dtype], "grad_V")
        assert_grad_close(npy(r["grad_G"]), z[f"critic_{tag}_grad_G"], GTOL[dtype], "grad_G")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cheat_v", [False, True])
@pytest.mark.parametrize("name", CASES)
def test_actor(name, cheat_v, dtype):
    z, cfg = load(name)
    eng = engine_for(cfg, dtype)
    x0, dw = dev(eng, z, ["x0", "dw"])
    thA, thV = dev(eng, z, ["theta_actor", "theta_critic"])
    ec = cfg["eqn_config"]
    r = eng.actor_step(thA, thV, x0, dw, ec["num_time_interval_actor"], ec["total_time_actor"], cheat_value=cheat_v,
                       need_grad=True, want=("delta", "coef"))
    tag = "cheatV" if cheat_v else "nn"
    same = (npy(r["coef"]) == z["prop_nn_coef"]).all(1)
    assert same.all() or dtype == "float32"
    np.testing.assert_allclose(npy(r["delta"])[same], z[f"actor_{tag}_y"][same], **VTOL[dtype])
    if same.all():
        np.testing.assert_allclose(float(npy(r["loss"])[0]), float(z[f"actor_{tag}_loss"]), rtol=1e-9 if dtype == "float64" else 1e-3, atol=1e-5)
        assert_grad_close(npy(r["grad_actor"]), z[f"actor_{tag}_grad"], GTOL[dtype], "grad_actor")
    # true loss: cheat value + cheat control (solver.py:42)
    r = eng.actor_step(None, None, x0, dw, ec["num_time_interval_actor"], ec["total_time_actor"], cheat_value=True, cheat_control=True)
    np.testing.assert_allclose(float(npy(r["loss"])[0]), float(z["actor_true_loss"]), rtol=1e-9 if dtype == "float64" else 1e-3, atol=1e-5)


@pytest.mark.parametrize("name", CASES)
def test_solver_api_three_adam_iterations(name):
    """ActorCriticSolver (drop-in API) on fixed data: err_* metrics, then three train iterations
    = Keras Adam + PiecewiseConstantDecay semantics (solver.py:16-21,99-107), float64."""
    from deeppde_actorcritic_b200 import equation, munchify
    from deeppde_actorcritic_b200.solver import ActorCriticSolver
    z, cfg = load(name)
    config = munchify(cfg)
    bsde = getattr(equation, config.eqn_config.eqn_name)(config.eqn_config)
    s = ActorCriticSolver(config, bsde, compute_dtype="float64", seed=1)
    s.model_actor.NN_control.theta.copy_(s.engine.tensor(z["theta_actor"]))
    s.model_critic.NN_value.theta.copy_(s.engine.tensor(z["theta_critic"]))
    s.model_critic.NN_value_grad.theta.copy_(s.engine.tensor(z["theta_critic_grad"]))
    inputs = (z["x0"], z["dw"], z["xb"])
    for k in ("err_value", "err_control", "err_value_grad", "err_value_infty", "err_cost"):
        np.testing.assert_allclose(float(getattr(s, k)(inputs)), float(z[k]), rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(float(s.loss_critic(inputs, False, False)), float(z["critic_nn_loss"]), rtol=1e-9)
    np.testing.assert_allclose(float(s.loss_actor(inputs, False, False, False)), float(z["actor_nn_loss"]), rtol=1e-9)
    delta, delta_b = s.model_critic(inputs, s.model_actor, False, False)
    np.testing.assert_allclose(npy(delta), z["critic_nn_delta"], rtol=1e-9, atol=1e-11)
    xs, dt, coef = s.model_critic.propagate(len(z["x0"]), z["x0"], z["dw"], s.model_actor.NN_control, False,
                                            cfg["eqn_config"]["total_time_critic"], cfg["eqn_config"]["num_time_interval_critic"], False)
    assert np.array_equal(npy(coef), z["prop_nn_coef"])
    for _ in range(3):
        s.train_step_critic(inputs)
        s.train_step_actor(inputs)
    for k, t in (("actor", s.model_actor.NN_control.theta), ("critic", s.model_critic.NN_value.theta),
                 ("critic_grad", s.model_critic.NN_value_grad.theta)):
        np.testing.assert_allclose(npy(t), z["theta_after3_" + k], rtol=1e-6, atol=1e-9)


# ----------------------------------------------------------------------------------------------
# Larger seeded runs against the oracle evaluated here (several tiles, ragged tail, wide layers)
BIG = {
    "lqr": ({"eqn_name": "LQR", "discount": 1.0, "p": 1.0, "q": 1.0, "beta": 1.0, "R": 1.0, "dim": 5, "control_dim": 5}, [40, 24], "naive", "TD1", 0.2, 10),
    "vdp": ({"eqn_name": "VDP", "discount": 1.0, "a": 1.0, "epsilon": 0.1, "q": 1.0, "R": 1.0, "dim": 10, "control_dim": 5}, [50, 50], "adaptive", "TD2", 0.6, 12),
    "ekn": ({"eqn_name": "EKN", "discount": 0, "a2": 1.2, "a3": 0.2, "R": 1.0, "dim": 7, "control_dim": 7}, [32, 32, 32], "adaptive", "TD1", 0.5, 10),
    "lqr_var": ({"eqn_name": "LQR_var", "discount": 1.0, "q": 1.0, "beta": 1.0, "epsilon": 0.05, "R": 1.0, "dim": 20, "control_dim": 20}, [200, 200, 200], "adaptive", "TD1", 0.2, 8),
}


@pytest.mark.parametrize("key", list(BIG))
def test_multi_tile_vs_oracle_f64(key):
    e, hid, scheme, td, T, N = BIG[key]
    e = dict(e, total_time_critic=T, total_time_actor=T, num_time_interval_critic=N, num_time_interval_actor=N)
    B = 77                                                         # 5 tiles of 16 paths, ragged tail
    cfg = {"eqn_config": e, "net_config": {"num_hiddens_actor": hid, "num_hiddens_critic": hid},
           "train_config": {"scheme": scheme, "TD_type": td, "sample_type": "normal", "train": "actor-critic"}}
    eqn = RE.make_ref_equation(e)
    np.random.seed(5)
    x0, dw, xb = eqn.sample_normal(B, N)
    rng = np.random.RandomState(9)
    th = {}
    for k in ("actor", "critic", "critic_grad"):
        i, h, o, _ = RS.net_dims(cfg, k)
        th[k] = RS.init_params(i, h, o, rng)
        th[k][-3 * o:-2 * o] = rng.normal(0, 0.1, o)
    tt = {k: torch.tensor(v) for k, v in th.items()}
    inputs = tuple(torch.tensor(a) for a in (x0, dw, xb))
    loss_c, gV, gG, delta, delta_b, aux = RS.grad_critic(eqn, cfg, tt, inputs, False)
    loss_a, gA, y, _ = RS.grad_actor(eqn, cfg, tt, inputs, False, False)

    eng = engine_for(cfg, "float64")
    d = [eng.tensor(a) for a in (x0, dw, xb)]
    thd = {k: eng.tensor(v) for k, v in th.items()}
    r = eng.critic_step(thd["actor"], thd["critic"], thd["critic_grad"], d[0], d[1], d[2], N, T, need_grad=True,
                        want=("delta", "delta_bdry", "coef", "dt"))
    assert np.array_equal(npy(r["coef"]), aux["coef"].numpy())
    np.testing.assert_allclose(npy(r["dt"]), aux["dt"].numpy(), rtol=1e-12)
    np.testing.assert_allclose(npy(r["delta"]), delta.numpy(), rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(npy(r["delta_bdry"]), delta_b.numpy(), rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(npy(r["loss"]).sum(), float(loss_c), rtol=1e-9)
    assert_grad_close(npy(r["grad_V"]), gV.numpy(), 1e-7, "grad_V")
    assert_grad_close(npy(r["grad_G"]), gG.numpy(), 1e-7, "grad_G")
    r = eng.actor_step(thd["actor"], thd["critic"], d[0], d[1], N, T, need_grad=True, want=("delta",))
    np.testing.assert_allclose(npy(r["delta"]), y.numpy(), rtol=1e-8, atol=1e-10)
    assert_grad_close(npy(r["grad_actor"]), gA.numpy(), 1e-7, "grad_actor")
    # sharding invariance: two shards with B_global = B sum to the whole
    h = 40
    parts = []
    for lo, hi in ((0, h), (h, B)):
        sl = [t[lo:hi].contiguous() for t in d]
        parts.append(eng.critic_step(thd["actor"], thd["critic"], thd["critic_grad"], sl[0], sl[1], sl[2], N, T, need_grad=True,
                                     B_global=B, path_offset=lo))
    np.testing.assert_allclose(npy(parts[0]["loss"] + parts[1]["loss"]).sum(), float(loss_c), rtol=1e-9)
    assert_grad_close(npy(parts[0]["grad_G"] + parts[1]["grad_G"]), gG.numpy(), 1e-7, "sharded grad_G")
    assert_grad_close(npy(parts[0]["grad_V"] + parts[1]["grad_V"]), gV.numpy(), 1e-7, "sharded grad_V")


# ----------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def hh():
    src = os.path.join(ROOT, "tests", "harness", "host_harness.cpp")
    so = os.path.join(ROOT, "tests", "harness", "_host_harness.so")
    if not os.path.exists(so):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src])
    return C.CDLL(so)


@pytest.mark.parametrize("key", list(BIG))
def test_f32_schedule_bit_exact_under_true_control(key, hh):
    """float32 kernel vs the float32 g++ build of the same per-path code, u = u_true: exit indices,
    coef and the adaptive dt schedule agree bit for bit; x agrees bit for bit."""
    e, hid, scheme, td, T, N = BIG[key]
    T, N = T * 4, 40
    e = dict(e, total_time_critic=T, total_time_actor=T, num_time_interval_critic=N, num_time_interval_actor=N)
    cfg = {"eqn_config": e, "net_config": {"num_hiddens_actor": [8], "num_hiddens_critic": [8]},
           "train_config": {"scheme": scheme, "TD_type": td}}
    B = 300
    eqn = RE.make_ref_equation(e)
    np.random.seed(3)
    x0, dw, _ = eqn.sample_normal(B, N)
    x0, dw = x0.astype(np.float32), dw.astype(np.float32)
    eng = engine_for(cfg, "float32")
    r = eng.critic_step(None, None, None, eng.tensor(x0), eng.tensor(dw), None, N, T, cheat_control=True, propagate_only=True,
                        want=("x_smp", "dt", "coef", "exit_index"))
    d, m = e["dim"], e["control_dim"]
    A = np.zeros((m, d), np.float32)
    b = np.zeros(m, np.float32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    xs_g, dt_g, cf_g = r["x_smp"].cpu().numpy(), r["dt"].cpu().numpy(), r["coef"].cpu().numpy()
    for i in range(B):
        xs, dts, cfs, y = np.zeros((d, N + 1), np.float32), np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros(1, np.float32)
        hh.hh_run_path_f32(C.byref(eng.cfg), N, C.c_double(T), P(A), P(b), P(np.ascontiguousarray(x0[i])), P(np.ascontiguousarray(dw[i])),
                           1, C.c_float(1.0), P(xs), P(dts), P(cfs), P(y), None, None)
        assert np.array_equal(cfs, cf_g[i]), f"path {i}: coef differs"
        nlive = int(cfs.sum())
        assert int(r["exit_index"][i]) == nlive
        assert np.array_equal(dts[:nlive].view(np.uint32), dt_g[i, :nlive].view(np.uint32)), f"path {i}: dt bits differ"
        assert np.array_equal(xs.view(np.uint32), xs_g[i].view(np.uint32)), f"path {i}: x bits differ"
    assert 0.02 < cf_g.mean() < 0.999


def test_philox_increments():
    e, hid, scheme, td, T, N = BIG["lqr_var"]
    N = 16
    e = dict(e, total_time_critic=T, total_time_actor=T, num_time_interval_critic=N, num_time_interval_actor=N)
    cfg = {"eqn_config": e, "net_config": {"num_hiddens_actor": [24, 24], "num_hiddens_critic": [24, 24]},
           "train_config": {"scheme": "adaptive", "TD_type": "TD1"}}
    eng = engine_for(cfg, "float32")
    B = 4096
    dwn = eng.philox_dw(_cabi.DW_PHILOX_NORMAL, 2024, 6, 0, B, N)
    a = dwn.double()
    assert abs(float(a.mean())) < 5e-3 and abs(float(a.var()) - 1) < 1e-2 and abs(float((a ** 4).mean()) - 3) < 0.1
    dwb = eng.philox_dw(_cabi.DW_PHILOX_BOUNDED, 2024, 7, 0, B, N)
    vals, counts = torch.unique(dwb, return_counts=True)
    np.testing.assert_allclose(vals.cpu().numpy(), [-np.sqrt(3.0), 0, np.sqrt(3.0)], rtol=1e-6)       # equation.py:31-32
    np.testing.assert_allclose((counts.double() / dwb.numel()).cpu().numpy(), [1 / 6, 4 / 6, 1 / 6], atol=3e-3)
    # sharding invariance of the generator: rows [lo, hi) of the global tensor == a shard generated at offset lo
    sh = eng.philox_dw(_cabi.DW_PHILOX_NORMAL, 2024, 6, 1000, 500, N)
    assert torch.equal(sh, dwn[1000:1500])
    # in-kernel generation == the materialised tensor fed externally (bitwise identical results)
    rng = np.random.RandomState(4)
    th = {}
    for k in ("actor", "critic", "critic_grad"):
        i, h, o, _ = RS.net_dims(cfg, k)
        th[k] = eng.tensor(RS.init_params(i, h, o, rng))
    x0, xb = eng.sample_x(2024, 6, 0, B)
    r = torch.linalg.norm(x0.double(), dim=1)
    assert float(r.max()) < 1.0 and abs(float((r ** e["dim"]).mean()) - 0.5) < 2e-2        # r^d uniform on (0,1)
    np.testing.assert_allclose(torch.linalg.norm(xb.double(), dim=1).cpu().numpy(), 1.0, rtol=1e-5)
    for mode, dw in ((_cabi.DW_PHILOX_NORMAL, dwn),):
        r1 = eng.critic_step(th["actor"], th["critic"], th["critic_grad"], x0, None, xb, N, T, need_grad=True, want=("delta",),
                             dw_mode=mode, seed=2024, stream_id=6)
        r2 = eng.critic_step(th["actor"], th["critic"], th["critic_grad"], x0, dw, xb, N, T, need_grad=True, want=("delta",))
        assert torch.equal(r1["delta"], r2["delta"]) and torch.equal(r1["grad_G"], r2["grad_G"]) and torch.equal(r1["loss"], r2["loss"])
        a1 = eng.actor_step(th["actor"], th["critic"], x0, None, N, T, need_grad=True, dw_mode=mode, seed=2024, stream_id=6)
        a2 = eng.actor_step(th["actor"], th["critic"], x0, dw, N, T, need_grad=True)
        assert torch.equal(a1["grad_actor"], a2["grad_actor"]) and torch.equal(a1["loss"], a2["loss"])


def test_error_paths():
    cfg = {"eqn_config": dict(BIG["lqr"][0], total_time_critic=0.2), "net_config": {"num_hiddens_actor": [8], "num_hiddens_critic": [8]},
           "train_config": {"scheme": "naive", "TD_type": "TD1"}}
    eng = engine_for(cfg, "float32")
    x0 = torch.zeros(4, 5, device="cuda")
    with pytest.raises(_cabi.DpbError):
        eng.critic_step(None, None, None, x0, None, None, 10, 0.2, propagate_only=True)          # actor weights missing
    with pytest.raises(ValueError):
        Engine(dict(cfg["eqn_config"], eqn_name="nope"), cfg["net_config"], cfg["train_config"])
    with pytest.raises(_cabi.DpbError):
        Engine(dict(cfg["eqn_config"], control_dim=3), cfg["net_config"], cfg["train_config"])


def test_checkpoint_resume_is_bit_exact(tmp_path):
    """train 6 iterations == train 3, checkpoint, reload into a fresh solver, train 3 (device sampling is keyed
    by (seed, iteration); exact path => bitwise equal weights)."""
    from deeppde_actorcritic_b200 import equation, munchify
    from deeppde_actorcritic_b200.solver import ActorCriticSolver
    z, cfg = load("lqr_d5_adaptive_normal_td1")
    cfg = json.loads(json.dumps(cfg))
    cfg["net_config"]["batch_size"] = 96

    def make():
        config = munchify(cfg)
        bsde = getattr(equation, config.eqn_config.eqn_name)(config.eqn_config)
        return ActorCriticSolver(config, bsde, compute_dtype="float32", seed=3, impl="exact")

    a = make()
    for _ in range(6):
        a.train_iteration()
    b = make()
    for _ in range(3):
        b.train_iteration()
    path = str(tmp_path / "ck.pt")
    b.save_checkpoint(path)
    c = make()
    c.load_checkpoint(path)
    for _ in range(3):
        c.train_iteration()
    for ta, tc in ((a.model_actor.NN_control.theta, c.model_actor.NN_control.theta),
                   (a.model_critic.NN_value.theta, c.model_critic.NN_value.theta),
                   (a.model_critic.NN_value_grad.theta, c.model_critic.NN_value_grad.theta)):
        assert torch.equal(ta, tc)
