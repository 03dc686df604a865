"""tcgen05 building blocks of the tensor path against plain float64 matmuls (GPU only)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from deeppde_actorcritic_b200 import _cabi


@pytest.mark.parametrize("K", [208, 128])
def test_tcgen05_product_forms(K):
    lib = _cabi.load()
    g = torch.Generator().manual_seed(K)
    A = torch.randn(128, K, generator=g).bfloat16().float()
    B = torch.randn(208, K, generator=g).bfloat16().float()
    Ad, Bd = A.cuda(), B.cuda()
    D = torch.full((3, 128, 208), float("nan"), device="cuda")
    rc = lib.dpb_tc_selftest(Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), K, None)
    assert rc == 0, lib.dpb_last_error(None)
    torch.cuda.synchronize()
    D = D.cpu().double()
    ref = A.double() @ B.double().T
    err = [(D[i] - ref).abs().max().item() for i in (0, 1)]
    Cm = torch.zeros(128, 208, dtype=torch.float64)
    Cm[:, :K] = B[:128, :K].double()
    ref3 = A[:, :128].double().T @ Cm
    err.append((D[2] - ref3).abs().max().item())
    print("tcgen05 selftest max abs errors (SS K-major, TS, SS MN-major):", err)
    tol = 1e-3 * (K ** 0.5)
    assert err[0] < tol, f"SS K-major product wrong: {err}"
    assert err[1] < tol, f"TS (A from TMEM) product wrong: {err}"
    assert err[2] < tol, f"SS MN-major product wrong: {err}"


def test_tcgen05_handshake_latency():
    """empty control <-> path round trip of the tensor path's protocol completes and is timed (diagnostic)"""
    lib = _cabi.load()
    out = np.zeros(2, dtype=np.int64)
    rc = lib.dpb_tc_handshake_cycles(out.ctypes.data_as(C.c_void_p), 2000)
    assert rc == 0, lib.dpb_last_error(None)
    print("tcgen05 hand-off round trip:", int(out[0]), "cycles")
    assert out[1] == 2000 and 0 < out[0] < 100000
    for g in (1, 2, 3, 4):
        for with_mma, publish in ((0, 2), (0, 0), (0, 1), (1, 2), (1, 0), (1, 1)):
            assert lib.dpb_tc_epilogue_cycles(out.ctypes.data_as(C.c_void_p), 500, g, with_mma, publish) == 0, lib.dpb_last_error(None)
            print(f"hidden-layer epilogue (200 wide) with {4 * g} warps, tensor pipe {'busy' if with_mma else 'idle'}, "
                  f"{'per-chunk publish' if publish & 1 else 'no publish'}, {'separate plane columns' if publish & 2 else 'in place'}: {int(out[0])} cycles")
            assert 0 < out[0] < 1000000


# ------------------------------------------------------------------------------------------------
# tensor path (impl="tensor": bf16x3 products on tcgen05, FP32 accumulation) vs golden / exact path
import glob
import json
import os

from deeppde_actorcritic_b200.engine import Engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if "samplers" not in p)
VTOL = dict(rtol=3e-4, atol=3e-4)          # stated tolerance of the tensor path: values


def _load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z, json.loads(str(z["config_json"]))


def _npy(t):
    return t.detach().cpu().double().numpy()


@pytest.mark.parametrize("name", CASES)
def test_tensor_forward_vs_golden(name):
    z, cfg = _load(name)
    eng = Engine(cfg["eqn_config"], cfg["net_config"], cfg["train_config"], dtype="float32", impl="tensor")
    x0, dw, xb = (eng.tensor(z[k]) for k in ("x0", "dw", "xb"))
    thA, thV, thG = (eng.tensor(z[k]) for k in ("theta_actor", "theta_critic", "theta_critic_grad"))
    ec = cfg["eqn_config"]
    N, T = ec["num_time_interval_critic"], ec["total_time_critic"]
    r = eng.critic_step(thA, None, None, x0, dw, None, N, T, propagate_only=True, want=("x_smp", "dt", "coef", "exit_index"))
    same = (_npy(r["coef"]) == z["prop_nn_coef"]).all(1)
    assert same.mean() >= 0.95
    np.testing.assert_allclose(_npy(r["x_smp"])[same], z["prop_nn_x"][same], rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(_npy(r["dt"])[same], z["prop_nn_dt"][same], rtol=5e-3, atol=1e-7)
    for cheat in (False, True):
        tag = "cheat" if cheat else "nn"
        r = eng.critic_step(thA, thV, thG, x0, dw, xb, N, T, cheat_control=cheat, want=("delta", "delta_bdry", "coef"))
        same = (_npy(r["coef"]) == z[f"prop_{tag}_coef"]).all(1)
        np.testing.assert_allclose(_npy(r["delta"])[same], z[f"critic_{tag}_delta"][same], **VTOL)
        np.testing.assert_allclose(_npy(r["delta_bdry"]), z[f"critic_{tag}_delta_bdry"], **VTOL)
        if same.all():
            np.testing.assert_allclose(_npy(r["loss"]).sum(), float(z[f"critic_{tag}_loss"]), rtol=1e-3)
    for cheat_v in (False, True):
        tag = "cheatV" if cheat_v else "nn"
        r = eng.actor_step(thA, thV, x0, dw, N, T, cheat_value=cheat_v, want=("delta", "coef"))
        same = (_npy(r["coef"]) == z["prop_nn_coef"]).all(1)
        np.testing.assert_allclose(_npy(r["delta"])[same], z[f"actor_{tag}_y"][same], **VTOL)
        if same.all():
            np.testing.assert_allclose(float(_npy(r["loss"])[0]), float(z[f"actor_{tag}_loss"]), rtol=1e-3, atol=1e-5)


def test_tensor_forward_vs_exact_d20_3x200():
    """lqr_var d=20, nets 3x200, several tiles with a ragged tail, in-kernel Philox increments:
    tensor path against the exact FP32 path on identical inputs."""
    e = {"eqn_name": "LQR_var", "discount": 1.0, "q": 1.0, "beta": 1.0, "epsilon": 0.05, "R": 1.0, "dim": 20, "control_dim": 20,
         "total_time_critic": 0.2, "total_time_actor": 0.2, "num_time_interval_critic": 25, "num_time_interval_actor": 25}
    net = {"num_hiddens_actor": [200, 200, 200], "num_hiddens_critic": [200, 200, 200]}
    tr = {"scheme": "adaptive", "TD_type": "TD1"}
    ex = Engine(e, net, tr, dtype="float32", impl="exact")
    tn = Engine(e, net, tr, dtype="float32", impl="tensor")
    from oracle import ref_solver as RS
    rng = np.random.RandomState(12)
    cfg = {"eqn_config": e, "net_config": net, "train_config": tr}
    th = {}
    for k in ("actor", "critic", "critic_grad"):
        i, h, o, _ = RS.net_dims(cfg, k)
        th[k] = ex.tensor(RS.init_params(i, h, o, rng))
    B, N, T = 1000, 25, 0.2
    x0, xb = ex.sample_x(5, 1, 0, B)
    kw = dict(dw_mode=1, seed=5, stream_id=3)
    a = ex.critic_step(th["actor"], th["critic"], th["critic_grad"], x0, None, xb, N, T, want=("delta", "delta_bdry", "coef", "exit_index"), **kw)
    b = tn.critic_step(th["actor"], th["critic"], th["critic_grad"], x0, None, xb, N, T, want=("delta", "delta_bdry", "coef", "exit_index"), **kw)
    same = (a["coef"] == b["coef"]).all(1).cpu().numpy()
    print("tensor vs exact: identical exit pattern on", same.mean(), "of paths; max |d delta| =",
          float((a["delta"] - b["delta"]).abs().cpu().numpy()[same].max()))
    assert same.mean() > 0.99
    np.testing.assert_allclose(_npy(b["delta"])[same], _npy(a["delta"])[same], **VTOL)
    np.testing.assert_allclose(_npy(b["delta_bdry"]), _npy(a["delta_bdry"]), **VTOL)
    ya = ex.actor_step(th["actor"], th["critic"], x0, None, N, T, want=("delta", "coef"), **kw)
    yb = tn.actor_step(th["actor"], th["critic"], x0, None, N, T, want=("delta", "coef"), **kw)
    same = (ya["coef"] == yb["coef"]).all(1).cpu().numpy()
    np.testing.assert_allclose(_npy(yb["delta"])[same], _npy(ya["delta"])[same], **VTOL)


GTOL_TENSOR = 2e-3      # stated tolerance of the tensor path: gradients, relative to the gradient's max-norm -- the SAME as the
                        # exact FP32 path's (test_gpu_parity.py).  Forward / dX products are bf16x3 (16 significant bits), the dW
                        # contraction reads FP16 operand images (11 bits, per-evaluation power-of-two scaling); observed on these
                        # fixtures (24 paths): <= 4.2e-4, at the BASELINE shapes (1024 paths): <= 2.1e-4


def _gerr(got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    scale = np.abs(ref).max()
    if scale == 0:
        return float(np.abs(got).max())
    return float(np.abs(got - ref).max() / scale)


@pytest.mark.parametrize("name", CASES)
def test_tensor_gradients_vs_golden(name):
    z, cfg = _load(name)
    eng = Engine(cfg["eqn_config"], cfg["net_config"], cfg["train_config"], dtype="float32", impl="tensor")
    x0, dw, xb = (eng.tensor(z[k]) for k in ("x0", "dw", "xb"))
    thA, thV, thG = (eng.tensor(z[k]) for k in ("theta_actor", "theta_critic", "theta_critic_grad"))
    ec = cfg["eqn_config"]
    N, T = ec["num_time_interval_critic"], ec["total_time_critic"]
    errs = {}
    for cheat in (False, True):
        tag = "cheat" if cheat else "nn"
        r = eng.critic_step(thA, thV, thG, x0, dw, xb, N, T, cheat_control=cheat, need_grad=True, want=("delta", "coef"))
        same = (_npy(r["coef"]) == z[f"prop_{tag}_coef"]).all(1)
        np.testing.assert_allclose(_npy(r["delta"])[same], z[f"critic_{tag}_delta"][same], **VTOL)
        if same.all():
            errs[f"gV_{tag}"] = _gerr(_npy(r["grad_V"]), z[f"critic_{tag}_grad_V"])
            errs[f"gG_{tag}"] = _gerr(_npy(r["grad_G"]), z[f"critic_{tag}_grad_G"])
    for cheat_v in (False, True):
        tag = "cheatV" if cheat_v else "nn"
        r = eng.actor_step(thA, thV, x0, dw, N, T, cheat_value=cheat_v, need_grad=True, want=("delta", "coef"))
        same = (_npy(r["coef"]) == z["prop_nn_coef"]).all(1)
        np.testing.assert_allclose(_npy(r["delta"])[same], z[f"actor_{tag}_y"][same], **VTOL)
        if same.all():
            errs[f"gA_{tag}"] = _gerr(_npy(r["grad_actor"]), z[f"actor_{tag}_grad"])
    print(name, {k: f"{v:.2e}" for k, v in errs.items()})
    for k, v in errs.items():
        assert v < GTOL_TENSOR, (k, v)


def test_tensor_gradients_vs_exact_d20_3x200():
    e = {"eqn_name": "LQR_var", "discount": 1.0, "q": 1.0, "beta": 1.0, "epsilon": 0.05, "R": 1.0, "dim": 20, "control_dim": 20,
         "total_time_critic": 0.2, "total_time_actor": 0.2, "num_time_interval_critic": 25, "num_time_interval_actor": 25}
    net = {"num_hiddens_actor": [200, 200, 200], "num_hiddens_critic": [200, 200, 200]}
    tr = {"scheme": "adaptive", "TD_type": "TD1"}
    ex = Engine(e, net, tr, dtype="float32", impl="exact")
    tn = Engine(e, net, tr, dtype="float32", impl="tensor")
    from oracle import ref_solver as RS
    rng = np.random.RandomState(12)
    cfg = {"eqn_config": e, "net_config": net, "train_config": tr}
    th = {}
    for k in ("actor", "critic", "critic_grad"):
        i, h, o, _ = RS.net_dims(cfg, k)
        th[k] = ex.tensor(RS.init_params(i, h, o, rng))
    B, N, T = 1000, 25, 0.2
    x0, xb = ex.sample_x(5, 1, 0, B)
    kw = dict(dw_mode=1, seed=5, stream_id=3)
    a = ex.critic_step(th["actor"], th["critic"], th["critic_grad"], x0, None, xb, N, T, need_grad=True, **kw)
    b = tn.critic_step(th["actor"], th["critic"], th["critic_grad"], x0, None, xb, N, T, need_grad=True, **kw)
    ya = ex.actor_step(th["actor"], th["critic"], x0, None, N, T, need_grad=True, **kw)
    yb = tn.actor_step(th["actor"], th["critic"], x0, None, N, T, need_grad=True, **kw)
    errs = {"gV": _gerr(_npy(b["grad_V"]), _npy(a["grad_V"])), "gG": _gerr(_npy(b["grad_G"]), _npy(a["grad_G"])),
            "gA": _gerr(_npy(yb["grad_actor"]), _npy(ya["grad_actor"])),
            "loss_c": abs(float(b["loss"].sum() - a["loss"].sum())) / abs(float(a["loss"].sum())),
            "loss_a": abs(float(yb["loss"][0] - ya["loss"][0]))}
    print("tensor vs exact gradients (max-norm relative):", {k: f"{v:.2e}" for k, v in errs.items()})
    assert errs["gV"] < GTOL_TENSOR and errs["gG"] < GTOL_TENSOR and errs["gA"] < GTOL_TENSOR
    assert errs["loss_c"] < 1e-4 and errs["loss_a"] < 1e-4


ODD = {
    # specialised <24, LQR> kernels, widths that are not multiples of 16, three tiles with a ragged tail
    "lqr_d13_w72_40": ({"eqn_name": "LQR", "discount": 1.0, "p": 1.0, "q": 1.0, "beta": 1.0, "R": 1.0, "dim": 13, "control_dim": 13}, [72, 40], [40, 72, 24], "adaptive", "TD1"),
    # generic <0,-1> kernels (VDP indexes its state cyclically), hidden width 50 as configs/vdp_d4.json
    "vdp_d10_w50": ({"eqn_name": "VDP", "discount": 1.0, "a": 1.0, "epsilon": 0.1, "q": 1.0, "R": 1.0, "dim": 10, "control_dim": 5}, [50, 50], [50, 50], "naive", "TD1"),
    # one hidden layer (actor) / four hidden layers (critic)
    "lqr_var_d6_L1_L4": ({"eqn_name": "LQR_var", "discount": 1.0, "q": 1.0, "beta": 1.0, "epsilon": 0.05, "R": 1.0, "dim": 6, "control_dim": 6}, [48], [32, 48, 32, 16], "adaptive", "TD1"),
    # ekn head (control_dim + 1 outputs) and a 255-wide layer: the widest the tensor path supports
    "ekn_d9_w255": ({"eqn_name": "EKN", "discount": 0, "a2": 1.2, "a3": 0.2, "R": 1.0, "dim": 9, "control_dim": 9}, [255, 64], [64, 255], "adaptive", "TD1"),
}


@pytest.mark.parametrize("key", list(ODD))
def test_tensor_vs_exact_odd_shapes(key):
    e, ha, hc, scheme, td = ODD[key]
    N, T, B = 12, 0.3, 300
    e = dict(e, total_time_critic=T, total_time_actor=T, num_time_interval_critic=N, num_time_interval_actor=N)
    net = {"num_hiddens_actor": ha, "num_hiddens_critic": hc}
    tr = {"scheme": scheme, "TD_type": td}
    ex = Engine(e, net, tr, dtype="float32", impl="exact")
    tn = Engine(e, net, tr, dtype="float32", impl="tensor")
    from oracle import ref_solver as RS
    rng = np.random.RandomState(3)
    cfg = {"eqn_config": e, "net_config": net, "train_config": tr}
    th = {}
    for k in ("actor", "critic", "critic_grad"):
        i, h, o, _ = RS.net_dims(cfg, k)
        p = RS.init_params(i, h, o, rng)
        p[-3 * o:-2 * o] = rng.normal(0, 0.1, o)
        th[k] = ex.tensor(p)
    x0, xb = ex.sample_x(11, 1, 0, B)
    x0 = x0 * 0.8
    kw = dict(dw_mode=2 if key.startswith("vdp") else 1, seed=11, stream_id=5)
    a = ex.critic_step(th["actor"], th["critic"], th["critic_grad"], x0, None, xb, N, T, need_grad=True, want=("delta", "delta_bdry", "coef"), **kw)
    b = tn.critic_step(th["actor"], th["critic"], th["critic_grad"], x0, None, xb, N, T, need_grad=True, want=("delta", "delta_bdry", "coef"), **kw)
    same = (a["coef"] == b["coef"]).all(1).cpu().numpy()
    assert same.mean() > 0.98
    np.testing.assert_allclose(_npy(b["delta"])[same], _npy(a["delta"])[same], **VTOL)
    np.testing.assert_allclose(_npy(b["delta_bdry"]), _npy(a["delta_bdry"]), **VTOL)
    ya = ex.actor_step(th["actor"], th["critic"], x0, None, N, T, need_grad=True, want=("delta", "coef"), **kw)
    yb = tn.actor_step(th["actor"], th["critic"], x0, None, N, T, need_grad=True, want=("delta", "coef"), **kw)
    same_a = (ya["coef"] == yb["coef"]).all(1).cpu().numpy()
    np.testing.assert_allclose(_npy(yb["delta"])[same_a], _npy(ya["delta"])[same_a], **VTOL)
    if same.all() and same_a.all():
        errs = {"gV": _gerr(_npy(b["grad_V"]), _npy(a["grad_V"])), "gG": _gerr(_npy(b["grad_G"]), _npy(a["grad_G"])),
                "gA": _gerr(_npy(yb["grad_actor"]), _npy(ya["grad_actor"]))}
        print(key, {k: f"{v:.2e}" for k, v in errs.items()})
        for k, v in errs.items():
            assert v < GTOL_TENSOR, (k, v)


def test_tensor_sharding_invariance_philox():
    """Two shards (path_offset, B_global) with in-kernel Philox increments == the whole batch:
    the generator is keyed by the GLOBAL path index; sums agree up to FP32 reduction order."""
    e = {"eqn_name": "LQR_var", "discount": 1.0, "q": 1.0, "beta": 1.0, "epsilon": 0.05, "R": 1.0, "dim": 8, "control_dim": 8,
         "total_time_critic": 0.3, "total_time_actor": 0.3, "num_time_interval_critic": 10, "num_time_interval_actor": 10}
    net = {"num_hiddens_actor": [40, 40], "num_hiddens_critic": [40, 40]}
    tr = {"scheme": "adaptive", "TD_type": "TD1"}
    tn = Engine(e, net, tr, dtype="float32", impl="tensor")
    from oracle import ref_solver as RS
    rng = np.random.RandomState(5)
    cfg = {"eqn_config": e, "net_config": net, "train_config": tr}
    th = {}
    for k in ("actor", "critic", "critic_grad"):
        i, h, o, _ = RS.net_dims(cfg, k)
        th[k] = tn.tensor(RS.init_params(i, h, o, rng))
    B, N, T, cut = 500, 10, 0.3, 200
    x0, xb = tn.sample_x(9, 4, 0, B)
    kw = dict(dw_mode=1, seed=9, stream_id=8, need_grad=True)
    whole = tn.critic_step(th["actor"], th["critic"], th["critic_grad"], x0, None, xb, N, T, want=("delta",), **kw)
    parts = [tn.critic_step(th["actor"], th["critic"], th["critic_grad"], x0[lo:hi].contiguous(), None, xb[lo:hi].contiguous(), N, T,
                            B_global=B, path_offset=lo, want=("delta",), **kw) for lo, hi in ((0, cut), (cut, B))]
    assert torch.equal(torch.cat([p["delta"] for p in parts]), whole["delta"])             # identical paths, identical arithmetic
    np.testing.assert_allclose(_npy(parts[0]["loss"] + parts[1]["loss"]), _npy(whole["loss"]), rtol=1e-5)
    for g in ("grad_V", "grad_G"):
        assert _gerr(_npy(parts[0][g] + parts[1][g]), _npy(whole[g])) < 1e-4
    wa = tn.actor_step(th["actor"], th["critic"], x0, None, N, T, **kw)
    pa = [tn.actor_step(th["actor"], th["critic"], x0[lo:hi].contiguous(), None, N, T, B_global=B, path_offset=lo, **kw) for lo, hi in ((0, cut), (cut, B))]
    assert _gerr(_npy(pa[0]["grad_actor"] + pa[1]["grad_actor"]), _npy(wa["grad_actor"])) < 1e-4


@pytest.mark.parametrize("impl", ["exact", "tensor"])
@pytest.mark.parametrize("train,scheme,sample,td", [("critic", "naive", "bounded", "TD2"), ("actor", "adaptive", "normal", "TD1"),
                                                    ("actor-critic", "naive", "normal", "TD1")])
def test_solver_modes_run(impl, train, scheme, sample, td):
    """every train / scheme / sample_type / TD_type field of the reference's JSON drives the solver on both
    implementations: a few device-sampled and host-sampled iterations change the right networks and stay finite."""
    from deeppde_actorcritic_b200 import equation, munchify
    from deeppde_actorcritic_b200.solver import ActorCriticSolver
    z, cfg = _load("vdp_d10_adaptive_bounded_td2")
    cfg = json.loads(json.dumps(cfg))
    cfg["train_config"] = {"sample_type": sample, "scheme": scheme, "TD_type": td, "train": train}
    cfg["net_config"]["batch_size"] = 160
    config = munchify(cfg)
    bsde = getattr(equation, config.eqn_config.eqn_name)(config.eqn_config)
    s = ActorCriticSolver(config, bsde, compute_dtype="float32", seed=2, impl=impl)
    before = [t.clone() for t in (s.model_actor.NN_control.theta, s.model_critic.NN_value.theta, s.model_critic.NN_value_grad.theta)]
    for _ in range(2):
        s.train_iteration()                      # device sampling
    s.sampler = "host"
    s.train_iteration()                          # the reference's NumPy samplers
    after = [s.model_actor.NN_control.theta, s.model_critic.NN_value.theta, s.model_critic.NN_value_grad.theta]
    changed = [not torch.equal(a, b) for a, b in zip(before, after)]
    assert all(torch.isfinite(t).all() for t in after)
    assert changed[0] == (train in ("actor", "actor-critic"))
    assert changed[1] == (train in ("critic", "actor-critic"))
    assert changed[2] == (train in ("critic", "actor-critic") and td == "TD1")
    valid = tuple(s.engine.tensor(a) for a in s.sample(64, s.N_c))
    assert np.isfinite(float(s.loss_critic(valid, False, False))) and np.isfinite(float(s.loss_actor(valid, False, False, False)))


# ------------------------------------------------------------------------------------------------
# round-2 additions: variants that had no parity test (VERDICT r01 "test debt")
def _oracle_vs_engines(cfg, B, seed, ekn_sigma_fix=False, impls=(("float64", "exact"), ("float32", "tensor")), cheat=False):
    """critic residuals / losses / gradients and actor cost / gradient of both implementations against the float64 oracle
    on the same (x0, dw, x_bdry) and weights; returns the error table"""
    from oracle import ref_equation as RE
    from oracle import ref_solver as RS
    e = cfg["eqn_config"]
    N, T = int(e["num_time_interval_critic"]), float(e["total_time_critic"])
    eqn = RE.make_ref_equation(e, ekn_sigma_fix=ekn_sigma_fix)
    np.random.seed(seed)
    sampler = eqn.sample_bounded if cfg["train_config"].get("sample_type") == "bounded" else eqn.sample_normal
    x0, dw, xb = sampler(B, N)
    x0, dw, xb = (a.astype(np.float32).astype(np.float64) for a in (x0 * 0.9, dw, xb))
    rng = np.random.RandomState(seed + 1)
    th = {}
    for k in ("actor", "critic", "critic_grad"):
        i, h, o, _ = RS.net_dims(cfg, k)
        p = RS.init_params(i, h, o, rng)
        p[-3 * o:-2 * o] = rng.normal(0, 0.1, o)
        th[k] = p.astype(np.float32).astype(np.float64)
    tt = {k: torch.tensor(v) for k, v in th.items()}
    inputs = tuple(torch.tensor(a) for a in (x0, dw, xb))
    loss_c, gV, gG, delta, delta_b, aux = RS.grad_critic(eqn, cfg, tt, inputs, cheat)
    loss_a, gA, y, aux_a = RS.grad_actor(eqn, cfg, tt, inputs, False, False)
    table = {}
    for dtype, impl in impls:
        eng = Engine(e, cfg["net_config"], cfg["train_config"], dtype=dtype, impl=impl, ekn_sigma_fix=ekn_sigma_fix)
        d = [eng.tensor(a) for a in (x0, dw, xb)]
        thd = {k: eng.tensor(v) for k, v in th.items()}
        r = eng.critic_step(thd["actor"], thd["critic"], thd["critic_grad"], d[0], d[1], d[2], N, T, need_grad=True, cheat_control=cheat,
                            want=("delta", "delta_bdry", "coef"))
        a = eng.actor_step(thd["actor"], thd["critic"], d[0], d[1], N, T, need_grad=True, want=("delta", "coef"))
        same = (_npy(r["coef"]) == aux["coef"].detach().numpy()).all(1)
        same_a = (_npy(a["coef"]) == aux_a["coef"].detach().numpy()).all(1)
        vt = dict(rtol=1e-8, atol=1e-10) if dtype == "float64" else VTOL
        np.testing.assert_allclose(_npy(r["delta"])[same], delta.numpy()[same], **vt)
        np.testing.assert_allclose(_npy(r["delta_bdry"]), delta_b.numpy(), **vt)
        np.testing.assert_allclose(_npy(a["delta"])[same_a], y.numpy()[same_a], **vt)
        table[impl] = {"same": float(same.mean()), "same_a": float(same_a.mean()),
                       "gV": _gerr(_npy(r["grad_V"]), gV.numpy()), "gG": _gerr(_npy(r["grad_G"]), gG.numpy()),
                       "gA": _gerr(_npy(a["grad_actor"]), gA.numpy()), "all_same": bool(same.all() and same_a.all())}
    return table


def test_ekn_sigma_fix_vs_oracle():
    """the consistent-sigma switch of ekn (sigma = sqrt(2 eps) instead of the reference's literal sqrt(2)) against
    RefEKN(sigma_fix=True) -- and the literal setting must differ from it (the switch really changes the dynamics)"""
    z, cfg = _load("ekn_d6_adaptive_normal_td1")
    cfg = json.loads(json.dumps(cfg))
    t = _oracle_vs_engines(cfg, 200, 31, ekn_sigma_fix=True)
    print("ekn sigma_fix:", t)
    assert t["exact"]["same"] == 1.0 and t["exact"]["same_a"] == 1.0
    for k in ("gV", "gG", "gA"):
        assert t["exact"][k] < 1e-7, (k, t["exact"][k])
        if t["tensor"]["all_same"]:
            assert t["tensor"][k] < GTOL_TENSOR, (k, t["tensor"][k])
    assert t["tensor"]["same"] > 0.98
    # literal sqrt(2): different trajectories
    e = cfg["eqn_config"]
    N, T = e["num_time_interval_critic"], e["total_time_critic"]
    outs = []
    for fix in (False, True):
        eng = Engine(e, cfg["net_config"], cfg["train_config"], dtype="float64", ekn_sigma_fix=fix)
        r = eng.critic_step(None, None, None, eng.tensor(z["x0"]), eng.tensor(z["dw"]), None, N, T, cheat_control=True, propagate_only=True, want=("x_smp",))
        outs.append(_npy(r["x_smp"]))
    assert np.abs(outs[0] - outs[1]).max() > 1e-2
    # the Equation class carries the switch (eqn_config.sigma_fix / constructor argument) into its engine
    from deeppde_actorcritic_b200 import equation, munchify
    assert equation.ekn(munchify(dict(e, sigma_fix=True))).sigma_fix and equation.EKN(munchify(e), sigma_fix=True).sigma_fix
    assert not equation.ekn(munchify(e)).sigma_fix


def test_vdp_d20_instantiation_vs_oracle():
    """configs/vdp_d20.json: d=20, control_dim=10 -- the <24, VDP, 10> instantiation of the tensor kernels (static cyclic
    neighbour indices), at the config's network sizes, against the oracle"""
    cfg = json.load(open(os.path.join(ROOT, "configs", "vdp_d20.json")))
    assert cfg["eqn_config"]["control_dim"] == 10 and cfg["eqn_config"]["dim"] == 20
    cfg["eqn_config"].update(num_time_interval_critic=20, num_time_interval_actor=20)
    t = _oracle_vs_engines(cfg, 300, 41)
    print("vdp_d20:", t)
    assert t["exact"]["same"] == 1.0 and t["tensor"]["same"] > 0.98
    for k in ("gV", "gG", "gA"):
        assert t["exact"][k] < 1e-7, (k, t["exact"][k])
        if t["tensor"]["all_same"]:
            assert t["tensor"][k] < GTOL_TENSOR, (k, t["tensor"][k])


@pytest.mark.parametrize("L", [5, 6])
def test_deep_networks_vs_oracle(L):
    """five and six hidden layers (the maximum the ABI accepts): the chunk schedule of the critic's NN_value phase holds
    7 (L+1) products (the round-1 schedule overflowed at L >= 5)"""
    z, cfg = _load("lqr_d5_adaptive_normal_td1")
    cfg = json.loads(json.dumps(cfg))
    cfg["net_config"]["num_hiddens_actor"] = [24, 40, 24, 16, 40, 24][:L]
    cfg["net_config"]["num_hiddens_critic"] = [40, 24, 24, 40, 16, 24][:L]
    t = _oracle_vs_engines(cfg, 200, 51 + L)
    print(f"L={L}:", t)
    assert t["exact"]["same"] == 1.0 and t["tensor"]["same"] > 0.98
    for k in ("gV", "gG", "gA"):
        assert t["exact"][k] < 1e-7, (k, t["exact"][k])
        if t["tensor"]["all_same"]:
            assert t["tensor"][k] < GTOL_TENSOR, (k, t["tensor"][k])


def test_tensor_checkpoint_resume(tmp_path):
    """train 6 iterations == train 3, checkpoint, reload into a fresh solver, train 3 on the TENSOR path: the device sampler
    is keyed by (seed, iteration), so the resumed run sees the same paths; the weights agree up to the FP32 summation order
    of the shared gradient slabs -- stated tolerance 2e-5 of each vector's max-norm.  The checkpoint also
    carries the history rows, the elapsed time and the NumPy RNG state."""
    from deeppde_actorcritic_b200 import equation, munchify
    from deeppde_actorcritic_b200.solver import ActorCriticSolver
    z, cfg = _load("lqr_d5_adaptive_normal_td1")
    cfg = json.loads(json.dumps(cfg))
    cfg["net_config"]["batch_size"] = 300

    def make():
        config = munchify(cfg)
        bsde = getattr(equation, config.eqn_config.eqn_name)(config.eqn_config)
        return ActorCriticSolver(config, bsde, compute_dtype="float32", seed=3, impl="tensor")

    a = make()
    for _ in range(6):
        a.train_iteration()
    b = make()
    for _ in range(3):
        b.train_iteration()
    b._history, b._elapsed = [[0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0]], 12.5
    np.random.seed(77)
    expect_next = np.random.RandomState(77).standard_normal(3)
    path = str(tmp_path / "ck.pt")
    b.save_checkpoint(path)
    assert not os.path.exists(path + ".tmp")
    np.random.seed(1)                                     # the resumed process starts from some other RNG state
    c = make()
    c.load_checkpoint(path)
    assert c._iter == 3 and c._history == b._history and c._elapsed == 12.5
    np.testing.assert_array_equal(np.random.standard_normal(3), expect_next)
    for _ in range(3):
        c.train_iteration()
    for ta, tc in ((a.model_actor.NN_control.theta, c.model_actor.NN_control.theta),
                   (a.model_critic.NN_value.theta, c.model_critic.NN_value.theta),
                   (a.model_critic.NN_value_grad.theta, c.model_critic.NN_value_grad.theta)):
        assert _gerr(_npy(tc), _npy(ta)) < 2e-5


@pytest.mark.parametrize("impl", ["exact", "tensor"])
def test_cuda_graph_iteration_matches_eager(impl):
    """a whole device-sampled iteration captured as ONE CUDA graph and replayed (the iteration counter and the
    Adam rates live in device memory) gives the weights of the kernel-by-kernel loop: bit for bit on the exact path, up to
    the FP32 slab-reduction order on the tensor path"""
    from deeppde_actorcritic_b200 import equation, munchify
    from deeppde_actorcritic_b200.solver import ActorCriticSolver
    z, cfg = _load("lqr_var_d8_adaptive_normal_td1")
    cfg = json.loads(json.dumps(cfg))
    cfg["net_config"]["batch_size"] = 300
    cfg["net_config"]["lr_boundaries_critic"], cfg["net_config"]["lr_values_critic"] = [2], [1e-3, 1e-4]     # the rate changes inside the run

    def make():
        config = munchify(cfg)
        bsde = getattr(equation, config.eqn_config.eqn_name)(config.eqn_config)
        return ActorCriticSolver(config, bsde, compute_dtype="float32", seed=5, impl=impl)

    a, b = make(), make()
    assert b.enable_cuda_graph()
    for _ in range(6):
        a.train_iteration()
        b.train_iteration()
    assert b._graph["graph"] is not None and b._graph["replays"] == 5 and b._graph["launches"] >= 10
    assert a.optimizer_critic.iterations == b.optimizer_critic.iterations == 6
    for ta, tb in ((a.model_actor.NN_control.theta, b.model_actor.NN_control.theta),
                   (a.model_critic.NN_value.theta, b.model_critic.NN_value.theta),
                   (a.model_critic.NN_value_grad.theta, b.model_critic.NN_value_grad.theta)):
        if impl == "exact":
            assert torch.equal(ta, tb)
        else:
            assert _gerr(_npy(tb), _npy(ta)) < 2e-5


def test_naive_lifetime_sort_is_a_relabelling():
    """naive scheme, tensor path: tiling the paths in order of their lifetime (forward-only pre-pass + counting sort, so that
    the tiles die as a whole) changes nothing per path -- same x0, same Philox increments, outputs at the path's own index:
    TD residuals, costs and exit indices are bit-identical to the unsorted run; gradients agree up to FP32 summation order"""
    cfg = json.load(open(os.path.join(ROOT, "configs", "bench_lqr_d5_naive_normal_td1.json")))
    e, net, tr = cfg["eqn_config"], cfg["net_config"], cfg["train_config"]
    assert tr["scheme"] == "naive"
    from oracle import ref_solver as RS
    rng = np.random.RandomState(8)
    # (the sort is taken only when there are more tiles than SMs: 200 tiles)
    B, N, T = 25500, int(e["num_time_interval_critic"]), float(e["total_time_critic"])
    res = []
    for sort in (False, True):
        eng = Engine(e, net, tr, dtype="float32", impl="tensor", lifetime_sort=sort)
        if not res:
            th = {}
            for k in ("actor", "critic", "critic_grad"):
                i, h, o, _ = RS.net_dims(cfg, k)
                th[k] = RS.init_params(i, h, o, rng)
        thd = {k: eng.tensor(v) for k, v in th.items()}
        x0, xb = eng.sample_x(3, 1, 0, B)
        kw = dict(dw_mode=1, seed=3, stream_id=7, need_grad=True)
        r = eng.critic_step(thd["actor"], thd["critic"], thd["critic_grad"], x0, None, xb, N, T, want=("delta", "delta_bdry", "exit_index", "coef"), **kw)
        a = eng.actor_step(thd["actor"], thd["critic"], x0, None, N, T, want=("delta", "exit_index"), **kw)
        torch.cuda.synchronize()
        res.append((r, a, eng.launch_count()))
    (r0, a0, l0), (r1, a1, l1) = res
    assert l1 == l0 + 8                                   # two pre-passes: forward-only rollout + histogram + scan + scatter
    for k in ("delta", "delta_bdry", "exit_index", "coef"):
        assert torch.equal(r0[k], r1[k]), k
    for k in ("delta", "exit_index"):
        assert torch.equal(a0[k], a1[k]), k
    live = float(r0["coef"].mean())
    assert live < 0.5                                     # the naive scheme really kills most path-steps here
    for g in ("grad_V", "grad_G"):
        assert _gerr(_npy(r1[g]), _npy(r0[g])) < 1e-5
    assert _gerr(_npy(a1["grad_actor"]), _npy(a0["grad_actor"])) < 1e-5
    np.testing.assert_allclose(_npy(r1["loss"]), _npy(r0["loss"]), rtol=1e-5)


def test_event_trace_fails_loudly_without_the_stats_build():
    """dpb_tc_trace: a library built without DPB_TC_STATS holds no trace and says so instead of returning stale zeros; a
    stats build returns a monotone event log for every traced role (diagnostic entry point, include/deeppde_b200.h)"""
    from deeppde_actorcritic_b200._cabi import DpbError
    cfg = json.load(open(os.path.join(ROOT, "configs", "lqr_d5.json")))
    e, net, tr = cfg["eqn_config"], cfg["net_config"], cfg["train_config"]
    from oracle import ref_solver as RS
    eng = Engine(e, net, tr, dtype="float32", impl="tensor")
    rng = np.random.RandomState(2)
    th = {}
    for k in ("actor", "critic", "critic_grad"):
        i, h, o, _ = RS.net_dims(cfg, k)
        th[k] = eng.tensor(RS.init_params(i, h, o, rng))
    B, N, T = 256, int(e["num_time_interval_critic"]), float(e["total_time_critic"])
    x0, xb = eng.sample_x(3, 1, 0, B)
    eng.critic_step(th["actor"], th["critic"], th["critic_grad"], x0, None, xb, N, T, dw_mode=1, seed=3, stream_id=7, need_grad=True)
    torch.cuda.synchronize()
    try:
        trace = eng.tc_trace(B, N)
    except DpbError as err:
        assert "DPB_TC_STATS" in str(err)
        return
    for r in (0, 2):                                       # owner thread 0 and the control warp always log
        w = trace[r][trace[r] != 0]
        assert len(w) > 10
        t = (w >> np.uint64(8)).astype(np.int64)
        assert (np.diff(t[: min(len(t), 200)]) >= 0).all()
