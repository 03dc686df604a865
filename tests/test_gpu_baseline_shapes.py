"""The benchmarked tensor path (impl="tensor": tcgen05 MLP layers) against the float64 ORACLE at the five BASELINE.json
shapes -- the configs' own dimension, step count and network sizes, B = 1024 paths:

    lqr_d5       LQR     d=5   N=50   nets 2x200   naive    / normal  / TD1
    vdp_d10      VDP     d=10  N=100  nets 3x200   adaptive / bounded / TD2
    ekn_d20      ekn     d=20  N=100  nets 3x200   adaptive / normal  / TD1
    lqr_var_d20  LQR_var d=20  N=100  nets 3x200   adaptive / normal  / TD1
    lqr_d20      LQR     d=20  N=100  nets 3x200   adaptive / normal  / TD1   (the 2^20-path scaling config)

The kernel generates its Brownian increments in-kernel (Philox); the oracle is fed the same increments materialised by
dpb_philox_dw (test_philox_increments in test_gpu_parity.py proves the two are the same bits).

Stated tolerances of the tensor path (FP32 arithmetic, bf16x3 products with FP32 accumulation):
  * exit pattern: identical to the oracle's except for paths whose decisive proposal lies within EXIT_MARGIN of the
    boundary (|R - |p|| < 2e-5: FP32 rounding of the state against the float64 oracle); at most 3 such paths per 1024;
  * step sizes: rtol 2e-5 + atol 5e-9 (in the boundary layer dt = (R-|x|)^2 / (3 d sigma^2) amplifies the FP32 rounding
    of |x| ~ 1; the smallest step is 1e-4 T/N = 2e-7);
  * values (delta, delta_bdry, y): 3e-4 (rtol and atol) on the paths with the same exit pattern;
  * losses: 1e-3 relative; gradients: 2e-3 of the gradient's max-norm (the FP32 exact path's tolerance).  When a path
    exits at a different step than in the oracle its O(1/B) contribution differs, so gradients and losses are
    compared on a seed for which the exit patterns agree on every path (the seeds below are fixed).

Under cheat_control the rollout involves no network, and the tensor path must reproduce the exact FP32 path's
schedule bit for bit: coef, dt, exit index and every state (north_star: "exit-step indices and the adaptive step
schedule must be bit-exact").
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from deeppde_actorcritic_b200.engine import Engine
from oracle import ref_equation as RE
from oracle import ref_solver as RS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = {
    "lqr_d5": "bench_lqr_d5_naive_normal_td1.json",
    "vdp_d10": "bench_vdp_d10_adaptive_bounded_td2.json",
    "ekn_d20": "bench_ekn_d20_adaptive_normal_td1.json",
    "lqr_var_d20": "bench_lqr_var_d20_adaptive_normal_td1.json",
    "lqr_d20": "lqr_d20.json",
}
B = 1024
VTOL = dict(rtol=3e-4, atol=3e-4)
GTOL = 2e-3
EXIT_MARGIN = 2e-5
SEEDS = (11, 12, 13, 14)


def _npy(t):
    return t.detach().cpu().double().numpy()


def _gerr(got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    scale = np.abs(ref).max()
    return float(np.abs(got).max()) if scale == 0 else float(np.abs(got - ref).max() / scale)


def _setup(key, seed):
    cfg = json.load(open(os.path.join(ROOT, "configs", SHAPES[key])))
    e, net, tr = cfg["eqn_config"], cfg["net_config"], cfg["train_config"]
    eng = Engine(e, net, tr, dtype="float32", impl="tensor")
    rng = np.random.RandomState(100 + seed)
    th = {}
    for k in ("actor", "critic", "critic_grad"):
        i, h, o, _ = RS.net_dims(cfg, k)
        p = RS.init_params(i, h, o, rng)
        p[-3 * o:-2 * o] = rng.normal(0, 0.1, o)          # non-zero last bias (the reference initialises it to zero)
        th[k] = p.astype(np.float32).astype(np.float64)   # both sides start from the same float32-representable weights
    N, T = int(e["num_time_interval_critic"]), float(e["total_time_critic"])
    dw_mode = 2 if tr["sample_type"] == "bounded" else 1
    x0, xb = eng.sample_x(seed, 1, 0, B)
    dw = eng.philox_dw(dw_mode, seed, 3, 0, B, N)
    return cfg, eng, th, x0, xb, dw, N, T, dw_mode


def _margins(coef_k, coef_o, x_k, x_o, R):
    """for every path whose exit pattern differs: |R - |p|| of the decisive proposal (the accepted side's next state)"""
    out = []
    for b in np.nonzero((coef_k != coef_o).any(1))[0]:
        t = int(np.nonzero(coef_k[b] != coef_o[b])[0][0])
        p = x_k[b, :, t + 1] if coef_k[b, t] > 0 else x_o[b, :, t + 1]
        out.append(abs(R - float(np.sqrt((p ** 2).sum()))))
    return out


@pytest.mark.parametrize("key", list(SHAPES))
def test_tensor_vs_oracle_baseline_shape(key):
    report = None
    for seed in SEEDS:
        cfg, eng, th, x0, xb, dw, N, T, dw_mode = _setup(key, seed)
        eqn = RE.make_ref_equation(cfg["eqn_config"])
        R = float(cfg["eqn_config"]["R"])
        thd = {k: eng.tensor(v) for k, v in th.items()}
        tt = {k: torch.tensor(v) for k, v in th.items()}
        inputs = tuple(t.detach().cpu().double() for t in (x0, dw, xb))
        kw = dict(dw_mode=dw_mode, seed=seed, stream_id=3)
        # ---- critic: TD residuals, loss, gradients
        loss_c, gV, gG, delta, delta_b, aux = RS.grad_critic(eqn, cfg, tt, inputs, False)
        r = eng.critic_step(thd["actor"], thd["critic"], thd["critic_grad"], x0, None, xb, N, T, need_grad=True,
                            want=("delta", "delta_bdry", "coef", "dt", "x_smp"), **kw)
        ck, co = _npy(r["coef"]), aux["coef"].detach().numpy()
        same = (ck == co).all(1)
        marg = _margins(ck, co, _npy(r["x_smp"]), aux["x"].detach().numpy(), R)
        assert len(marg) <= 3 and all(m < EXIT_MARGIN for m in marg), (key, seed, "critic rollout exit mismatches", marg)
        np.testing.assert_allclose(_npy(r["dt"])[same], aux["dt"].detach().numpy()[same], rtol=2e-5, atol=5e-9)
        np.testing.assert_allclose(_npy(r["delta"])[same], delta.numpy()[same], **VTOL)
        np.testing.assert_allclose(_npy(r["delta_bdry"]), delta_b.numpy(), **VTOL)
        # ---- actor: cost, loss, gradient through the whole trajectory
        loss_a, gA, y, aux_a = RS.grad_actor(eqn, cfg, tt, inputs, False, False)
        a = eng.actor_step(thd["actor"], thd["critic"], x0, None, N, T, need_grad=True, want=("delta", "coef", "x_smp"), **kw)
        ak, ao = _npy(a["coef"]), aux_a["coef"].detach().numpy()
        same_a = (ak == ao).all(1)
        marg_a = _margins(ak, ao, _npy(a["x_smp"]), aux_a["x"].detach().numpy(), R)
        assert len(marg_a) <= 3 and all(m < EXIT_MARGIN for m in marg_a), (key, seed, "actor rollout exit mismatches", marg_a)
        np.testing.assert_allclose(_npy(a["delta"])[same_a], y.numpy()[same_a], **VTOL)
        report = dict(seed=seed, mismatch_critic=int((~same).sum()), mismatch_actor=int((~same_a).sum()), margins=marg + marg_a)
        if same.all() and same_a.all():
            errs = {"loss_c": abs(float(r["loss"].sum()) - float(loss_c)) / max(1.0, abs(float(loss_c))),
                    "loss_a": abs(float(a["loss"][0]) - float(loss_a)) / max(1.0, abs(float(loss_a))),
                    "gV": _gerr(_npy(r["grad_V"]), gV.numpy()), "gG": _gerr(_npy(r["grad_G"]), gG.numpy()),
                    "gA": _gerr(_npy(a["grad_actor"]), gA.numpy())}
            print(key, report, {k: f"{v:.2e}" for k, v in errs.items()},
                  "max |d delta| = %.2e" % np.abs(_npy(r["delta"]) - delta.numpy()).max())
            assert errs["loss_c"] < 1e-3 and errs["loss_a"] < 1e-3, errs
            for k in ("gV", "gG", "gA"):
                assert errs[k] < GTOL, (key, k, errs[k])
            return
        print(key, "exit pattern differs on this seed (within the margin), trying the next:", report)
    pytest.fail(f"{key}: no seed in {SEEDS} with an identical exit pattern: {report}")


@pytest.mark.parametrize("key", list(SHAPES))
def test_tensor_schedule_bit_exact_under_cheat_control(key):
    """no network in the rollout => the tensor kernels must reproduce the exact FP32 path bit for bit"""
    cfg, tn, th, x0, xb, dw, N, T, dw_mode = _setup(key, 21)
    ex = Engine(cfg["eqn_config"], cfg["net_config"], cfg["train_config"], dtype="float32", impl="exact")
    kw = dict(dw_mode=dw_mode, seed=21, stream_id=3, cheat_control=True)
    want = ("coef", "dt", "exit_index", "x_smp")
    thd = {k: tn.tensor(v) for k, v in th.items()}
    a = ex.critic_step(None, thd["critic"], thd["critic_grad"], x0, None, xb, N, T, want=want + ("delta",), **kw)
    b = tn.critic_step(None, thd["critic"], thd["critic_grad"], x0, None, xb, N, T, want=want + ("delta",), **kw)
    for k in want:
        assert torch.equal(a[k], b[k]), f"{key}: critic rollout {k} differs between impl=tensor and impl=exact under cheat_control"
    np.testing.assert_allclose(_npy(b["delta"]), _npy(a["delta"]), **VTOL)
    # the same with externally supplied increments, and for the propagate-only entry (Equation.propagate_*)
    kw2 = dict(cheat_control=True, propagate_only=True)
    a = ex.critic_step(None, None, None, x0, dw, None, N, T, want=want, **kw2)
    b = tn.critic_step(None, None, None, x0, dw, None, N, T, want=want, **kw2)
    for k in want:
        assert torch.equal(a[k], b[k]), f"{key}: propagate {k} differs between impl=tensor and impl=exact under cheat_control"
    # ... and both equal the INDEPENDENT float32 restatement of the reference's scheme (oracle/ref_schedule_f32.py, NumPy),
    # bit for bit: states, step sizes, coef, exit index
    from oracle.ref_schedule_f32 import ScheduleF32
    o = ScheduleF32(cfg["eqn_config"], cfg["train_config"]["scheme"], T, N)
    xs_o, dt_o, cf_o, ex_o = o.propagate(x0.cpu().numpy(), dw.cpu().numpy())
    assert np.array_equal(b["coef"].cpu().numpy(), cf_o) and np.array_equal(b["exit_index"].cpu().numpy(), ex_o)
    assert np.array_equal(b["dt"].cpu().numpy().view(np.uint32), dt_o.view(np.uint32)), f"{key}: dt bits differ from the float32 oracle"
    assert np.array_equal(b["x_smp"].cpu().numpy().view(np.uint32), xs_o.view(np.uint32)), f"{key}: state bits differ from the float32 oracle"
    ya = ex.actor_step(None, thd["critic"], x0, None, N, T, want=want + ("delta",), **kw)
    yb = tn.actor_step(None, thd["critic"], x0, None, N, T, want=want + ("delta",), **kw)
    for k in want:
        assert torch.equal(ya[k], yb[k]), f"{key}: actor rollout {k} differs between impl=tensor and impl=exact under cheat_control"
    np.testing.assert_allclose(_npy(yb["delta"]), _npy(ya["delta"]), **VTOL)
    live = float(a["coef"].mean())
    print(key, f"bit-identical schedule on {B} paths x {N} steps (live fraction {live:.3f})")
