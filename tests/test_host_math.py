"""Per-path arithmetic of csrc/dpb_eqn.h (the code the CUDA kernels run per path) compiled with g++
and checked against the oracle: closed forms, step schedule (bit-exact in float64), actor cost and
the reverse recursion (vs torch autograd through the oracle's rollout).  CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from deeppde_actorcritic_b200 import _cabi
from oracle import ref_equation as RE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "harness", "host_harness.cpp")
SO = os.path.join(ROOT, "tests", "harness", "_host_harness.so")

EQNS = {
    "LQR": {"eqn_name": "LQR", "discount": 1.0, "p": 1.3, "q": 0.7, "beta": 0.9, "R": 1.0},
    "VDP": {"eqn_name": "VDP", "discount": 1.0, "a": 1.0, "epsilon": 0.1, "q": 1.2, "R": 1.0},
    "ekn": {"eqn_name": "ekn", "discount": 0.0, "a2": 1.2, "a3": 0.2, "R": 1.0},
    "LQR_var": {"eqn_name": "LQR_var", "discount": 1.0, "q": 1.1, "beta": 0.8, "epsilon": 0.05, "R": 1.0},
}
DIMS = {"LQR": (5, 5), "VDP": (6, 3), "ekn": (4, 4), "LQR_var": (6, 6)}


@pytest.fixture(scope="module")
def hh():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(
            os.path.getmtime(SRC), os.path.getmtime(os.path.join(ROOT, "deeppde_actorcritic_b200", "csrc", "dpb_eqn.h"))):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", SO, SRC])
    return C.CDLL(SO)


def make_cfg(name, scheme, dim=None, m=None):
    e = dict(EQNS[name])
    d0, m0 = DIMS[name]
    e["dim"], e["control_dim"] = dim or d0, m or m0
    c = _cabi.dpb_config()
    c.dtype = 1
    c.eqn = _cabi.EQN_IDS[name]
    c.dim, c.control_dim = e["dim"], e["control_dim"]
    c.scheme = _cabi.SCHEME_IDS[scheme]
    c.td_type = 1
    c.R, c.discount = e["R"], e["discount"]
    for k in ("p", "q", "beta", "a", "epsilon", "a2", "a3"):
        setattr(c, k, e.get(k, 0.0))
    return e, c


def dptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("name", list(EQNS))
def test_closed_forms(hh, name):
    e, c = make_cfg(name, "naive")
    eqn = RE.make_ref_equation(e)
    rng = np.random.RandomState(3)
    d, m = e["dim"], e["control_dim"]
    for _ in range(5):
        x = rng.normal(0, 0.4, d)
        u = rng.normal(0, 0.5, m)
        out = np.zeros(3 + m + d)
        hh.hh_closed_forms_f64(C.byref(c), dptr(x), dptr(u), dptr(out))
        xt, ut = torch.tensor(x)[None], torch.tensor(u)[None]
        np.testing.assert_allclose(out[0], float(eqn.V_true(xt)), rtol=1e-13)
        np.testing.assert_allclose(out[1], float(eqn.Z(xt)), rtol=1e-13)
        np.testing.assert_allclose(out[2], float(eqn.w(xt, ut)), rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(out[3:3 + m], eqn.u_true(xt)[0].numpy(), rtol=1e-13)
        np.testing.assert_allclose(out[3 + m:], eqn.V_grad_true(xt)[0].numpy(), rtol=1e-13)


@pytest.mark.parametrize("name", list(EQNS))
@pytest.mark.parametrize("scheme,T", [("naive", 0.3), ("adaptive", 1.5), ("adaptive", 0.12)])
@pytest.mark.parametrize("cheat", [0, 1])
def test_rollout_and_adjoint(hh, name, scheme, T, cheat):
    e, c = make_cfg(name, scheme)
    eqn = RE.make_ref_equation(e)
    d, m = e["dim"], e["control_dim"]
    N, B = 12, 16
    rng = np.random.RandomState(5)
    A = rng.normal(0, 0.4, (m, d))
    b = rng.normal(0, 0.1, m)
    np.random.seed(17)
    x0, dw, _ = eqn.sample_normal(B, N)
    if scheme == "naive":
        x0 *= 0.9
    elif T < 1:
        x0 *= 0.5                                      # mix of inner and boundary-layer steps
    At = torch.tensor(A, requires_grad=True)
    bt = torch.tensor(b, requires_grad=True)
    control = (lambda x: eqn.u_true(x)) if cheat else (lambda x: x @ At.T + bt)
    x0t, dwt = torch.tensor(x0), torch.tensor(dw)
    if scheme == "naive":
        xs, dts, coefs = eqn.propagate_naive(x0t, dwt, control, T, N)
    else:
        xs, dts, coefs = eqn.propagate_adaptive(x0t, dwt, control, T, N)
    y = 0
    disc = torch.ones(B, 1, dtype=torch.float64)
    for t in range(N):
        xt = xs[:, :, t]
        w = eqn.w(xt, control(xt))
        y = y + coefs[:, t:t + 1] * w * dts[:, t:t + 1] * disc
        disc = disc * torch.exp(-eqn.gamma * dts[:, t:t + 1] * coefs[:, t:t + 1])
    y = y + eqn.V_true(xs[:, :, -1]) * disc
    loss = y.mean()
    if not cheat:
        gA_ref, gb_ref = torch.autograd.grad(loss, [At, bt])
    assert 0.05 < float(coefs.detach().mean()) <= 1.0
    if scheme == "adaptive":                         # boundary-layer steps and exits both occur
        assert float((dts.detach() != T / N).double().mean()) > 0.2
        if T < 1:
            assert float((dts.detach() == T / N).double().mean()) > 0.02

    gA_sum, gb_sum = np.zeros((m, d)), np.zeros(m)
    for i in range(B):
        xs_o, dt_o, cf_o = np.zeros((d, N + 1)), np.zeros(N), np.zeros(N)
        y_o = np.zeros(1)
        gA, gb = np.zeros((m, d)), np.zeros(m)
        hh.hh_run_path_f64(C.byref(c), N, C.c_double(T), dptr(A), dptr(b), dptr(np.ascontiguousarray(x0[i])),
                           dptr(np.ascontiguousarray(dw[i])), cheat, C.c_double(1.0 / B), dptr(xs_o), dptr(dt_o),
                           dptr(cf_o), dptr(y_o), None if cheat else dptr(gA), None if cheat else dptr(gb))
        assert np.array_equal(cf_o, coefs[i].detach().numpy())                          # exit pattern: exact
        np.testing.assert_allclose(dt_o, dts[i].detach().numpy(), rtol=1e-13, atol=0)
        np.testing.assert_allclose(xs_o, xs[i].detach().numpy(), rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(y_o[0], float(y[i].detach()), rtol=1e-11, atol=1e-13)
        gA_sum += gA
        gb_sum += gb
    if not cheat:
        np.testing.assert_allclose(gA_sum, gA_ref.numpy(), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(gb_sum, gb_ref.numpy(), rtol=1e-9, atol=1e-12)


def test_ekn_head(hh):
    rng = np.random.RandomState(2)
    m = 5
    for sign in (1.0, -1.0):
        y = rng.normal(0, 1, m + 1)
        y[m] = sign * abs(y[m])
        ubar = rng.normal(0, 1, m)
        u, ybar = np.zeros(m), np.zeros(m + 1)
        hh.hh_ekn_head_f64(dptr(y), dptr(ubar), m, dptr(u), dptr(ybar))
        yt = torch.tensor(y, requires_grad=True)
        ut = yt[:m] / (1e-15 + torch.relu(yt[m]) + torch.sum(yt[:m] ** 2) ** 0.5)
        (g,) = torch.autograd.grad(torch.sum(ut * torch.tensor(ubar)), [yt])
        np.testing.assert_allclose(u, ut.detach().numpy(), rtol=1e-13)
        np.testing.assert_allclose(ybar, g.numpy(), rtol=1e-11, atol=1e-14)


# ------------------------------------------------------------------------------------------------
# FLOAT32: the header (g++ build, the very code the CUDA kernels run per path) against the independent NumPy float32
# restatement of the reference's schemes (oracle/ref_schedule_f32.py) -- bit for bit
@pytest.mark.parametrize("name", list(EQNS))
@pytest.mark.parametrize("scheme,T", [("naive", 0.3), ("adaptive", 1.5), ("adaptive", 0.12)])
def test_f32_schedule_header_vs_numpy_oracle(hh, name, scheme, T):
    from oracle.ref_schedule_f32 import ScheduleF32
    N, B = 30, 200
    d, m = (20, 20) if name in ("LQR", "LQR_var", "ekn") else (20, 10)          # the BASELINE dimension
    e, c = make_cfg(name, scheme, d, m)
    c.dtype = 0
    rng = np.random.RandomState(17)
    g = rng.standard_normal((B, d))
    sc = 0.6 if scheme == "naive" else 1.0          # (naive: paths that start near the boundary at d=20 leave at once)
    x0 = (sc * rng.uniform(0, 1, (B, 1)) ** (1.0 / d) * g / np.sqrt((g ** 2).sum(1, keepdims=True))).astype(np.float32)
    dw = rng.standard_normal((B, d, N)).astype(np.float32)
    o = ScheduleF32(e, scheme, T, N)
    xs_o, dt_o, cf_o, ex_o = o.propagate(x0, dw)
    A, b = np.zeros((m, d), np.float32), np.zeros(m, np.float32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    for i in range(B):
        xs, dts, cfs, y = np.zeros((d, N + 1), np.float32), np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros(1, np.float32)
        hh.hh_run_path_f32(C.byref(c), N, C.c_double(T), P(A), P(b), P(np.ascontiguousarray(x0[i])), P(np.ascontiguousarray(dw[i])),
                           1, C.c_float(1.0), P(xs), P(dts), P(cfs), P(y), None, None)
        assert np.array_equal(cfs, cf_o[i]), f"path {i}: coef differs"
        assert np.array_equal(dts.view(np.uint32), dt_o[i].view(np.uint32)), f"path {i}: dt bits differ"
        assert np.array_equal(xs.view(np.uint32), xs_o[i].view(np.uint32)), f"path {i}: x bits differ"
    assert 0.003 < cf_o.mean() < 0.9999 and (dt_o != dt_o[0, 0]).any() == (scheme == "adaptive")
    # and the float32 schedule shadows the float64 oracle's: same exit pattern except within rounding of the boundary
    eqn = RE.make_ref_equation(e)
    x64, dw64 = torch.tensor(x0.astype(np.float64)), torch.tensor(dw.astype(np.float64))
    prop = eqn.propagate_naive if scheme == "naive" else eqn.propagate_adaptive
    _, _, coef64 = prop(x64, dw64, lambda x: eqn.u_true(x), T, N)
    assert (coef64.numpy() == cf_o).all(1).mean() > 0.97
