"""The C-ABI library loads and exports every symbol include/deeppde_b200.h declares; entry points
that need a GPU fail loudly (no CPU fallback).  CPU only: no compute call succeeds here."""
import ctypes as C
import os
import re

import pytest

from deeppde_actorcritic_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _cabi.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "deeppde_b200.h")).read()
    declared = set(re.findall(r"\b(dpb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert b"deeppde_b200" in lib.dpb_version()


def _cfg():
    c = _cabi.dpb_config()
    c.dtype, c.eqn, c.dim, c.control_dim, c.scheme, c.td_type = 0, 0, 5, 5, 1, 1
    c.n_hidden_actor = c.n_hidden_critic = 2
    for i in range(2):
        c.hidden_actor[i] = c.hidden_critic[i] = 200
    c.R, c.discount, c.p, c.q, c.beta = 1.0, 1.0, 1.0, 1.0, 1.0
    return c


def test_create_and_layout_queries(lib):
    c = _cfg()
    h = C.c_void_p()
    assert lib.dpb_create(C.byref(h), C.byref(c)) == 0
    # parameter counts of SURVEY 8a row 11 (d=5, 2x200): actor 42,825 / V 42,013 / G 42,825
    assert lib.dpb_param_count(h, 0) == 42825 and lib.dpb_param_count(h, 1) == 42013 and lib.dpb_param_count(h, 2) == 42825
    assert lib.dpb_workspace_bytes(h, 1024, 50) > 0
    assert lib.dpb_staging_bytes(h, 1024, 50, 0) > lib.dpb_staging_bytes(h, 1024, 50, 1) > 0
    assert lib.dpb_launch_count(h) == 0
    lib.dpb_destroy(h)
    c.dim = 40
    assert lib.dpb_create(C.byref(h), C.byref(c)) == _cabi.DPB_ERR_ARG
    assert b"dim" in lib.dpb_last_error(None)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    c = _cfg()
    h = C.c_void_p()
    assert lib.dpb_create(C.byref(h), C.byref(c)) == 0
    buf = (C.c_float * 64)()
    inp = _cabi.dpb_inputs()
    inp.x0 = C.addressof(buf)
    inp.x_bdry = C.addressof(buf)
    inp.dw_mode = 1
    ws = (C.c_char * 16)()
    rc = lib.dpb_critic_step(h, C.addressof(buf), C.addressof(buf), C.addressof(buf), C.byref(inp), 1, 0, 1, 10, 0.2, 0,
                             None, None, None, None, C.addressof(ws), 1 << 40, None)
    assert rc == _cabi.DPB_ERR_CUDA and b"no CPU fallback" in lib.dpb_last_error(h)
    assert lib.dpb_adam_step(h, C.addressof(buf), C.addressof(buf), C.addressof(buf), C.addressof(buf), 8, 1e-3, None, .9, .999, 1e-8, None) == _cabi.DPB_ERR_CUDA
    lib.dpb_destroy(h)
    from deeppde_actorcritic_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine({"eqn_name": "LQR", "dim": 5, "control_dim": 5, "R": 1.0, "discount": 1.0, "p": 1, "q": 1, "beta": 1},
               {"num_hiddens_actor": [8], "num_hiddens_critic": [8]}, {"scheme": "naive", "TD_type": "TD1"})
