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
