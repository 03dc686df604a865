"""Prints the cost of one tcgen05.mma (M=128, K=16, bf16) as a function of N (diagnostic; see dpb_tc_mma_cycles)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from deeppde_actorcritic_b200 import _cabi
lib = _cabi.load()
out = np.zeros(2, dtype=np.int64)
for ts in (1, 0):
    for per_commit in (3, 39):
        row = []
        for n in (16, 32, 48, 64, 96, 112, 128, 160, 208, 256):
            rc = lib.dpb_tc_mma_cycles(out.ctypes.data_as(C.c_void_p), n, 2000 // per_commit + 1, ts, per_commit)
            assert rc == 0, lib.dpb_last_error(None)
            row.append(f"N={n}: {out[0]} ({out[1]})")
        print(f"A in {'TMEM' if ts else 'smem'}, {per_commit} MMAs per commit: cycles per MMA total (issue): " + "  ".join(row))
