import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from deeppde_actorcritic_b200 import _cabi
lib = _cabi.load()
out = np.zeros(2, dtype=np.int64)
for n in (16, 96, 112, 208):
    row = []
    for pc in (1, 2, 3, 6, 9, 12, 18, 39):
        rc = lib.dpb_tc_mma_cycles(out.ctypes.data_as(C.c_void_p), n, 4000 // pc + 1, 1, pc)
        row.append(f"{pc}/commit: {out[0]*pc}")
    print(f"TS N={n}: cycles per (group + commit): " + "  ".join(row))
