"""ctypes binding of include/deeppde_b200.h (libdeeppde_b200.so).

This is the only place the package touches the C ABI.  There is no CPU fallback: if the shared
library is missing, or a compute entry point is called without a CUDA device, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

DPB_MAX_HIDDEN = 6
DPB_MAX_DIM = 32

DPB_OK, DPB_ERR_ARG, DPB_ERR_CUDA, DPB_ERR_WORKSPACE = 0, 1, 2, 3
EQN_IDS = {"LQR": 0, "VDP": 1, "ekn": 2, "EKN": 2, "LQR_var": 3}
SCHEME_IDS = {"naive": 0, "adaptive": 1}
TD_IDS = {"TD1": 1, "TD2": 2}
DTYPE_IDS = {"float32": 0, "float64": 1}
NET_ACTOR, NET_CRITIC, NET_CRITIC_GRAD = 0, 1, 2
DW_EXTERNAL, DW_PHILOX_NORMAL, DW_PHILOX_BOUNDED = 0, 1, 2
FLAG_CHEAT_CONTROL, FLAG_CHEAT_VALUE, FLAG_NEED_GRAD, FLAG_PROPAGATE_ONLY = 1, 2, 4, 8
CF_V_TRUE, CF_U_TRUE, CF_V_GRAD_TRUE, CF_Z, CF_W, CF_DRIFT, CF_SIGMA = 0, 1, 2, 3, 4, 5, 6
IMPL_EXACT, IMPL_TENSOR = 0, 1


class dpb_config(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("eqn", C.c_int32), ("dim", C.c_int32), ("control_dim", C.c_int32),
        ("scheme", C.c_int32), ("td_type", C.c_int32), ("ekn_sigma_fix", C.c_int32),
        ("n_hidden_actor", C.c_int32), ("n_hidden_critic", C.c_int32),
        ("hidden_actor", C.c_int32 * DPB_MAX_HIDDEN), ("hidden_critic", C.c_int32 * DPB_MAX_HIDDEN),
        ("impl", C.c_int32), ("reserved", C.c_int32 * 3),
        ("R", C.c_double), ("discount", C.c_double),
        ("p", C.c_double), ("q", C.c_double), ("beta", C.c_double),
        ("a", C.c_double), ("epsilon", C.c_double), ("a2", C.c_double), ("a3", C.c_double),
    ]


class dpb_inputs(C.Structure):
    _fields_ = [
        ("x0", C.c_void_p), ("dw", C.c_void_p), ("x_bdry", C.c_void_p),
        ("dw_mode", C.c_int32), ("reserved", C.c_int32),
        ("seed", C.c_uint64), ("stream", C.c_uint64), ("stream_base", C.c_void_p),
    ]


class dpb_path_outputs(C.Structure):
    _fields_ = [
        ("x_smp", C.c_void_p), ("dt", C.c_void_p), ("coef", C.c_void_p),
        ("delta", C.c_void_p), ("delta_bdry", C.c_void_p), ("exit_index", C.c_void_p),
    ]


_HERE = os.path.dirname(os.path.abspath(__file__))
# (DPB_LIB_PATH: an experiment build of the same library, e.g. another helper-group layout -- tools/variant_builds.sh)
LIB_PATH = os.environ.get("DPB_LIB_PATH") or os.path.join(_HERE, "libdeeppde_b200.so")

# name -> (restype, argtypes); every symbol include/deeppde_b200.h declares
_P, _I64, _I32, _U32, _U64, _D = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_uint64, C.c_double
SYMBOLS = {
    "dpb_create": (C.c_int, [C.POINTER(_P), C.POINTER(dpb_config)]),
    "dpb_destroy": (C.c_int, [_P]),
    "dpb_last_error": (C.c_char_p, [_P]),
    "dpb_version": (C.c_char_p, []),
    "dpb_param_count": (_I64, [_P, C.c_int]),
    "dpb_workspace_bytes": (_I64, [_P, _I64, _I32]),
    "dpb_staging_bytes": (_I64, [_P, _I64, _I32, _I32]),
    "dpb_critic_step": (C.c_int, [_P, _P, _P, _P, C.POINTER(dpb_inputs), _I64, _I64, _I64, _I32, _D, _U32,
                                  _P, _P, _P, C.POINTER(dpb_path_outputs), _P, _I64, _P]),
    "dpb_actor_step": (C.c_int, [_P, _P, _P, C.POINTER(dpb_inputs), _I64, _I64, _I64, _I32, _D, _U32,
                                 _P, _P, C.POINTER(dpb_path_outputs), _P, _I64, _P]),
    "dpb_mlp_forward": (C.c_int, [_P, C.c_int, _P, _P, _I64, _P, _P, _I64, _P]),
    "dpb_closed_form": (C.c_int, [_P, C.c_int, _P, _P, _I64, _P, _P]),
    "dpb_diffusion": (C.c_int, [_P, _P, _P, _P, _I64, _P, _P]),
    "dpb_err_metrics": (C.c_int, [_P, _P, _P, _I64, _P, _P]),
    "dpb_adam_step": (C.c_int, [_P, _P, _P, _P, _P, _I64, _D, _P, _D, _D, _D, _P]),
    "dpb_philox_dw": (C.c_int, [_P, _I32, _U64, _U64, _I64, _I64, _I32, _P, _P]),
    "dpb_sample_x": (C.c_int, [_P, _U64, _U64, _P, _I64, _I64, _P, _P, _P]),
    "dpb_critic_step_host": (C.c_int, [_P, _P, _P, _P, C.POINTER(dpb_inputs), _I64, _I64, _I64, _I32, _D, _U32,
                                       _P, _P, _P, _P, _I64, _P]),
    "dpb_actor_step_host": (C.c_int, [_P, _P, _P, C.POINTER(dpb_inputs), _I64, _I64, _I64, _I32, _D, _U32,
                                      _P, _P, _P, _I64, _P]),
    "dpb_launch_count": (_I64, [_P]),
    "dpb_set_timing": (C.c_int, [_P, C.c_int]),
    "dpb_last_kernel_ms": (_D, [_P]),
    "dpb_tc_selftest": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "dpb_tc_handshake_cycles": (C.c_int, [_P, C.c_int]),
    "dpb_tc_epilogue_cycles": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "dpb_tc_stats": (C.c_int, [_P, _P, _I64, _I32, _P]),
    "dpb_tc_trace": (C.c_int, [_P, _P, _I64, _I32, _P]),
    "dpb_tc_mma_cycles": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int]),
}

_lib = None


def load():
    """dlopen libdeeppde_b200.so (built in-tree by __graft_entry__.build()) and type every symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)        # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


class DpbError(RuntimeError):
    pass


def check(lib, handle, rc):
    if rc != 0:
        msg = lib.dpb_last_error(handle)
        raise DpbError(f"libdeeppde_b200 error {rc}: {msg.decode() if msg else '?'}")
