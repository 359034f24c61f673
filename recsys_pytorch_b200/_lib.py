"""ctypes binding of libb200rec.so (include/b200rec.h).

The library is the product; there is NO fallback: importing the engine without
the built shared object, or calling it without a CUDA device, raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200REC_LIB: profiling builds only (tools/build_ablate.sh); the product library is always the in-tree one
LIB_PATH = os.environ.get("B200REC_LIB") or os.path.join(_HERE, "libb200rec.so")

OK, EINVAL, ECUDA, ENOMEM, EUNSUPPORTED = 0, -1, -2, -3, -4
SINK_UPDATE, SINK_STAGE, SINK_GRAD, SINK_NONE = 0, 1, 2, 3
F_USERS_UNIQUE, F_TMA_GATHER, F_ITEM_DELTA, F_GENERIC, F_ASYNC_GATHER, F_ITEM_DELTA_BF16, F_L2_HINTS = 1, 2, 4, 8, 16, 32, 64
F_P2P_ROUND_ROBIN, F_P2P_NO_UWRITE, F_P2P_NO_UREAD, F_P2P_PURE_SEQUENTIAL = 256, 512, 1024, 2048
GATHER_FLAGS = {"ldg": 0, "tma": F_TMA_GATHER, "async": F_ASYNC_GATHER, "generic": F_GENERIC, "ldg_hints": F_L2_HINTS}
SCORE_EXACT, SCORE_TC = 0, 1


class B200RecError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libb200rec error {code}: {msg}")
        self.code = code


class BprArgs(C.Structure):
    """struct b200rec_bpr_args (include/b200rec.h)."""
    _fields_ = [
        ("U", C.c_void_p), ("V", C.c_void_p),
        ("ld", C.c_int32), ("d", C.c_int32),
        ("num_users", C.c_int32), ("num_items", C.c_int32),
        ("users", C.c_void_p), ("pos", C.c_void_p), ("neg", C.c_void_p),
        ("B", C.c_int32),
        ("csr_indptr", C.c_void_p), ("csr_indices", C.c_void_p),
        ("seed", C.c_uint64), ("step", C.c_uint64),
        ("out_pos", C.c_void_p), ("out_neg", C.c_void_p),
        ("lr", C.c_float), ("reg", C.c_float),
        ("sink", C.c_int32), ("flags", C.c_int32),
        ("stage", C.c_void_p), ("gU", C.c_void_p), ("gV", C.c_void_p),
        ("loss_sum", C.c_void_p), ("x_out", C.c_void_p),
        ("item_lo", C.c_int32), ("item_hi", C.c_int32),
        ("udelta", C.c_void_p), ("inv_batch", C.c_float),
    ]


MAX_RANKS, PEER_HANDLE_BYTES = 16, 64


class P2PRouteArgs(C.Structure):
    """struct b200rec_p2p_route_args (include/b200rec.h)."""
    _fields_ = [
        ("users", C.c_void_p), ("pos", C.c_void_p), ("neg", C.c_void_p), ("B", C.c_int32),
        ("csr_indptr", C.c_void_p), ("csr_indices", C.c_void_p),
        ("seed", C.c_uint64), ("step", C.c_uint64),
        ("world", C.c_int32), ("rank", C.c_int32),
        ("item_bounds", C.c_int32 * (MAX_RANKS + 1)), ("head", C.c_int32),
        ("out_u", C.c_void_p), ("out_i", C.c_void_p), ("out_j", C.c_void_p), ("out_cnt", C.c_void_p),
        ("cap", C.c_int32),
        ("dbg_pos", C.c_void_p), ("dbg_neg", C.c_void_p),
    ]


class P2PStepArgs(C.Structure):
    """struct b200rec_p2p_step_args (include/b200rec.h)."""
    _fields_ = [
        ("world", C.c_int32), ("rank", C.c_int32), ("ld", C.c_int32), ("d", C.c_int32),
        ("U_peer", C.c_void_p * MAX_RANKS), ("V_peer", C.c_void_p * MAX_RANKS),
        ("item_bounds", C.c_int32 * (MAX_RANKS + 1)), ("head", C.c_int32),
        ("Vh", C.c_void_p), ("dVh", C.c_void_p),
        ("in_u", C.c_void_p * MAX_RANKS), ("in_i", C.c_void_p * MAX_RANKS), ("in_j", C.c_void_p * MAX_RANKS),
        ("in_cnt", C.c_void_p * MAX_RANKS),
        ("lr", C.c_float), ("reg", C.c_float), ("inv_batch", C.c_float),
        ("flags", C.c_int32),
        ("loss_sum", C.c_void_p), ("n_processed", C.c_void_p),
    ]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
_PROTOS = {
    "b200rec_last_error": (C.c_char_p, []),
    "b200rec_version": (_I, []),
    "b200rec_launch_count": (_L, []),
    "b200rec_last_step_kernel": (C.c_char_p, []),
    "b200rec_top_k_array_index": (_I, [_P, _I, _I, _I, _P]),
    "b200rec_evaluate_holdout": (_I, [_I, _P, _I, _P, _I, _P, _P, _P]),
    "b200rec_evaluate_loo": (_I, [_I, _P, _I, _P, _I, _P, _P]),
    "b200rec_mf_forward": (_I, [_P, _P, _I, _I, _P, _P, _I, _P, _P]),
    "b200rec_bpr_step": (_I, [C.POINTER(BprArgs), _P]),
    "b200rec_pointwise_step": (_I, [_P, _P, _I, _I, _P, _P, _P, _I, _I, _F, _F, _I, _P, _P, _P, _F, _P]),
    "b200rec_sample_triples": (_I, [_P, _I, _P, _P, _I, C.c_uint64, C.c_uint64, _P, _P, _P]),
    "b200rec_bpr_apply": (_I, [_P, _P, _I, _P, _P, _P, _I, _P, _P]),
    "b200rec_rows_add": (_I, [_P, _I, _P, _I, _P, _I, _F, _P]),
    "b200rec_add_bf16": (_I, [_P, _P, _L, _P]),
    "b200rec_sgd_dense": (_I, [_P, _P, _L, _F, _P]),
    "b200rec_snap_apply": (_I, [_P, _P, _P, _F, _L, _P]),
    "b200rec_add_clear": (_I, [_P, _P, _L, _P]),
    "b200rec_delta_diff": (_I, [_P, _P, _P, _P, _L, _P]),
    "b200rec_delta_apply": (_I, [_P, _P, _P, _L, _P]),
    "b200rec_adam_dense": (_I, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _I, _P]),
    "b200rec_p2p_route": (_I, [C.POINTER(P2PRouteArgs), _P]),
    "b200rec_p2p_step": (_I, [C.POINTER(P2PStepArgs), _P]),
    "b200rec_peer_alloc": (_I, [_L, C.POINTER(C.c_void_p)]),
    "b200rec_peer_free": (_I, [_P]),
    "b200rec_peer_export": (_I, [_P, _P]),
    "b200rec_peer_import": (_I, [_P, C.POINTER(C.c_void_p)]),
    "b200rec_peer_close": (_I, [_P]),
    "b200rec_peer_copy": (_I, [_P, _P, _L, _P]),
    "b200rec_l2_persist": (_I, [_P, _L, _F, _P]),
    "b200rec_adam_rows": (_I, [_P, _P, _P, _P, _P, _I, _P, _I, _F, _F, _F, _F, _I, _P]),
    "b200rec_score_topk_workspace": (_L, [_I, _I, _I, _I, _I]),
    "b200rec_score_topk": (_I, [_P, _P, _I, _I, _P, _I, _I, _P, _P, _I, _P, _P, _P, _L, _I, _P]),
    "b200rec_debug_tc_layout": (_I, [_I, _I, _I, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "b200rec_debug_tc_scores": (_I, [_P, _P, _I, _I, _P, _I, _I, _P, _P, _L, _P]),
    "b200rec_predict_dense": (_I, [_P, _P, _I, _I, _P, _I, _I, _P, _P, _P, _P]),
    "b200rec_topk_rows": (_I, [_P, _L, _I, _I, _I, _P, _P]),
    "b200rec_holdout_metrics": (_I, [_P, _I, _I, _P, _P, _P, _P, _I, _P, _P]),
    "b200rec_loo_metrics": (_I, [_P, _I, _I, _P, _P, _P, _P, _I, _P, _P]),
    "b200rec_column_means": (_I, [_P, _L, _I, _P, _P]),
    "b200rec_spmm_csr": (_I, [_P, _P, _P, _I, _P, _I, _I, _P, _I, _P, _I, _F, _I, _P]),
    "b200rec_ngcf_layer_forward": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, C.c_uint64, C.c_uint64, _P, _P, _P, _F, _P]),
    "b200rec_ngcf_layer_backward": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, C.c_uint64, C.c_uint64, _F,
                                        _P, _P, _P, _P, _P, _P, _P]),
    "b200rec_spmm_csr_split": (_I, [_P, _P, _P, _I, _P, _I, _I, _P, _I, _P, _I, _F, _I, _L, _P, _P, _I, _P, _P, _I, _P, _P]),
}
EXPORTS = tuple(_PROTOS)

_lib = None


def lib():
    """The loaded library (loads on first use; raises if it was never built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "recsys_pytorch_b200 has no CPU / PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(handle, name)      # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(code):
    if code != OK:
        raise B200RecError(code, lib().b200rec_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device/host address of a tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t, name, dtype=None):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise B200RecError(ECUDA, f"{name} must be a CUDA tensor: the engine has no CPU path")
    if dtype is not None and t.dtype != dtype:
        raise B200RecError(EINVAL, f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise B200RecError(EINVAL, f"{name} must be contiguous")
    return t


def last_step_kernel() -> str:
    return lib().b200rec_last_step_kernel().decode()


def launch_count() -> int:
    return int(lib().b200rec_launch_count())
