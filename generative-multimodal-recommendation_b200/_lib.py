"""ctypes binding of libgmr.so (the C ABI of include/gmr.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError carrying
``gmr_last_error()`` is raised.  ``load()`` never compiles anything; building is
``genmmrec_b200.build.build()`` / ``__graft_entry__.build()``.
"""
import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libgmr.so")

GMR_SCORE_FP32 = 0
GMR_SCORE_TC = 1
GMR_SCORE_TC_SPLIT = 2
GMR_MAX_TOPK = 256
GMR_PEER_HANDLE_BYTES = 64

_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol include/gmr.h declares
SIGNATURES = {
    "gmr_last_error": (C.c_char_p, []),
    "gmr_abi_version": (C.c_int, []),
    "gmr_device_info": (C.c_int, [C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i64)]),
    "gmr_spmm_plan_create": (C.c_int, [C.POINTER(_vp), _vp, _i64, _i64, _i32, _vp]),
    "gmr_spmm_plan_destroy": (C.c_int, [_vp]),
    "gmr_spmm_plan_stats": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "gmr_spmm_workspace_bytes": (_i64, [_vp, _i32]),
    "gmr_spmm_csr_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i32, _f32, _f32, _vp, _i64, _vp]),
    "gmr_spmm_csr_f32_push": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _i64, _i64, _i32, _f32, _vp, _i64,
                                        _vp]),
    "gmr_spmm_blocked_plan_create": (C.c_int, [C.POINTER(_vp), _vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "gmr_spmm_blocked_plan_set_values": (C.c_int, [_vp, _vp, _vp]),
    "gmr_spmm_blocked_plan_destroy": (C.c_int, [_vp]),
    "gmr_spmm_blocked_plan_stats": (C.c_int, [_vp, C.POINTER(_i64)]),
    "gmr_spmm_blocked_workspace_bytes": (_i64, [_vp, _i32]),
    "gmr_spmm_blocked_f32": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i32, _f32, _f32, _vp, _i64, _vp]),
    "gmr_rows_push_f32": (C.c_int, [_vp, _i64, _i64, _i32, _vp, _i32, _i64, _i64, _vp]),
    "gmr_score_topk_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32, _i32]),
    "gmr_score_mask_topk_f32": (C.c_int, [_vp, _i64, _vp, _i32, _vp, _i64, _vp, _i32, _i32, _vp, _vp, _i32, _i32, _vp,
                                          _vp, _vp, _i64, _vp]),
    "gmr_score_tc_fallback_rows": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, C.POINTER(_i32), _vp]),
    "gmr_score_tc_stats": (C.c_int, [_vp, _i32, _i32, _i32, _i32, C.POINTER(C.c_uint64), _vp]),
    "gmr_scores_f32": (C.c_int, [_vp, _i64, _vp, _i32, _vp, _i64, _vp, _i32, _i32, _vp, _i64, _vp]),
    "gmr_hits_metrics_workspace_bytes": (_i64, [_i32, _i32]),
    "gmr_hits_metrics": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "gmr_rows_axpby_norm_f32": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _f32, _f32, _f32,
                                          _f32, _vp]),
    "gmr_rows_normalize_mix_f32": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _f32, _f32, _f32,
                                             _f32, _vp]),
    "gmr_dense_proj_workspace_bytes": (_i64, [_i32, _i32]),
    "gmr_dense_proj_f32": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _i64, _i32, _vp, _i64, _vp, _i64, _vp]),
    "gmr_bpr_scores_f32": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "gmr_bpr_scores_backward_f32": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _i64,
                                              _vp]),
    "gmr_spmm_plan_enable_narrow": (C.c_int, [_vp, _vp, _vp]),
    "gmr_spmm_narrow_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i32, _i32, _f32, _f32, _vp, _i64, _vp]),
    "gmr_spmm_sell_create": (C.c_int, [C.POINTER(_vp), _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "gmr_spmm_sell_set_values": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "gmr_spmm_sell_destroy": (C.c_int, [_vp]),
    "gmr_spmm_sell_bytes": (_i64, [_vp]),
    "gmr_spmm_sell_f32": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i32, _i32, _f32, _f32, _vp, _i64, _vp]),
    "gmr_cols_push_f32": (C.c_int, [_vp, _i64, _i64, _i32, _i64, _i64, _vp, _i32, _vp, _i64, _i64, _i64, _i32, _i64, _i64, _vp]),
    "gmr_slabs_to_rows_f32": (C.c_int, [_vp, _i32, _i64, _i64, _i32, _vp, _i64, _vp]),
    "gmr_rows_sumsq_f32": (C.c_int, [_vp, _i64, _i64, _i32, _vp, _vp]),
    "gmr_rows_axpby_ss_f32": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i32, _i64, _vp, _i64, _i64, _i32, _f32, _f32,
                                        _f32, _f32, _vp]),
    "gmr_peer_barrier": (C.c_int, [_vp, _i32, _i32, _vp, _vp]),
    "gmr_peer_alloc": (C.c_int, [C.POINTER(_vp), _i64]),
    "gmr_peer_free": (C.c_int, [_vp]),
    "gmr_peer_export": (C.c_int, [_vp, C.c_char_p]),
    "gmr_peer_open": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "gmr_peer_close": (C.c_int, [_vp]),
}

_lib = None


def load():
    """Load libgmr.so and declare every signature.  Raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libgmr.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for the CUDA hot path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == ABI drift, fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.gmr_abi_version() != 1:
        raise RuntimeError("libgmr.so ABI version %d, binding expects 1" % lib.gmr_abi_version())
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().gmr_last_error()
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))


def device_info():
    lib = load()
    sm, maj, mnr, l2 = _i32(), _i32(), _i32(), _i64()
    check(lib.gmr_device_info(C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(l2)), "gmr_device_info")
    return {"sm_count": sm.value, "cc": (maj.value, mnr.value), "l2_bytes": l2.value}
