"""ctypes binding of oracle/libgmr_oracle.so (the plain-C restatement, gmr_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of gmr_oracle.c.  ``build()`` compiles the library with
gcc through oracle/Makefile; nothing here touches CUDA.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libgmr_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_DIR, "gmr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _DIR, "-s", "libgmr_oracle.so"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def spmm_coo(rows, cols, vals, x, n_rows):
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    cols = np.ascontiguousarray(cols, dtype=np.int64)
    vals = np.ascontiguousarray(vals, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty((n_rows, x.shape[1]), dtype=np.float32)
    lib().oracle_spmm_coo_f32(_p(rows, C.c_int64), _p(cols, C.c_int64), _p(vals, C.c_float), C.c_int64(vals.size),
                              _p(x, C.c_float), C.c_int64(x.shape[1]), C.c_int32(x.shape[1]),
                              _p(y, C.c_float), C.c_int64(x.shape[1]), C.c_int64(n_rows))
    return y


def spmm_csr(rowptr, col, val, x, alpha=1.0, beta=0.0, y=None):
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    n_rows = rowptr.size - 1
    if y is None:
        y = np.zeros((n_rows, x.shape[1]), dtype=np.float32)
    else:
        y = np.ascontiguousarray(y, dtype=np.float32).copy()
    lib().oracle_spmm_csr_f32(_p(rowptr, C.c_int32), _p(col, C.c_int32), _p(val, C.c_float), _p(x, C.c_float),
                              C.c_int64(x.shape[1]), C.c_int32(x.shape[1]), _p(y, C.c_float),
                              C.c_int64(x.shape[1]), C.c_int64(n_rows), C.c_float(alpha), C.c_float(beta))
    return y


def spmm_csr_f64(rowptr, col, val, x):
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    n_rows = rowptr.size - 1
    y = np.empty((n_rows, x.shape[1]), dtype=np.float64)
    lib().oracle_spmm_csr_f64(_p(rowptr, C.c_int32), _p(col, C.c_int32), _p(val, C.c_float), _p(x, C.c_float),
                              C.c_int64(x.shape[1]), C.c_int32(x.shape[1]), _p(y, C.c_double),
                              C.c_int64(x.shape[1]), C.c_int64(n_rows))
    return y


def score_mask_topk(eu, users, ei, bias, mask_rowptr, mask_items, k):
    eu = np.ascontiguousarray(eu, dtype=np.float32)
    ei = np.ascontiguousarray(ei, dtype=np.float32)
    users = None if users is None else np.ascontiguousarray(users, dtype=np.int64)
    b = eu.shape[0] if users is None else users.size
    bias = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    if mask_rowptr is not None:
        mask_rowptr = np.ascontiguousarray(mask_rowptr, dtype=np.int64)
        mask_items = np.ascontiguousarray(mask_items, dtype=np.int32)
    ids = np.empty((b, k), dtype=np.int32)
    sc = np.empty((b, k), dtype=np.float32)
    lib().oracle_score_mask_topk_f32(_p(eu, C.c_float), C.c_int64(eu.shape[1]), _p(users, C.c_int64), C.c_int32(b),
                                     _p(ei, C.c_float), C.c_int64(ei.shape[1]), _p(bias, C.c_float),
                                     C.c_int32(ei.shape[0]), C.c_int32(ei.shape[1]), _p(mask_rowptr, C.c_int64),
                                     _p(mask_items, C.c_int32), C.c_int32(k), _p(ids, C.c_int32), _p(sc, C.c_float))
    return ids, sc


def hits(topk, gt_rowptr, gt_items):
    topk = np.ascontiguousarray(topk, dtype=np.int32)
    gt_rowptr = np.ascontiguousarray(gt_rowptr, dtype=np.int64)
    gt_items = np.ascontiguousarray(gt_items, dtype=np.int32)
    u, k = topk.shape
    hit = np.empty((u, k), dtype=np.uint8)
    lib().oracle_hits(_p(topk, C.c_int32), _p(gt_rowptr, C.c_int64), _p(gt_items, C.c_int32), C.c_int32(u),
                      C.c_int32(k), _p(hit, C.c_uint8))
    return hit


def metrics(hit, gt_len):
    hit = np.ascontiguousarray(hit, dtype=np.uint8)
    gt_len = np.ascontiguousarray(gt_len, dtype=np.int64)
    u, k = hit.shape
    out = [np.empty(k, dtype=np.float64) for _ in range(4)]
    lib().oracle_metrics(_p(hit, C.c_uint8), _p(gt_len, C.c_int64), C.c_int32(u), C.c_int32(k),
                         *[_p(o, C.c_double) for o in out])
    return dict(zip(("recall", "ndcg", "precision", "map"), out))
