"""Tensor-level operators over libgmr.so.

PyTorch is used here for device memory, streams and autograd bookkeeping only; every
computation below is a call into the hand-written sm_100a kernels through the C ABI
(``include/gmr.h``).  Nothing in this module has a CPU or eager-PyTorch fallback.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib

# Kernel launches issued through this module since import (bench.py reports the delta of a timed
# region as ``gpu_launches``).
LAUNCHES = 0

# When set to a list, every operator appends (name, meta, start_event, end_event): CUDA events recorded
# on the launching stream around the C-ABI call (bench.py derives per-kernel durations from them).
PROFILE = None

_workspaces = {}
# Superseded workspaces are kept alive once a CUDA graph has captured the step (Trainer.graphed): a captured kernel has the
# old buffer's address baked in, and handing that memory back to the allocator would let a later replay write into someone
# else's tensor.  Grow-only buffers and a handful of growth events bound what this holds.
_retired = []
_graphs_captured = 0


def note_graph_captured():
    """Called by Trainer.graphed after a capture: from now on replaced workspaces are retired, not freed."""
    global _graphs_captured
    _graphs_captured += 1


def _prof_begin():
    if PROFILE is None:
        return None
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


def _prof_end(name, ev, **meta):
    if ev is not None:
        end = torch.cuda.Event(enable_timing=True)
        end.record()
        PROFILE.append((name, meta, ev, end))


def _ws(device, nbytes, tag="default"):
    """Grow-only per-device scratch buffer (keeps allocations out of the timed/captured path)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None and _graphs_captured > 0:
            _retired.append(buf)
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _rows(t, name):
    """(pointer, leading dimension) of a 2-D fp32 CUDA tensor whose rows are contiguous."""
    if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2):
        raise TypeError("%s must be a 2-D float32 CUDA tensor, got %s %s on %s" % (name, t.dtype, tuple(t.shape), t.device))
    if t.shape[1] > 1 and t.stride(1) != 1:
        raise ValueError("%s must have unit stride along its last dimension" % name)
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))
    return _ptr(t), int(ld)


class GraphCSR:
    """A sparse fp32 matrix in CSR form on one GPU plus its SpMM plan (the row schedule).

    The reference keeps its graphs as uncoalesced COO tensors built with the legacy
    ``torch.sparse.FloatTensor`` ctor (GenMMRec/src/models/diffmm.py:105-107) and pays a coalesce +
    CSR conversion inside every ``torch.sparse.mm`` on CUDA; here the conversion happens once.
    Duplicate entries are kept (they add up, exactly like the uncoalesced COO product).
    """

    def __init__(self, rowptr, col, val, shape, chunk_nnz=0):
        assert rowptr.dtype == torch.int32 and col.dtype == torch.int32 and val.dtype == torch.float32
        assert rowptr.is_cuda and col.is_cuda and val.is_cuda
        self.rowptr, self.col, self.val = rowptr.contiguous(), col.contiguous(), val.contiguous()
        self.shape = (int(shape[0]), int(shape[1]))
        self.device = rowptr.device
        self.chunk_nnz = chunk_nnz
        self._plan = None
        self._bplans = {}   # block_cols -> [handle, version of `val` the plan snapshotted]
        self._t = None
        self._coo = None

    # ---- construction -------------------------------------------------------------------------
    @classmethod
    def from_coo(cls, indices, values, shape, device, chunk_nnz=0):
        """indices [2, nnz] (numpy or tensor, any int type), values [nnz]; entries keep their
        relative order inside each row (stable sort by row)."""
        idx = torch.as_tensor(np.asarray(indices) if not torch.is_tensor(indices) else indices)
        val = torch.as_tensor(np.asarray(values) if not torch.is_tensor(values) else values)
        idx = idx.to(device=device, dtype=torch.int64)
        val = val.to(device=device, dtype=torch.float32)
        n_rows = int(shape[0])
        if idx.shape[1] >= 2 ** 31 or max(shape) >= 2 ** 31:
            raise ValueError("graph exceeds the int32 index range of libgmr")
        if idx.shape[1] > 0 and bool((idx[0, 1:] < idx[0, :-1]).any()):
            order = torch.sort(idx[0], stable=True).indices
            idx, val = idx[:, order], val[order]
        counts = torch.bincount(idx[0], minlength=n_rows) if idx.shape[1] > 0 else \
            torch.zeros(n_rows, dtype=torch.int64, device=device)
        rowptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=device)
        rowptr[1:] = torch.cumsum(counts, 0)
        g = cls(rowptr.to(torch.int32), idx[1].to(torch.int32), val, shape, chunk_nnz)
        return g

    @classmethod
    def from_torch_sparse(cls, t, device=None, chunk_nnz=0):
        """From a (possibly uncoalesced) torch sparse COO tensor, e.g. the reference's own graphs."""
        device = device if device is not None else t.device
        return cls.from_coo(t._indices(), t._values(), tuple(t.shape), device, chunk_nnz)

    @property
    def nnz(self):
        return int(self.col.numel())

    def to_torch_coo(self):
        """Uncoalesced COO view (row-major order) for interop / checks."""
        counts = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(self.shape[0], device=self.device), counts)
        return torch.sparse_coo_tensor(torch.stack([rows, self.col.to(torch.int64)]), self.val, self.shape)

    def t(self):
        """Transposed graph (cached); needed by the SpMM backward, and equal to `self` only for
        symmetric matrices (norm_adj is; the edge-dropped modality graphs and the kNN graphs are not)."""
        if self._t is None:
            counts = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64)
            rows = torch.repeat_interleave(torch.arange(self.shape[0], device=self.device), counts)
            self._t = GraphCSR.from_coo(torch.stack([self.col.to(torch.int64), rows]), self.val,
                                        (self.shape[1], self.shape[0]), self.device, self.chunk_nnz)
            self._t._t = self
        return self._t

    def row_block(self, r0, r1):
        """CSR of rows [r0, r1) (all columns): the shard a rank owns under row sharding."""
        rp = self.rowptr[r0:r1 + 1].to(torch.int64)
        lo, hi = int(rp[0]), int(rp[-1])
        return GraphCSR((rp - lo).to(torch.int32), self.col[lo:hi].clone(), self.val[lo:hi].clone(),
                        (r1 - r0, self.shape[1]), self.chunk_nnz)

    # ---- plan -----------------------------------------------------------------------------------
    @property
    def plan(self):
        if self._plan is None:
            lib = _lib.load()
            h = C.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(lib.gmr_spmm_plan_create(C.byref(h), _ptr(self.rowptr), self.shape[0], self.shape[1],
                                                    self.chunk_nnz, _stream()), "gmr_spmm_plan_create")
            self._plan = h
        return self._plan

    @property
    def narrow_plan(self):
        """The plan with its length-sorted virtual-row order (what ``spmm_narrow`` walks), built once."""
        if not self.__dict__.get("_narrow_ready"):
            lib = _lib.load()
            with torch.cuda.device(self.device):
                _lib.check(lib.gmr_spmm_plan_enable_narrow(self.plan, _ptr(self.rowptr), _stream()), "gmr_spmm_plan_enable_narrow")
            self._narrow_ready = True
        return self.plan

    def sell(self, d, chains):
        """Sliced-ELL snapshot for ``spmm_narrow`` with rows of ``d`` floats (csrc/spmm.cu): built once per lane layout,
        values re-read when ``val`` was modified in place since."""
        lib = _lib.load()
        rpw = 32 // ((d // 4) * chains)
        if "_sells" not in self.__dict__:
            self._sells = {}
        key = (rpw, int(chains))   # the snapshot stores a block chain-major: one per lane layout AND chain count
        ent = self._sells.get(key)
        with torch.cuda.device(self.device):
            if ent is None:
                h = C.c_void_p()
                _lib.check(lib.gmr_spmm_sell_create(C.byref(h), self.plan, _ptr(self.rowptr), _ptr(self.col), _ptr(self.val),
                                                    int(d), int(chains), _stream()), "gmr_spmm_sell_create")
                ent = [h, self.val._version]
                self._sells[key] = ent
                self._narrow_ready = True
            elif ent[1] != self.val._version:
                _lib.check(lib.gmr_spmm_sell_set_values(ent[0], _ptr(self.rowptr), _ptr(self.col), _ptr(self.val), _stream()),
                           "gmr_spmm_sell_set_values")
                ent[1] = self.val._version
        return ent[0]

    def blocked_plan(self, block_cols):
        """Column-blocked plan (K1b, csrc/spmm_flat.cu) with blocks of `block_cols` columns.  The plan snapshots the
        matrix; if `val` was modified in place since (its torch version counter moved) the values are re-read."""
        lib = _lib.load()
        block_cols = int(min(max(1, block_cols), max(1, self.shape[1])))
        ent = self._bplans.get(block_cols)
        if ent is None:
            h = C.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(lib.gmr_spmm_blocked_plan_create(C.byref(h), _ptr(self.rowptr), _ptr(self.col), _ptr(self.val),
                                                            self.shape[0], self.shape[1], block_cols, _stream()),
                           "gmr_spmm_blocked_plan_create")
            ent = [h, self.val._version]
            self._bplans[block_cols] = ent
        elif ent[1] != self.val._version:
            with torch.cuda.device(self.device):
                _lib.check(lib.gmr_spmm_blocked_plan_set_values(ent[0], _ptr(self.val), _stream()),
                           "gmr_spmm_blocked_plan_set_values")
            ent[1] = self.val._version
        return ent[0]

    def blocked_plan_stats(self, block_cols):
        lib = _lib.load()
        out = (C.c_int64 * 8)()
        _lib.check(lib.gmr_spmm_blocked_plan_stats(self.blocked_plan(block_cols), out), "gmr_spmm_blocked_plan_stats")
        keys = ("blocks", "block_cols", "tiles", "segments", "long_segments", "slots", "reduce_small", "reduce_big")
        return dict(zip(keys, [int(v) for v in out]))

    def plan_stats(self):
        lib = _lib.load()
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(lib.gmr_spmm_plan_stats(self.plan, C.byref(a), C.byref(b), C.byref(c)), "gmr_spmm_plan_stats")
        return {"chunks": a.value, "split_rows": b.value, "nnz": c.value}

    def algorithmic_bytes(self, d):
        """CSR read once + X read once + Y written once."""
        return self.nnz * 8 + (self.shape[0] + 1) * 4 + self.shape[1] * d * 4 + self.shape[0] * d * 4

    def __del__(self):
        try:
            if self._plan is not None and _lib._lib is not None:
                _lib._lib.gmr_spmm_plan_destroy(self._plan)
            if _lib._lib is not None:
                for h, _ in self.__dict__.get("_sells", {}).values():
                    _lib._lib.gmr_spmm_sell_destroy(h)
                for h, _ in self._bplans.values():
                    _lib._lib.gmr_spmm_blocked_plan_destroy(h)
        except Exception:
            pass


# Column blocking (K1b, csrc/spmm_flat.cu) is OPT-IN: on the 1M x 500k workload it ties the row-centric kernels (2.47 vs
# 2.41 ms for the full graph, profiles/r02_spmm_flat_probe_v6.json) while cutting DRAM traffic 4.5x, so the default stays
# K1.  GMR_SPMM_BLOCKED: 0 (default) = row-centric kernels; auto = block when the gathered table n_cols * D * 4 exceeds
# GMR_SPMM_BLOCK_MIN_MB (96); 1 = always the nonzero-centric kernel.  GMR_SPMM_BLOCK_MB (48) is the slice of X one pass
# gathers from.
def _block_cols_for(a, d, x, out):
    mode = os.environ.get("GMR_SPMM_BLOCKED", "0")
    if mode == "0" or d % 4 != 0:
        return None
    if x.data_ptr() % 16 or out.data_ptr() % 16 or (x.shape[0] > 1 and x.stride(0) % 4) or (out.shape[0] > 1 and out.stride(0) % 4):
        return None
    row_bytes = d * 4
    table = a.shape[1] * row_bytes
    if mode != "1" and table <= (int(os.environ.get("GMR_SPMM_BLOCK_MIN_MB", "96")) << 20):
        return None
    target = int(os.environ.get("GMR_SPMM_BLOCK_MB", "48")) << 20
    return max(1, min(a.shape[1], target // row_bytes))


def spmm_raw(a, x, out=None, alpha=1.0, beta=0.0):
    """Y = alpha * A @ X + beta * Y through K1 (no autograd).  `out` may be a column slice of a
    wider row-major buffer; so may `x`."""
    global LAUNCHES
    lib = _lib.load()
    if x.shape[0] != a.shape[1]:
        raise ValueError("spmm: A is %s but X has %d rows" % (a.shape, x.shape[0]))
    d = int(x.shape[1])
    if out is None:
        if beta != 0.0:
            raise ValueError("spmm: beta != 0 needs an `out` to accumulate into")
        out = torch.empty((a.shape[0], d), dtype=torch.float32, device=x.device)
    xp, ldx = _rows(x, "X")
    yp, ldy = _rows(out, "Y")
    if out.shape != (a.shape[0], d):
        raise ValueError("spmm: out has shape %s, expected %s" % (tuple(out.shape), (a.shape[0], d)))
    bc = _block_cols_for(a, d, x, out) if a.shape[0] > 0 else None
    if bc is not None:
        return spmm_blocked(a, x, bc, out=out, alpha=alpha, beta=beta)
    plan = a.plan
    need = lib.gmr_spmm_workspace_bytes(plan, d)
    ws = _ws(x.device, need, "spmm") if need > 0 else None
    with torch.cuda.device(x.device):
        ev = _prof_begin()
        _lib.check(lib.gmr_spmm_csr_f32(plan, _ptr(a.rowptr), _ptr(a.col), _ptr(a.val), xp, ldx, yp, ldy, d,
                                        float(alpha), float(beta), _ptr(ws), need, _stream()), "gmr_spmm_csr_f32")
        _prof_end("spmm", ev, alg_bytes=a.algorithmic_bytes(d), nnz=a.nnz, d=d, rows=a.shape[0], cols=a.shape[1],
                  kernel="row")
    LAUNCHES += 1 + (1 if need > 0 else 0)
    return out


def spmm_blocked(a, x, block_cols, out=None, alpha=1.0, beta=0.0):
    """Y = alpha * A @ X + beta * Y through K1b (column-blocked, nonzero-centric; csrc/spmm_flat.cu) with blocks of
    `block_cols` columns.  D % 4 == 0 and 16-byte aligned rows required."""
    global LAUNCHES
    lib = _lib.load()
    if x.shape[0] != a.shape[1]:
        raise ValueError("spmm: A is %s but X has %d rows" % (a.shape, x.shape[0]))
    d = int(x.shape[1])
    if out is None:
        if beta != 0.0:
            raise ValueError("spmm: beta != 0 needs an `out` to accumulate into")
        out = torch.empty((a.shape[0], d), dtype=torch.float32, device=x.device)
    xp, ldx = _rows(x, "X")
    yp, ldy = _rows(out, "Y")
    if out.shape != (a.shape[0], d):
        raise ValueError("spmm: out has shape %s, expected %s" % (tuple(out.shape), (a.shape[0], d)))
    bplan = a.blocked_plan(block_cols)
    need = lib.gmr_spmm_blocked_workspace_bytes(bplan, d)
    ws = _ws(x.device, need, "spmm") if need > 0 else None
    with torch.cuda.device(x.device):
        ev = _prof_begin()
        _lib.check(lib.gmr_spmm_blocked_f32(bplan, xp, ldx, yp, ldy, d, float(alpha), float(beta), _ptr(ws), need,
                                            _stream()), "gmr_spmm_blocked_f32")
        _prof_end("spmm", ev, alg_bytes=a.algorithmic_bytes(d), nnz=a.nnz, d=d, rows=a.shape[0], cols=a.shape[1],
                  kernel="flat/blocked", block_cols=int(block_cols))
    LAUNCHES += -(-a.shape[1] // max(1, int(block_cols))) + 2
    return out


class _SpMM(torch.autograd.Function):
    """torch.sparse.mm(A, X) with constant A: backward is K1 on the transposed graph
    (operator contract of SURVEY.md section 8b)."""

    @staticmethod
    def forward(ctx, x, a):
        ctx.graph = a
        return spmm_raw(a, x.contiguous() if x.stride(-1) != 1 else x)

    @staticmethod
    def backward(ctx, grad):
        g = grad if grad.stride(-1) == 1 else grad.contiguous()
        return spmm_raw(ctx.graph.t(), g), None


def spmm(a, x):
    """Drop-in for ``torch.sparse.mm(A, X)`` / ``torch.spmm(A, X)`` with A a GraphCSR."""
    if x.requires_grad and torch.is_grad_enabled():
        return _SpMM.apply(x, a)
    return spmm_raw(a, x)


def rows_axpby_norm(x, y=None, z=None, a=1.0, b=0.0, c=0.0, eps=1e-12, out=None):
    """out = a*x + b*y + c * z / max(||z||_2, eps) row-wise (glue of diffmm.py:138-167)."""
    global LAUNCHES
    lib = _lib.load()
    if out is None:
        out = torch.empty_like(x)
    xp, ldx = _rows(x, "x")
    yp, ldy = _rows(y, "y") if y is not None else (C.c_void_p(0), 0)
    zp, ldz = _rows(z, "z") if z is not None else (C.c_void_p(0), 0)
    op, ldo = _rows(out, "out")
    with torch.cuda.device(x.device):
        _lib.check(lib.gmr_rows_axpby_norm_f32(xp, ldx, yp, ldy, zp, ldz, op, ldo, x.shape[0], x.shape[1], float(a),
                                               float(b), float(c), float(eps), _stream()), "gmr_rows_axpby_norm_f32")
    LAUNCHES += 1
    return out


def spmm_narrow(a, x, chains=2, out=None, alpha=1.0, beta=0.0, sell=None):
    """Y = alpha * A @ X + beta * Y for rows of 8 / 16 / 32 / 64 floats (the column shard a rank owns under
    ``dist.ColShardedDiffMM``): several virtual rows per warp, per-column operation order of the wide kernel
    (``chains=2``: what ``spmm_raw`` does up to 64 columns; ``chains=1``: what it does for 65..128), so the result equals
    the corresponding columns of the wide product bit for bit."""
    global LAUNCHES
    lib = _lib.load()
    if x.shape[0] != a.shape[1]:
        raise ValueError("spmm: A is %s but X has %d rows" % (a.shape, x.shape[0]))
    d = int(x.shape[1])
    if out is None:
        if beta != 0.0:
            raise ValueError("spmm: beta != 0 needs an `out` to accumulate into")
        out = torch.empty((a.shape[0], d), dtype=torch.float32, device=x.device)
    xp, ldx = _rows(x, "X")
    yp, ldy = _rows(out, "Y")
    if out.shape != (a.shape[0], d):
        raise ValueError("spmm: out has shape %s, expected %s" % (tuple(out.shape), (a.shape[0], d)))
    if sell is None:
        sell = os.environ.get("GMR_SPMM_SELL", "1") != "0"
    plan = a.plan
    snap = a.sell(d, chains) if sell else None   # sliced-ELL snapshot (default) or the CSR stream as it is
    if snap is None:
        plan = a.narrow_plan
    need = lib.gmr_spmm_workspace_bytes(plan, d)
    ws = _ws(x.device, need, "spmm") if need > 0 else None
    with torch.cuda.device(x.device):
        ev = _prof_begin()
        if snap is not None:
            _lib.check(lib.gmr_spmm_sell_f32(snap, xp, ldx, yp, ldy, d, int(chains), float(alpha), float(beta), _ptr(ws), need,
                                             _stream()), "gmr_spmm_sell_f32")
        else:
            _lib.check(lib.gmr_spmm_narrow_f32(plan, _ptr(a.rowptr), _ptr(a.col), _ptr(a.val), xp, ldx, yp, ldy, d, int(chains),
                                               float(alpha), float(beta), _ptr(ws), need, _stream()), "gmr_spmm_narrow_f32")
        _prof_end("spmm", ev, alg_bytes=a.algorithmic_bytes(d), nnz=a.nnz, d=d, rows=a.shape[0], cols=a.shape[1],
                  kernel="%s/%d" % ("sell" if snap is not None else "narrow", chains))
    LAUNCHES += 1 + (1 if need > 0 else 0)
    return out


def cols_push(src, dc, ptr_table, n_peers, ldy, src_col0=0, src_col_step=0, row_bounds=None, dst_row_offset=0, dst_col0=0, tag="",
              n_seg=1, src_seg_step=0, dst_seg_step=0):
    """Column-slice copies into peer buffers (``gmr_cols_push_f32``): for every peer p and source row r,
    ``peer[p][dst_row_offset + r - first_p, dst_col0 : dst_col0 + dc] = src[r, src_col0 + p * src_col_step : ... + dc]``;
    with ``row_bounds`` (device int64 [n_peers + 1]) row r goes only to the peer whose range holds it; ``n_seg`` segments
    (source columns ``src_seg_step`` apart, destination columns ``dst_seg_step`` apart) move in one launch."""
    global LAUNCHES
    lib = _lib.load()
    sp, ld = _rows(src, "src")
    with torch.cuda.device(src.device):
        ev = _prof_begin()
        _lib.check(lib.gmr_cols_push_f32(sp, ld, int(src.shape[0]), int(dc), int(src_col0), int(src_col_step),
                                         C.c_void_p(ptr_table.data_ptr()), int(n_peers), _ptr(row_bounds), int(dst_row_offset),
                                         int(ldy), int(dst_col0), int(n_seg), int(src_seg_step), int(dst_seg_step), _stream()),
                   "gmr_cols_push_f32")
        _prof_end("cols_push" + (":" + tag if tag else ""), ev,
                  bytes=4.0 * src.shape[0] * dc * n_seg * (1 if row_bounds is not None else n_peers),
                  peers=int(n_peers))
    LAUNCHES += 1


def slabs_to_rows(slabs, n_slabs, slab_rows, n_rows, dc, out):
    """out[r, p * dc : (p + 1) * dc] = slab p's row r, for the ``n_slabs`` contiguous [slab_rows, dc] slabs of ``slabs``."""
    global LAUNCHES
    lib = _lib.load()
    op, ldo = _rows(out, "out")
    with torch.cuda.device(out.device):
        ev = _prof_begin()
        _lib.check(lib.gmr_slabs_to_rows_f32(_ptr(slabs), int(n_slabs), int(slab_rows) * int(dc), int(n_rows), int(dc), op, ldo,
                                             _stream()), "gmr_slabs_to_rows_f32")
        _prof_end("slabs_to_rows", ev, bytes=8.0 * n_rows * dc * n_slabs)
    LAUNCHES += 1
    return out


def rows_sumsq(z, out=None):
    """out[r] = sum_d z[r, d]^2 (a rank's part of the squared row norms; D in 4, 8, 16, 32, 64)."""
    global LAUNCHES
    lib = _lib.load()
    zp, ldz = _rows(z, "z")
    if out is None:
        out = torch.empty(z.shape[0], dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device):
        _lib.check(lib.gmr_rows_sumsq_f32(zp, ldz, int(z.shape[0]), int(z.shape[1]), _ptr(out), _stream()), "gmr_rows_sumsq_f32")
    LAUNCHES += 1
    return out


def rows_axpby_ss(x, y=None, z=None, ss_parts=None, n_parts=1, ss_stride=0, a=1.0, b=0.0, c=0.0, eps=1e-12, out=None):
    """out = a*x + b*y + c * z / max(sqrt(ss), eps) with ss[r] the balanced-tree sum of ``ss_parts[k * ss_stride + r]``,
    k < n_parts: ``rows_axpby_norm`` for a column shard whose row norms were formed across the ranks."""
    global LAUNCHES
    lib = _lib.load()
    if out is None:
        out = torch.empty_like(x)
    xp, ldx = _rows(x, "x")
    yp, ldy = _rows(y, "y") if y is not None else (C.c_void_p(0), 0)
    zp, ldz = _rows(z, "z") if z is not None else (C.c_void_p(0), 0)
    op, ldo = _rows(out, "out")
    with torch.cuda.device(x.device):
        _lib.check(lib.gmr_rows_axpby_ss_f32(xp, ldx, yp, ldy, zp, ldz, _ptr(ss_parts), int(n_parts), int(ss_stride), op, ldo,
                                             x.shape[0], x.shape[1], float(a), float(b), float(c), float(eps), _stream()),
                   "gmr_rows_axpby_ss_f32")
    LAUNCHES += 1
    return out


def rows_normalize_mix(x1, x2, w1, w2, y=None, slope=1.0, eps=1e-12, out=None):
    """out[:, :D] = w1 * normalize(lrelu(x1)) + w2 * normalize(lrelu(x2)); with ``y``, out[:, D:2D] = out[:, :D] + y
    (modality mix of diffmm.py:131-145 in one pass)."""
    global LAUNCHES
    lib = _lib.load()
    n, d = int(x1.shape[0]), int(x1.shape[1])
    width = 2 * d if y is not None else d
    if out is None:
        out = torch.empty((n, width), dtype=torch.float32, device=x1.device)
    p1, ld1 = _rows(x1, "x1")
    p2, ld2 = _rows(x2, "x2")
    py, ldy = _rows(y, "y") if y is not None else (C.c_void_p(0), 0)
    po, ldo = _rows(out, "out")
    if out.shape[0] != n or out.shape[1] < width:
        raise ValueError("rows_normalize_mix: out has shape %s, need [%d, >= %d]" % (tuple(out.shape), n, width))
    with torch.cuda.device(x1.device):
        _lib.check(lib.gmr_rows_normalize_mix_f32(p1, ld1, p2, ld2, py, ldy, po, ldo, n, d, float(w1), float(w2),
                                                  float(slope), float(eps), _stream()), "gmr_rows_normalize_mix_f32")
    LAUNCHES += 1
    return out


def dense_proj_supported(a, w):
    k, n = int(w.shape[0]), int(w.shape[1])
    return (n == 64 and k % 32 == 0 and a.dim() == 2 and a.shape[1] == k and a.stride(1) == 1 and a.stride(0) % 4 == 0
            and a.data_ptr() % 16 == 0 and a.dtype == torch.float32 and w.dtype == torch.float32)


def dense_proj(a, w, out=None):
    """C = A @ W (fp32, fp32-level accuracy) through the split-TF32 tcgen05 kernel (K3 family).  A [M, K] with
    contiguous rows, W [K, 64].  Shapes the kernel does not take (N != 64, K % 32 != 0) raise; callers that want the
    library GEMM for those check ``dense_proj_supported`` first."""
    global LAUNCHES
    lib = _lib.load()
    m, k, n = int(a.shape[0]), int(a.shape[1]), int(w.shape[1])
    if out is None:
        out = torch.empty((m, n), dtype=torch.float32, device=a.device)
    ap, lda = _rows(a, "A")
    wc = w if (w.stride(1) == 1) else w.contiguous()
    wp, ldw = _rows(wc, "W")
    cp, ldc = _rows(out, "C")
    need = lib.gmr_dense_proj_workspace_bytes(k, n)
    ws = _ws(a.device, need, "proj")
    with torch.cuda.device(a.device):
        _lib.check(lib.gmr_dense_proj_f32(ap, lda, m, k, wp, ldw, n, cp, ldc, _ptr(ws), need, _stream()),
                   "gmr_dense_proj_f32")
    LAUNCHES += 2
    return out


class _BprScores(torch.autograd.Function):
    """(pos_score, neg_score) of BPR triples without materialising the three [B, D] gathers (csrc/train_ops.cu)."""

    @staticmethod
    def forward(ctx, eu, ei, users, pos, neg):
        global LAUNCHES
        lib = _lib.load()
        eup, ldu = _rows(eu, "eu")
        eip, ldi = _rows(ei, "ei")
        b, d = int(users.numel()), int(eu.shape[1])
        ps = torch.empty(b, dtype=torch.float32, device=eu.device)
        ns = torch.empty(b, dtype=torch.float32, device=eu.device)
        with torch.cuda.device(eu.device):
            _lib.check(lib.gmr_bpr_scores_f32(eup, ldu, eip, ldi, _ptr(users), _ptr(pos), _ptr(neg), b, d, _ptr(ps), _ptr(ns),
                                              _stream()), "gmr_bpr_scores_f32")
        LAUNCHES += 1
        ctx.save_for_backward(eu, ei, users, pos, neg)
        return ps, ns

    @staticmethod
    def backward(ctx, g_pos, g_neg):
        global LAUNCHES
        lib = _lib.load()
        eu, ei, users, pos, neg = ctx.saved_tensors
        eup, ldu = _rows(eu, "eu")
        eip, ldi = _rows(ei, "ei")
        d_eu, d_ei = torch.zeros_like(eu, memory_format=torch.contiguous_format), torch.zeros_like(ei, memory_format=torch.contiguous_format)
        b, d = int(users.numel()), int(eu.shape[1])
        with torch.cuda.device(eu.device):
            _lib.check(lib.gmr_bpr_scores_backward_f32(eup, ldu, eip, ldi, _ptr(users), _ptr(pos), _ptr(neg), b, d,
                                                       _ptr(g_pos.contiguous()), _ptr(g_neg.contiguous()), _ptr(d_eu),
                                                       int(d_eu.stride(0)), _ptr(d_ei), int(d_ei.stride(0)), _stream()),
                       "gmr_bpr_scores_backward_f32")
        LAUNCHES += 1
        return d_eu, d_ei, None, None, None


def bpr_scores(eu, ei, users, pos, neg):
    """``((eu[users] * ei[pos]).sum(-1), (eu[users] * ei[neg]).sum(-1))``, differentiable w.r.t. ``eu`` and ``ei``, as one
    fused gather + dot kernel forward and one fused scatter kernel backward (the BPR terms of every model's
    ``calculate_loss``)."""
    def ok(t):
        return t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and (t.shape[1] == 1 or t.stride(1) == 1)

    if not (ok(eu) and ok(ei)):
        eu, ei = eu.contiguous().float(), ei.contiguous().float()
    idx = [t.to(torch.int64).contiguous() for t in (users, pos, neg)]
    return _BprScores.apply(eu, ei, *idx)


def linear_proj(x, linear):
    """``linear(x)`` for an ``nn.Linear`` with 64 outputs over a wide feature table (GUME's image/text_reduce_dim,
    gume.py:232-233; VBPR's item_linear, vbpr.py:69-75): the split-TF32 tcgen05 kernel under ``no_grad`` when the shape
    fits it, the library GEMM otherwise (training needs autograd; other shapes are not implemented)."""
    w = linear.weight
    if (not torch.is_grad_enabled()) and x.is_cuda and x.dim() == 2 and w.shape[0] == 64 and w.shape[1] % 32 == 0:
        wt = w.detach().t().contiguous()           # [K, 64]
        xc = x.detach()
        if dense_proj_supported(xc, wt):
            out = dense_proj(xc, wt)
            if linear.bias is not None:
                out += linear.bias.detach()
            return out
    return linear(x)


def tc_supported(d, k, precision="tc"):
    """Shapes the tcgen05 scoring paths accept (operand tiles + ring stages must fit in shared memory)."""
    if precision == "tc_split":   # D > 256 runs K-chunked: both operands streamed, accumulation over 3 D / 64 atoms in TMEM
        return d % 64 == 0 and 64 <= d <= 8192 and k <= 248
    return d % 64 == 0 and 64 <= d <= 256 and k <= 256


_PRECISIONS = {"fp32": _lib.GMR_SCORE_FP32, "tc": _lib.GMR_SCORE_TC, "tc_split": _lib.GMR_SCORE_TC_SPLIT}


def score_mask_topk(eu, ei, k, users=None, bias=None, mask_rowptr=None, mask_items=None, precision="fp32",
                    return_scores=True):
    """Fused full-sort scoring + train-history mask + top-K (K2).

    eu [*, D], ei [I, D] fp32; users int64 [B] row ids into eu (None: all rows in order);
    mask_rowptr int64 [B+1] / mask_items int32 ascending within each row.  Returns
    (ids int32 [B, k], scores fp32 [B, k] or None) ordered by (score desc, item id asc).
    """
    global LAUNCHES
    lib = _lib.load()
    if precision == "auto":  # tensor cores whenever the shape is supported; every path returns the same result
        dd = int(ei.shape[1])
        precision = "tc" if tc_supported(dd, k) else ("tc_split" if tc_supported(dd, k, "tc_split") else "fp32")
    mode = _PRECISIONS[precision]
    eup, lde_u = _rows(eu, "eu")
    eip, lde_i = _rows(ei, "ei")
    if eu.shape[1] != ei.shape[1]:
        raise ValueError("score_mask_topk: embedding widths differ (%d vs %d)" % (eu.shape[1], ei.shape[1]))
    b = int(users.numel()) if users is not None else int(eu.shape[0])
    i, d = int(ei.shape[0]), int(ei.shape[1])
    if users is not None and not (users.dtype == torch.int64 and users.is_cuda and users.is_contiguous()):
        raise TypeError("users must be a contiguous int64 CUDA tensor")
    if bias is not None and not (bias.dtype == torch.float32 and bias.is_cuda and bias.is_contiguous() and bias.numel() == i):
        raise TypeError("bias must be a contiguous float32 CUDA tensor of length n_items")
    if mask_rowptr is not None:
        if not (mask_rowptr.dtype == torch.int64 and mask_rowptr.numel() == b + 1 and mask_rowptr.is_contiguous()):
            raise TypeError("mask_rowptr must be contiguous int64 of length B + 1")
        if not (mask_items.dtype == torch.int32 and mask_items.is_contiguous()):
            raise TypeError("mask_items must be contiguous int32")
    ids = torch.empty((b, k), dtype=torch.int32, device=eu.device)
    scores = torch.empty((b, k), dtype=torch.float32, device=eu.device) if return_scores else None
    need = lib.gmr_score_topk_workspace_bytes(b, i, d, k, mode)
    ws = _ws(eu.device, need, "score")
    with torch.cuda.device(eu.device):
        ev = _prof_begin()
        _lib.check(lib.gmr_score_mask_topk_f32(eup, lde_u, _ptr(users), b, eip, lde_i, _ptr(bias), i, d,
                                               _ptr(mask_rowptr), _ptr(mask_items), k, mode, _ptr(ids), _ptr(scores),
                                               _ptr(ws), need, _stream()), "gmr_score_mask_topk_f32")
        _prof_end("score_topk", ev, flops=2.0 * b * i * d, users=b, items=i, d=d, k=k, precision=precision)
    # own kernels only (the cub radix sort of the item norms is library code and not counted):
    # tc: item norms + 2 operand-prep kernels + 2 sweeps + checkpoint + finalize + fp32 redo (+ bias absmax);
    # tc_split: 2 split kernels + fused + redo
    LAUNCHES += {_lib.GMR_SCORE_FP32: 1, _lib.GMR_SCORE_TC: 8 + (1 if bias is not None else 0), _lib.GMR_SCORE_TC_SPLIT: 4}[mode]
    global _last_score_call
    _last_score_call = (ws, b, i, d, k, mode)
    return ids, scores


_last_score_call = None


def last_tc_fallback_rows():
    """Rows of the most recent ``precision='tc'`` call that were redone on the exact fp32 path because
    their candidate margin could not certify exactness (diagnostic; synchronises)."""
    if _last_score_call is None or _last_score_call[5] == _lib.GMR_SCORE_FP32:
        return 0
    ws, b, i, d, k, mode = _last_score_call
    out = C.c_int32(0)
    _lib.check(_lib.load().gmr_score_tc_fallback_rows(_ptr(ws), b, i, d, k, mode, C.byref(out), _stream()),
               "gmr_score_tc_fallback_rows")
    return int(out.value)


def last_tc_stats():
    """Counters of the most recent ``precision='tc'`` call (all zero unless GMR_SCREEN_STATS is set in the
    environment): append-path chunks, appended candidates, cheap/exact prunes, re-scored candidates, item tiles swept, rows settled by
    the exact head."""
    if _last_score_call is None or _last_score_call[5] != _lib.GMR_SCORE_TC:
        return {}
    ws, b, i, d, k, _ = _last_score_call
    out = (C.c_uint64 * 8)()
    _lib.check(_lib.load().gmr_score_tc_stats(_ptr(ws), b, i, d, k, out, _stream()), "gmr_score_tc_stats")
    return {"slow_chunks": out[0], "appends": out[1], "cheap_prunes": out[2], "exact_prunes": out[3],
            "head_pairs_64bit": out[4], "rescored": out[5], "tiles_swept": out[6], "head_rows": out[7]}


def scores_dense(eu, ei, users=None, bias=None):
    """fp32 [B, I] score matrix (API parity with ``full_sort_predict``; the fused evaluation never
    materialises it)."""
    global LAUNCHES
    lib = _lib.load()
    eup, lde_u = _rows(eu, "eu")
    eip, lde_i = _rows(ei, "ei")
    b = int(users.numel()) if users is not None else int(eu.shape[0])
    i, d = int(ei.shape[0]), int(ei.shape[1])
    out = torch.empty((b, i), dtype=torch.float32, device=eu.device)
    with torch.cuda.device(eu.device):
        _lib.check(lib.gmr_scores_f32(eup, lde_u, _ptr(users), b, eip, lde_i, _ptr(bias), i, d, _ptr(out), i,
                                      _stream()), "gmr_scores_f32")
    LAUNCHES += 1
    return out


def hits_metrics(topk, gt_rowptr, gt_items, return_hit=False):
    """Hit matrix + per-position SUMS over users of recall / ndcg / precision / map (K4).
    topk int32 [U, K]; ground truth CSR (int64 rowptr, int32 items ascending within a row)."""
    global LAUNCHES
    lib = _lib.load()
    if not (topk.dtype == torch.int32 and topk.is_cuda and topk.is_contiguous() and topk.dim() == 2):
        raise TypeError("topk must be a contiguous int32 CUDA tensor [U, K]")
    u, k = int(topk.shape[0]), int(topk.shape[1])
    if not (gt_rowptr.dtype == torch.int64 and gt_rowptr.numel() == u + 1 and gt_items.dtype == torch.int32):
        raise TypeError("ground truth must be (int64 rowptr [U+1], int32 items)")
    hit = torch.empty((u, k), dtype=torch.uint8, device=topk.device) if return_hit else None
    sums = torch.empty((4, k), dtype=torch.float64, device=topk.device)
    need = lib.gmr_hits_metrics_workspace_bytes(u, k)
    ws = _ws(topk.device, need, "metrics")
    with torch.cuda.device(topk.device):
        ev = _prof_begin()
        _lib.check(lib.gmr_hits_metrics(_ptr(topk), _ptr(gt_rowptr.contiguous()), _ptr(gt_items.contiguous()), u, k,
                                        _ptr(hit), _ptr(sums), _ptr(ws), need, _stream()), "gmr_hits_metrics")
        _prof_end("hits_metrics", ev, users=u, k=k)
    LAUNCHES += 2
    return sums, hit
