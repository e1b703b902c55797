"""One-process-per-GPU sharding of the hot path.

Propagation is row-sharded: rank g owns a contiguous block of user rows and a contiguous block of item
rows of every graph.  A layer output that the NEXT SpMM needs in full is produced by the fused kernel
``gmr_spmm_csr_f32_push``: each rank computes its rows and stores them straight into every peer's copy
of the gathered operand through NVLink peer memory (CUDA IPC mappings) -- the all-gather rides inside
the SpMM epilogue, tile by tile, instead of following it as a separate NCCL collective.  A tiny
stream-ordered NCCL all-reduce then acts as the cross-rank barrier.  Evaluation is sharded by the same
user blocks; every rank scores its users against the full item table and only the [4, K] float64
metric sums are all-reduced.

``torch.distributed`` is plumbing here (rendezvous, handle exchange, barrier, the small dense
all-gathers around the torch projections); the data path of the gathered SpMM operand is our kernel.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib, ops


def world_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def block_bounds(n, world):
    """Contiguous, near-equal blocks: bounds[g] .. bounds[g + 1]."""
    return [n * g // world for g in range(world + 1)]


def nnz_balanced_bounds(rowptr, world):
    """Row blocks holding about the same number of nonzeros (power-law rows make equal row counts
    unbalanced).  ``rowptr`` is a host or device int tensor of length n_rows + 1."""
    rp = rowptr.to(torch.int64).cpu()
    total = int(rp[-1])
    targets = torch.tensor([total * g // world for g in range(world + 1)], dtype=torch.int64)
    b = torch.searchsorted(rp, targets).tolist()
    b[0], b[-1] = 0, rp.numel() - 1
    for g in range(1, world + 1):
        b[g] = max(b[g], b[g - 1])
    return b


class _RawCudaBuffer(object):
    """Expose a raw device pointer to torch through the CUDA array interface."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


class PeerBuffer(object):
    """A [rows, cols] fp32 buffer replicated on every rank, each replica writable by all peers.

    Allocation is a plain ``cudaMalloc`` inside libgmr (IPC-exportable, unlike pool memory); the 64-byte
    handles travel through ``torch.distributed``; ``ptr_table`` is a device array with the address of
    every rank's replica as seen from THIS process (own replica = local pointer)."""

    def __init__(self, rows, cols, device, group=None):
        import torch.distributed as dist

        self.shape = (int(rows), int(cols))
        self.device = torch.device(device)
        lib = _lib.load()
        nbytes = max(rows * cols * 4, 256)
        self._base = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.gmr_peer_alloc(C.byref(self._base), nbytes), "gmr_peer_alloc")
        self.tensor = torch.as_tensor(_RawCudaBuffer(self._base.value, self.shape), device=self.device)
        self.tensor.zero_()
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._opened = []
        ptrs = [0] * self.world
        ptrs[self.rank] = self._base.value
        if self.world > 1:
            handle = C.create_string_buffer(_lib.GMR_PEER_HANDLE_BYTES)
            _lib.check(lib.gmr_peer_export(self._base, handle), "gmr_peer_export")
            mine = torch.tensor(list(handle.raw), dtype=torch.uint8, device=self.device)
            gathered = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(gathered, mine, group=group)
            for g in range(self.world):
                if g == self.rank:
                    continue
                raw = bytes(gathered[g].cpu().tolist())
                p = C.c_void_p()
                with torch.cuda.device(self.device):
                    _lib.check(lib.gmr_peer_open(raw, C.byref(p)), "gmr_peer_open")
                self._opened.append(p)
                ptrs[g] = p.value
        self.ptr_table = torch.tensor(ptrs, dtype=torch.int64, device=self.device)

    def close(self):
        lib = _lib.load()
        for p in self._opened:
            lib.gmr_peer_close(p)
        self._opened = []
        if self._base is not None and self._base.value:
            self.tensor = None
            lib.gmr_peer_free(self._base)
            self._base = None


def spmm_push(a, x, ptr_table, n_peers, row_offset, ldy, alpha=1.0):
    """Fused SpMM + all-gather: rows of ``alpha * A @ X`` go to ``peer[p] + (row_offset + r) * ldy`` for every
    peer p (K1 push variant).  ``ptr_table`` is a device int64 array of peer base addresses (a column
    offset may already be folded into the addresses)."""
    lib = _lib.load()
    d = int(x.shape[1])
    xp, ldx = ops._rows(x, "X")
    plan = a.plan
    need = lib.gmr_spmm_workspace_bytes(plan, d)
    ws = ops._ws(x.device, need, "spmm") if need > 0 else None
    with torch.cuda.device(x.device):
        ev = ops._prof_begin()
        _lib.check(lib.gmr_spmm_csr_f32_push(plan, ops._ptr(a.rowptr), ops._ptr(a.col), ops._ptr(a.val), xp, ldx,
                                             C.c_void_p(ptr_table.data_ptr()), int(n_peers), int(row_offset), int(ldy), d,
                                             float(alpha), ops._ptr(ws), need, ops._stream()), "gmr_spmm_csr_f32_push")
        ops._prof_end("spmm_push", ev, alg_bytes=a.algorithmic_bytes(d), nnz=a.nnz, d=d, rows=a.shape[0], cols=a.shape[1],
                      peers=int(n_peers))
    ops.LAUNCHES += 1 + (1 if need > 0 else 0)


def stream_barrier(token):
    """Cross-rank barrier ordered on the current CUDA stream (no host synchronisation): a 1-element NCCL
    all-reduce completes only after every rank has enqueued -- and therefore finished -- the work before it."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(token)


def all_gather_rows(local, sizes, group=None):
    """Concatenate row blocks of unequal heights from all ranks (NCCL all-gather on a padded buffer)."""
    import torch.distributed as dist

    world = len(sizes)
    if world == 1:
        return local
    cols, mx = local.shape[1], max(sizes)
    padded = local if local.shape[0] == mx else torch.cat(
        [local, torch.zeros((mx - local.shape[0], cols), dtype=local.dtype, device=local.device)])
    out = torch.empty((world * mx, cols), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if all(sz == mx for sz in sizes):
        return out
    return torch.cat([out[g * mx:g * mx + sizes[g]] for g in range(world)])


class ShardedDiffMM(object):
    """Row-sharded ``forward_MM`` of a DiffMM model replica (parameters replicated, graphs sharded).

    Rank g owns users [u0, u1) and items [i0, i1).  Same regrouped dataflow as the single-GPU
    ``DiffMM._forward_mm_fused``, cut by output rows:

        [Z | Z + I0][i-block]   sharded projections + gmr_rows_normalize_mix     NCCL all-gather -> Xi [I, 2d]
        Hz_g      = R_hat[g] . Z            K1 PUSH into every peer's Hz [U, d]   (fused all-gather over NVLink)
        modal_u,g = R_hat[g] . (Z + I0)     K1, local
        modal_i,g = R_hat'[g] . (U0 + Hz)   K1, local
        modal_g  += lambda (w0 A_v + w1 A_t)[g] . [U0; I0]                        K1, local, in place
        modal     = NCCL all-gather of the row blocks;   L_g = A[g] . modal       K1, local
        E_g       = modal_g + L_g + ris_lambda n(modal_g)

    and the result rows (users of the block, items of the block) are returned with the block bounds.
    """

    def __init__(self, model, group=None):
        import torch.distributed as dist

        self.model = model
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        m = model
        self.dev = m.device
        nu, ni, d = m.n_users, m.n_items, m.latdim
        adj = m.norm_adj
        self.ub = nnz_balanced_bounds(adj.ui.rowptr, self.world)
        self.ib = nnz_balanced_bounds(adj.iu.rowptr, self.world)
        g = self.rank
        self.u0, self.u1, self.i0, self.i1 = self.ub[g], self.ub[g + 1], self.ib[g], self.ib[g + 1]
        self.r_ui = adj.ui.row_block(self.u0, self.u1)       # [U_g, I]
        self.r_iu = adj.iu.row_block(self.i0, self.i1)       # [I_g, U]
        self.hz = PeerBuffer(nu, d, self.dev, group)         # R_hat Z, gathered by the push SpMM
        self.token = torch.zeros(1, device=self.dev)
        self._graphs_version = None
        self._mix_sig = None

    def _graph_blocks(self, w0, w1):
        m = self.model
        nu = m.n_users
        full = m.norm_adj.full
        if self._graphs_version is not full:
            self.adj_u = full.row_block(self.u0, self.u1)
            self.adj_i = full.row_block(nu + self.i0, nu + self.i1)
            self._graphs_version = full
        mix = m._modal_mix_graph(m.image_UI_matrix, m.text_UI_matrix, w0, w1)
        sig = (w0, w1, m.ris_adj_lambda, id(mix))
        if self._mix_sig != sig:
            self.mix_u = mix.row_block(self.u0, self.u1)
            self.mix_i = mix.row_block(nu + self.i0, nu + self.i1)
            self._mix_sig = sig
        return self

    @torch.no_grad()
    def forward_MM(self):
        m = self.model
        nu, ni, d = m.n_users, m.n_items, m.latdim
        u0, u1, i0, i1 = self.u0, self.u1, self.i0, self.i1
        w0, w1 = m._modal_weights_host()
        self._graph_blocks(w0, w1)
        e0 = m._packed_e0()
        U0, I0 = e0[:nu], e0[nu:]
        usz = [self.ub[g + 1] - self.ub[g] for g in range(self.world)]
        isz = [self.ib[g + 1] - self.ib[g] for g in range(self.world)]
        # 1. sharded projections -> gathered Xi = [Z | Z + I0]
        pv = torch.mm(m.v_feat[i0:i1], m.image_trans.detach())
        pt = torch.mm(m.t_feat[i0:i1], m.text_trans.detach())
        xi_local = ops.rows_normalize_mix(pv, pt, w0, w1, y=I0[i0:i1], slope=m.leakyrelu.negative_slope)
        xi = all_gather_rows(xi_local, isz, self.group)
        # 2. user rows: R_hat Z pushed to every replica of Hz (fused all-gather); modal_u stays local
        spmm_push(self.r_ui, xi[:, :d], self.hz.ptr_table, self.world, u0, d)
        modal_u = ops.spmm_raw(self.r_ui, xi[:, d:])
        stream_barrier(self.token)
        # 3. item rows: modal_i = R_hat' (U0 + R_hat Z)
        xu = torch.add(U0, self.hz.tensor)
        modal_i = ops.spmm_raw(self.r_iu, xu)
        stream_barrier(self.token)   # nobody may overwrite Hz (next call) before every rank has read it
        # 4. modality graphs, accumulated in place
        ops.spmm_raw(self.mix_u, e0, out=modal_u, beta=1.0)
        ops.spmm_raw(self.mix_i, e0, out=modal_i, beta=1.0)
        # 5. GCN layers over the full adjacency
        acc_u = acc_i = None
        last_u, last_i = modal_u, modal_i
        for layer in range(m.gnn_layer):
            full = torch.cat([all_gather_rows(last_u, usz, self.group), all_gather_rows(last_i, isz, self.group)])
            last_u, last_i = ops.spmm_raw(self.adj_u, full), ops.spmm_raw(self.adj_i, full)
            if layer == 0:
                acc_u = ops.rows_axpby_norm(modal_u, last_u, modal_u, a=1.0, b=1.0, c=m.ris_lambda)
                acc_i = ops.rows_axpby_norm(modal_i, last_i, modal_i, a=1.0, b=1.0, c=m.ris_lambda)
            else:
                ops.rows_axpby_norm(acc_u, last_u, None, a=1.0, b=1.0, out=acc_u)
                ops.rows_axpby_norm(acc_i, last_i, None, a=1.0, b=1.0, out=acc_i)
        if acc_u is None:
            acc_u = ops.rows_axpby_norm(modal_u, None, modal_u, a=1.0, c=m.ris_lambda)
            acc_i = ops.rows_axpby_norm(modal_i, None, modal_i, a=1.0, c=m.ris_lambda)
        return acc_u, acc_i

    def close(self):
        """Release the peer-mapped buffers (CUDA IPC mappings of the other ranks' replicas)."""
        self.hz.close()

    @torch.no_grad()
    def eval_factors(self):
        """(user rows of this rank's block, full item table): the item block is all-gathered."""
        out_u, out_i = self.forward_MM()
        isz = [self.ib[g + 1] - self.ib[g] for g in range(self.world)]
        return out_u, all_gather_rows(out_i, isz, self.group)


def shard_eval_by_user_block(loader, u0, u1):
    """The eval users whose id falls in [u0, u1), in loader order, with their mask / ground-truth CSR
    slices and user ids rebased to the block (so they index the rank-local user rows)."""
    sel = torch.nonzero((loader.eval_u >= u0) & (loader.eval_u < u1)).flatten()

    class Shard(object):
        pass

    s = Shard()
    s.positions = sel
    s.eval_u = (loader.eval_u[sel] - u0).contiguous()
    for ptr, items in (("mask_rowptr", "mask_items"), ("gt_rowptr", "gt_items")):
        rp = getattr(loader, ptr)
        lens = (rp[1:] - rp[:-1])[sel]
        new_rp = torch.zeros(sel.numel() + 1, dtype=torch.int64, device=rp.device)
        new_rp[1:] = torch.cumsum(lens, 0)
        total = int(new_rp[-1])
        src = torch.repeat_interleave(rp[:-1][sel] - new_rp[:-1], lens) + torch.arange(total, device=rp.device)
        setattr(s, ptr, new_rp)
        setattr(s, items, getattr(loader, items)[src].contiguous())
    s.eval_len_list = np.asarray(loader.eval_len_list)[sel.cpu().numpy()]
    s.get_eval_len_list = lambda: s.eval_len_list
    return s
