"""One-process-per-GPU sharding of the hot path.

Propagation is row-sharded: rank g owns a contiguous block of user rows and a contiguous block of item
rows of every graph.  A layer output that the NEXT SpMM needs in full lives in a buffer replicated on
every rank (CUDA IPC peer mappings); each rank computes its row block with the local SpMM kernel (K1) and
then stores the block into every peer's replica over NVLink with ``gmr_rows_push_f32`` (a copy kernel on a
second stream, overlapped with the SpMMs that do not depend on it).  A tiny stream-ordered NCCL all-reduce
then acts as the cross-rank barrier.  The fused variant ``gmr_spmm_csr_f32_push`` (stores from the SpMM
epilogue) exists and is tested, but is NOT on this path: remote stores stall the gather warps and it
measured slower than SpMM + copy kernel.  Evaluation is sharded by the same user
blocks; every rank scores its users against the full item table and only the [4, K] float64 metric sums
are all-reduced.

``torch.distributed`` is plumbing here (rendezvous, handle exchange, barriers); the data path of the
gathered SpMM operand is our kernels over peer memory.
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
    ws = ops._ws(x.device, need, "spmm_push") if need > 0 else None   # own buffer: may run beside spmm_raw on another stream
    with torch.cuda.device(x.device):
        ev = ops._prof_begin()
        _lib.check(lib.gmr_spmm_csr_f32_push(plan, ops._ptr(a.rowptr), ops._ptr(a.col), ops._ptr(a.val), xp, ldx,
                                             C.c_void_p(ptr_table.data_ptr()), int(n_peers), int(row_offset), int(ldy), d,
                                             float(alpha), ops._ptr(ws), need, ops._stream()), "gmr_spmm_csr_f32_push")
        ops._prof_end("spmm_push", ev, alg_bytes=a.algorithmic_bytes(d), nnz=a.nnz, d=d, rows=a.shape[0], cols=a.shape[1],
                      peers=int(n_peers))
    ops.LAUNCHES += 1 + (1 if need > 0 else 0)


def rows_push(x, ptr_table, n_peers, row_offset, ldy):
    """Dense all-gather by peer stores: the rows of ``x`` land at row ``row_offset`` of every peer's replica."""
    lib = _lib.load()
    xp, ldx = ops._rows(x, "x")
    with torch.cuda.device(x.device):
        ev = ops._prof_begin()
        _lib.check(lib.gmr_rows_push_f32(xp, ldx, int(x.shape[0]), int(x.shape[1]), C.c_void_p(ptr_table.data_ptr()),
                                         int(n_peers), int(row_offset), int(ldy), ops._stream()), "gmr_rows_push_f32")
        ops._prof_end("rows_push", ev, bytes=4.0 * x.shape[0] * x.shape[1] * n_peers, peers=int(n_peers))
    ops.LAUNCHES += 1


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
    ``DiffMM._forward_mm_fused``, cut by output rows.  Every operand a later SpMM needs in full
    lives in a buffer replicated on all ranks and is filled by each rank storing its row block into every replica over
    NVLink peer memory -- from the SpMM epilogue itself where the block comes out of an SpMM (``spmm_push``), from a
    copy kernel otherwise (``rows_push``) -- followed by a stream-ordered barrier:

        [Z | Z + I0][i-block]   sharded projections + gmr_rows_normalize_mix      rows_push -> Xi [I, 2d]
        P_u, P_i                modality-graph terms lambda (w0 A_v + w1 A_t)[g] E0, K1, local, while Xi travels
        [Hz | modal_u]_g = [0 | P_u] + R_hat[g] . Xi                              K1, ONE 128-wide local pass
        xu_g      = U0_g + Hz_g             rows_push -> Xu [U, d];   modal_u,g rows_push -> modal [N, d]
        modal_i,g = P_i + R_hat'[g] . Xu    K1, local, while modal_u travels;     rows_push -> modal
        L_g       = A[g] . modal            K1, local (item rows first: the final item block travels during the user rows)
        E_g       = modal_g + L_g + ris_lambda n(modal_g);  item block rows_push -> final item table [I, d]

    and the result rows (users of the block, full item table) are returned with the block bounds.
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
        # replicated operands, each filled by every rank pushing its row block
        self.xi_all = PeerBuffer(ni, 2 * d, self.dev, group)  # [Z | Z + I0],  Z = w0 n(F_v) + w1 n(F_t)
        self.xu_all = PeerBuffer(nu, d, self.dev, group)      # U0 + R_hat Z
        self.modal_all = PeerBuffer(nu + ni, d, self.dev, group)   # modal, then each further layer output
        self.items_all = PeerBuffer(ni, d, self.dev, group)  # final item table
        self.token = torch.zeros(1, device=self.dev)      # barrier token of the compute stream
        self.token_c = torch.zeros(1, device=self.dev)    # barrier token of the communication stream
        self.comm = torch.cuda.Stream(device=self.dev)
        ug, ig = self.u1 - self.u0, self.i1 - self.i0
        mk = lambda r: torch.empty((r, d), dtype=torch.float32, device=self.dev)
        # row-block work buffers that cross streams (allocated once: the caching allocator never sees them freed)
        self.buf = {"xi_local": torch.empty((ig, 2 * d), dtype=torch.float32, device=self.dev),
                    "yu": torch.empty((ug, 2 * d), dtype=torch.float32, device=self.dev),
                    "modal_i": mk(ig), "xu_g": mk(ug), "acc_i": mk(ig)}
        self._adj_ref = None
        self._mix_sig = None

    def _graph_blocks(self, w0, w1):
        m = self.model
        nu = m.n_users
        full = m.norm_adj.full
        if self._adj_ref is not full:
            self.adj_u = full.row_block(self.u0, self.u1)
            self.adj_i = full.row_block(nu + self.i0, nu + self.i1)
            self._adj_ref = full
        mix = m._modal_mix_graph(m.image_UI_matrix, m.text_UI_matrix, w0, w1)
        sig = (w0, w1, m.ris_adj_lambda, id(mix))
        if self._mix_sig != sig:
            self.mix_u = mix.row_block(self.u0, self.u1)
            self.mix_i = mix.row_block(nu + self.i0, nu + self.i1)
            self._mix_sig = sig
            self._mix_ref = mix
        return self

    def _on_comm(self, after, fn):
        """Run ``fn`` on the communication stream once ``after`` (an event of the compute stream) has fired; returns
        an event of the communication stream that fires when fn's work -- including the cross-rank barrier it ends
        with, if any -- is done."""
        self.comm.wait_event(after)
        with torch.cuda.stream(self.comm):
            fn()
            ev = torch.cuda.Event()
            ev.record()
        return ev

    @staticmethod
    def _mark():
        ev = torch.cuda.Event()
        ev.record()
        return ev

    @torch.no_grad()
    def forward_MM(self, push_items=False):
        """Row blocks (users, items) of the propagated embeddings.  Peer stores run on a second stream and overlap
        the SpMMs that do not depend on them: the modality-graph terms are computed while Xi travels, modal_u travels
        during the item-row SpMM, the final item block during the user-row layer."""
        m = self.model
        nu, ni, d = m.n_users, m.n_items, m.latdim
        u0, u1, i0, i1 = self.u0, self.u1, self.i0, self.i1
        w0, w1 = m._modal_weights_host()
        self._graph_blocks(w0, w1)
        e0 = m._packed_e0()
        U0, I0 = e0[:nu], e0[nu:]
        W = self.world
        cur = torch.cuda.current_stream()
        stream_barrier(self.token)       # step fence: every rank has finished reading the replicated buffers
        # 1. sharded projections -> [Z | Z + I0] of the item block -> every replica (comm stream): the owner adds
        #    its I0 rows, so no rank ever runs a full-size elementwise pass
        pv = m._project(m.v_feat[i0:i1], m.image_trans.detach())
        pt = m._project(m.t_feat[i0:i1], m.text_trans.detach())
        xi_local = ops.rows_normalize_mix(pv, pt, w0, w1, y=I0[i0:i1], slope=m.leakyrelu.negative_slope,
                                          out=self.buf["xi_local"])

        def push_xi():
            rows_push(xi_local, self.xi_all.ptr_table, W, i0, 2 * d)
            stream_barrier(self.token_c)
        ev_z = self._on_comm(self._mark(), push_xi)
        #    meanwhile: the modality-graph terms (no remote data) seed the accumulators of the two big passes
        yu = self.buf["yu"]                                      # [U_g, 2d] = R_hat Z | modal_u
        yu[:, :d].zero_()
        ops.spmm_raw(self.mix_u, e0, out=yu[:, d:])
        modal_i = ops.spmm_raw(self.mix_i, e0, out=self.buf["modal_i"])
        # 2. user rows: ONE 128-wide pass  yu += R_hat[g] [Z | Z + I0];  xu_g = U0_g + R_hat Z -> every replica of xu
        cur.wait_event(ev_z)
        ops.spmm_raw(self.r_ui, self.xi_all.tensor, out=yu, beta=1.0)
        modal_u = yu[:, d:]
        xu_g = ops.rows_axpby_norm(U0[u0:u1], yu[:, :d], None, a=1.0, b=1.0, out=self.buf["xu_g"])

        def push_xu():
            rows_push(xu_g, self.xu_all.ptr_table, W, u0, d)
            stream_barrier(self.token_c)
        ev_hz = self._on_comm(self._mark(), push_xu)
        ev_mu = self._on_comm(self._mark(), lambda: rows_push(modal_u, self.modal_all.ptr_table, W, u0, d))
        # 3. item rows: modal_i += R_hat'[g] (U0 + R_hat Z)   (modal_u travels meanwhile)
        cur.wait_event(ev_hz)
        ops.spmm_raw(self.r_iu, self.xu_all.tensor, out=modal_i, beta=1.0)
        if m.gnn_layer == 0:
            cur.wait_event(ev_mu)
            acc_u = ops.rows_axpby_norm(modal_u, None, modal_u, a=1.0, c=m.ris_lambda)
            acc_i = ops.rows_axpby_norm(modal_i, None, modal_i, a=1.0, c=m.ris_lambda)
            return self._finish(acc_u, acc_i, push_items)

        def push_mi():
            rows_push(modal_i, self.modal_all.ptr_table, W, nu + i0, d)
            stream_barrier(self.token_c)
        ev_m = self._on_comm(self._mark(), push_mi)
        # 4. first GCN layer: item rows first, so that the final item block can travel during the user rows
        cur.wait_event(ev_m)
        full = self.modal_all.tensor
        last_i = ops.spmm_raw(self.adj_i, full)
        acc_i = ops.rows_axpby_norm(modal_i, last_i, modal_i, a=1.0, b=1.0, c=m.ris_lambda, out=self.buf["acc_i"])
        ev_items = None
        if m.gnn_layer == 1 and push_items:
            def push_items_fn():
                rows_push(acc_i, self.items_all.ptr_table, W, i0, d)
                stream_barrier(self.token_c)
            ev_items = self._on_comm(self._mark(), push_items_fn)
        last_u = ops.spmm_raw(self.adj_u, full)
        acc_u = ops.rows_axpby_norm(modal_u, last_u, modal_u, a=1.0, b=1.0, c=m.ris_lambda)
        # 5. further layers (not overlapped): gather the previous layer output, multiply, accumulate
        for _ in range(m.gnn_layer - 1):
            stream_barrier(self.token)   # every rank has read `full` before it is overwritten
            rows_push(last_u, self.modal_all.ptr_table, W, u0, d)
            rows_push(last_i, self.modal_all.ptr_table, W, nu + i0, d)
            stream_barrier(self.token)
            last_u, last_i = ops.spmm_raw(self.adj_u, full), ops.spmm_raw(self.adj_i, full)
            ops.rows_axpby_norm(acc_u, last_u, None, a=1.0, b=1.0, out=acc_u)
            ops.rows_axpby_norm(acc_i, last_i, None, a=1.0, b=1.0, out=acc_i)
        return self._finish(acc_u, acc_i, push_items, ev_items)

    def _finish(self, acc_u, acc_i, push_items, ev_items=None):
        cur = torch.cuda.current_stream()
        if push_items and ev_items is None:
            rows_push(acc_i, self.items_all.ptr_table, self.world, self.i0, self.model.latdim)
            stream_barrier(self.token)
        if ev_items is not None:
            cur.wait_event(ev_items)
        cur.wait_stream(self.comm)   # join: nothing of this call is left on the communication stream
        return acc_u, acc_i

    def close(self):
        """Release the peer-mapped buffers (CUDA IPC mappings of the other ranks' replicas)."""
        for buf in (self.xi_all, self.xu_all, self.modal_all, self.items_all):
            buf.close()

    @torch.no_grad()
    def eval_factors(self):
        """(user rows of this rank's block, full item table): every rank stores its item block into every replica."""
        out_u, _ = self.forward_MM(push_items=True)
        return out_u, self.items_all.tensor


class _LocalReplicas(object):
    """``world`` same-shaped buffers in ONE process standing in for the replicas of a PeerBuffer (single-GPU emulation of
    the ranks in the tests: every 'rank' sees the same pointer table)."""

    def __init__(self, rows, cols, device, world):
        self.tensors = [torch.zeros((int(rows), int(cols)), dtype=torch.float32, device=device) for _ in range(world)]
        self.ptr_table = torch.tensor([t.data_ptr() for t in self.tensors], dtype=torch.int64, device=device)

    def view(self, rank):
        class V(object):
            pass
        v = V()
        v.tensor, v.ptr_table, v.close = self.tensors[rank], self.ptr_table, (lambda: None)
        return v


class PeerBarrier(object):
    """Stream-ordered cross-rank barrier through flags in peer memory (``gmr_peer_barrier``): one 32-thread kernel per
    rank instead of a 1-element NCCL all-reduce.  ``check()`` (host, synchronising) raises if a peer ever failed to arrive."""

    def __init__(self, device, group=None, replicas=None, rank=None, world=None):
        if replicas is None:
            self.flags = PeerBuffer(1, 64, device, group)
            self.rank, self.world = self.flags.rank, self.flags.world
        else:
            self.flags, self.rank, self.world = replicas, rank, world
        self.state = torch.zeros(2, dtype=torch.int32, device=device)
        self.device = torch.device(device)

    def __call__(self):
        if self.world <= 1:
            return
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(lib.gmr_peer_barrier(C.c_void_p(self.flags.ptr_table.data_ptr()), self.rank, self.world,
                                            ops._ptr(self.state), ops._stream()), "gmr_peer_barrier")
        ops.LAUNCHES += 1

    def check(self):
        n = int(self.state[1].item())
        if n:
            raise RuntimeError("gmr_peer_barrier: %d wait(s) timed out -- a peer rank never arrived" % n)

    def close(self):
        self.flags.close()


def col_shard_supported(d, world):
    """Column sharding needs a power-of-two world whose column share is 8, 16 or 32 floats."""
    return world >= 2 and (world & (world - 1)) == 0 and d % world == 0 and (d // world) in (8, 16, 32)


class ColShardedDiffMM(object):
    """COLUMN-sharded ``forward_MM`` of a DiffMM replica: rank g owns columns [g dc, (g + 1) dc) of every embedding row.

    ``A X`` is column-separable, so every SpMM of the regrouped dataflow (``DiffMM._forward_mm_fused``) runs locally over
    the WHOLE graph on a [N, dc] operand that fits the L2 -- no all-gather of layer outputs (the row-sharded form moves
    ~0.9 GB per rank and step at the 1M x 500k shape and loses the gather reuse of the single-GPU pass).  What is
    exchanged, by peer stores (``gmr_cols_push_f32``) behind three barriers per step:

        P0  projections of the item block (row-sharded: the dense features stay where they are) -> [Z | Z + I0] of the
            block -> column slices to their owners (all-to-all, 2 x 4 dc bytes per item and peer); meanwhile
            e0_c = [U0; I0][:, cols], modal_c = lambda (w0 A_v + w1 A_t) e0_c
        P1  [R_hat Z | modal_u]_c += R_hat Xi_c (ONE 2dc-wide pass), xu_c = U0_c + R_hat Z, modal_i,c += R_hat' xu_c,
            L_c = A modal_c; this rank's part of |modal|^2 per row -> every rank (4 bytes per row and peer)
        P2  E_c = modal_c + L_c + ris_lambda modal_c / |modal|  (parts added as the single-GPU kernel's reduction tree);
            item columns -> every rank's item table, user columns -> the rank that scores those users

    The narrow SpMM keeps the wide kernels' per-column operation order and the norm parts combine like the 64-column
    row kernel's butterfly, so user block and item table equal the single-GPU result bit for bit."""

    def __init__(self, model, group=None, _emulate=None):
        import torch.distributed as dist

        self.model = m = model
        self.group = group
        self.dev = m.device
        nu, ni, d = m.n_users, m.n_items, m.latdim
        if _emulate is None:
            self.world = dist.get_world_size(group) if dist.is_initialized() else 1
            self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        else:
            self.rank, self.world = _emulate["rank"], _emulate["world"]
        W, g = self.world, self.rank
        if not col_shard_supported(d, W):
            raise ValueError("column sharding needs a power-of-two world with 8, 16 or 32 columns per rank (d=%d, world=%d)" % (d, W))
        self.dc = dc = d // W
        self.c0 = g * dc
        self.ub, self.ib = block_bounds(nu, W), block_bounds(ni, W)
        self.u0, self.u1, self.i0, self.i1 = self.ub[g], self.ub[g + 1], self.ib[g], self.ib[g + 1]
        n = nu + ni
        self.n_pad = (n + 3) // 4 * 4
        max_ub = max(self.ub[k + 1] - self.ub[k] for k in range(W))
        self.max_ub = max_ub
        # items_slab / users_slab: W column slabs [rows, dc], slab g written contiguously by rank g
        shapes = {"xi_c": (ni, 2 * dc), "ss_all": (W * self.n_pad // 4, 4), "items_slab": (W * ni, dc), "users_slab": (W * max_ub, dc)}
        if _emulate is None:
            self.peer = {k: PeerBuffer(r, c, self.dev, group) for k, (r, c) in shapes.items()}
            self.barrier = PeerBarrier(self.dev, group) if os.environ.get("GMR_PEER_BARRIER", "1") == "1" else None
        else:
            self.peer = {k: _emulate["replicas"][k].view(g) for k in shapes}
            self.barrier = None
        self._shapes = shapes
        self.token = torch.zeros(1, device=self.dev)
        mk = lambda r, c: torch.empty((r, c), dtype=torch.float32, device=self.dev)
        self.buf = {"xi_local": mk(self.i1 - self.i0, 2 * d), "e0_c": mk(n, dc), "yu_c": mk(nu, 2 * dc), "modal_c": mk(n, dc), "xu_c": mk(nu, dc),
                    "last_c": mk(n, dc), "emb_c": mk(n, dc), "ss": torch.zeros(self.n_pad, dtype=torch.float32, device=self.dev),
                    "items_all": mk(ni, d), "users_blk": mk(max_ub, d)}
        self.self_table = {k: torch.tensor([v.data_ptr()], dtype=torch.int64, device=self.dev) for k, v in self.buf.items()}
        self.ub_dev = torch.tensor(self.ub, dtype=torch.int64, device=self.dev)

    @classmethod
    def emulate(cls, model, world):
        """``world`` instances in one process over shared local replicas (tests): run with ``emulated_eval_factors``."""
        nu, ni, d = model.n_users, model.n_items, model.latdim
        dc = d // world
        n_pad = (nu + ni + 3) // 4 * 4
        ub = block_bounds(nu, world)
        max_ub = max(ub[k + 1] - ub[k] for k in range(world))
        shapes = {"xi_c": (ni, 2 * dc), "ss_all": (world * n_pad // 4, 4), "items_slab": (world * ni, dc), "users_slab": (world * max_ub, dc)}
        reps = {k: _LocalReplicas(r, c, model.device, world) for k, (r, c) in shapes.items()}
        return [cls(model, _emulate={"rank": g, "world": world, "replicas": reps}) for g in range(world)]

    @staticmethod
    def emulated_eval_factors(ranks):
        """Run the three phases of every emulated rank in lockstep; returns [(user block, item table)] per rank."""
        for fn in ("_p0", "_p1", "_p2", "_p3"):
            for r in ranks:
                getattr(r, fn)()
        return [r._result() for r in ranks]

    def _sync(self):
        if self.barrier is not None:
            self.barrier()
        else:
            stream_barrier(self.token)

    # ---- phases ------------------------------------------------------------------------------------------------
    def _p0(self):
        m = self.model
        nu, d, dc, W = m.n_users, m.latdim, self.dc, self.world
        i0, i1 = self.i0, self.i1
        w0, w1 = m._modal_weights_host()
        e0 = m._packed_e0()
        pv = m._project(m.v_feat[i0:i1], m.image_trans.detach())
        pt = m._project(m.t_feat[i0:i1], m.text_trans.detach())
        xi_local = ops.rows_normalize_mix(pv, pt, w0, w1, y=e0[nu + i0:nu + i1], slope=m.leakyrelu.negative_slope,
                                          out=self.buf["xi_local"])
        xi_tab = self.peer["xi_c"].ptr_table
        # all-to-all of column slices: peer h receives [Z[:, cols_h] | (Z + I0)[:, cols_h]] of this item block
        # (both halves in one launch: a destination row is the two slices side by side, so the remote stores are contiguous)
        ops.cols_push(xi_local, dc, xi_tab, W, 2 * dc, src_col0=0, src_col_step=dc, dst_row_offset=i0, dst_col0=0, tag="xi",
                      n_seg=2, src_seg_step=d, dst_seg_step=dc)
        # local work under the transfer: this rank's columns of [U0; I0], and the modality-graph term as the seed of modal_c
        e0_c = self.buf["e0_c"]
        ops.cols_push(e0, dc, self.self_table["e0_c"], 1, dc, src_col0=self.c0, tag="e0_local")
        mix = m._modal_mix_graph(m.image_UI_matrix, m.text_UI_matrix, w0, w1)
        ops.spmm_narrow(mix, e0_c, chains=2, out=self.buf["modal_c"])

    def _p1(self):
        m = self.model
        nu, dc = m.n_users, self.dc
        adj = m.norm_adj
        yu, modal, e0_c = self.buf["yu_c"], self.buf["modal_c"], self.buf["e0_c"]
        ops.spmm_narrow(adj.ui, self.peer["xi_c"].tensor, chains=1, out=yu)                    # [R_hat Z | R_hat (Z + I0)] (cols)
        xu = ops.rows_axpby_ss(e0_c[:nu], yu[:, :dc], a=1.0, b=1.0, out=self.buf["xu_c"])      # U0 + R_hat Z
        ops.rows_axpby_ss(modal[:nu], yu[:, dc:], a=1.0, b=1.0, out=modal[:nu])                # modal_u = mix term + R_hat (Z + I0)
        ops.spmm_narrow(adj.iu, xu, chains=2, out=modal[nu:], beta=1.0)                        # modal_i
        if m.gnn_layer >= 1:
            ops.spmm_narrow(adj.full, modal, chains=2, out=self.buf["last_c"])
        ss = ops.rows_sumsq(modal, out=self.buf["ss"])
        ops.cols_push(ss.view(-1, 4), 4, self.peer["ss_all"].ptr_table, self.world, 4, dst_row_offset=self.rank * (self.n_pad // 4),
                      tag="norms")

    def _p2(self):
        m = self.model
        nu, dc, W = m.n_users, self.dc, self.world
        modal, last, emb = self.buf["modal_c"], self.buf["last_c"], self.buf["emb_c"]
        ss_all = self.peer["ss_all"].tensor
        ops.rows_axpby_ss(modal, last if m.gnn_layer >= 1 else None, modal, ss_parts=ss_all, n_parts=W, ss_stride=self.n_pad,
                          a=1.0, b=1.0 if m.gnn_layer >= 1 else 0.0, c=m.ris_lambda, out=emb)
        for _ in range(max(0, m.gnn_layer - 1)):
            nxt = ops.spmm_narrow(m.norm_adj.full, last, chains=2)
            ops.rows_axpby_ss(emb, nxt, a=1.0, b=1.0, out=emb)
            last = nxt
        # column slabs, each stored contiguously into slab `rank` of every peer (items) / of the peer that scores the rows (users)
        ni = m.n_items
        ops.cols_push(emb[nu:], dc, self.peer["items_slab"].ptr_table, W, dc, dst_row_offset=self.rank * ni, tag="items")
        ops.cols_push(emb[:nu], dc, self.peer["users_slab"].ptr_table, W, dc, row_bounds=self.ub_dev,
                      dst_row_offset=self.rank * self.max_ub, tag="users")

    def _p3(self):
        """After the last barrier: the slabs the peers stored, side by side as 64-column rows (local)."""
        m, W, dc = self.model, self.world, self.dc
        ops.slabs_to_rows(self.peer["items_slab"].tensor, W, m.n_items, m.n_items, dc, self.buf["items_all"])
        ops.slabs_to_rows(self.peer["users_slab"].tensor, W, self.max_ub, self.u1 - self.u0, dc, self.buf["users_blk"])

    def _result(self):
        return self.buf["users_blk"][:self.u1 - self.u0], self.buf["items_all"]

    @torch.no_grad()
    def eval_factors(self):
        """(user rows of this rank's block, full item table), both 64 columns wide, ready for the fused scoring."""
        self._p0()
        self._sync()
        self._p1()
        self._sync()
        self._p2()
        self._sync()
        self._p3()
        return self._result()

    def forward_MM(self, push_items=False):
        ue, ie = self.eval_factors()
        return ue, ie[self.i0:self.i1]

    def close(self):
        for b in self.peer.values():
            b.close()
        if self.barrier is not None:
            self.barrier.close()


class ShardedGCNChain(object):
    """Row-sharded LightGCN-style propagation  c = sum_g w_g * mean_{l=0..L} G^l E0  over one or more N x N graphs:
    ``LightGCN.forward`` (GenMMRec/src/models/lightgcn.py:115-127, one graph) and the content embedding GenRecV1
    scores with (``user_item_GCN`` x 2, GenMMRec/src/models/genrecv1.py:255-264,335-341, the normalised adjacency and
    the generated-edge graph).  Rank g owns users [u0, u1) and items [i0, i1) of every graph; a layer output that the
    next layer needs in full is stored by every rank into every replica (``rows_push``) behind a stream-ordered
    barrier; with one layer (the shipped GenRecV1 / LightGCN settings use 1-4) the first product reads the replicated
    parameters directly.  The propagated item block is pushed into the replicated item table for scoring."""

    def __init__(self, model, graphs_fn, weights_fn, embeddings_fn, n_layers, group=None):
        """graphs_fn() -> list of GraphCSR [N, N]; weights_fn() -> list of floats; embeddings_fn() -> (U0, I0)."""
        import torch.distributed as dist

        self.model, self.group = model, group
        self.graphs_fn, self.weights_fn, self.embeddings_fn, self.n_layers = graphs_fn, weights_fn, embeddings_fn, n_layers
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dev = model.device
        nu, ni = model.n_users, model.n_items
        u0_, i0_ = embeddings_fn()
        self.d = int(u0_.shape[1])
        first = graphs_fn()[0]
        counts = (first.rowptr[1:] - first.rowptr[:-1]).to(torch.int64).cpu()
        rp_u = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(counts[:nu], 0)])
        rp_i = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(counts[nu:], 0)])
        self.ub, self.ib = nnz_balanced_bounds(rp_u, self.world), nnz_balanced_bounds(rp_i, self.world)
        g = self.rank
        self.u0, self.u1, self.i0, self.i1 = self.ub[g], self.ub[g + 1], self.ib[g], self.ib[g + 1]
        self.layer_all = PeerBuffer(nu + ni, self.d, self.dev, group)
        self.items_all = PeerBuffer(ni, self.d, self.dev, group)
        self.token = torch.zeros(1, device=self.dev)
        self._blocks = {}

    def _row_blocks(self, graph):
        key = id(graph)
        hit = self._blocks.get(key)
        if hit is None or hit[0] is not graph:
            nu = self.model.n_users
            hit = (graph, graph.row_block(self.u0, self.u1), graph.row_block(nu + self.i0, nu + self.i1))
            self._blocks[key] = hit
        return hit[1], hit[2]

    @torch.no_grad()
    def forward(self):
        nu, d, W = self.model.n_users, self.d, self.world
        U0, I0 = self.embeddings_fn()
        e0 = torch.cat([U0.detach(), I0.detach()])
        out_u = out_i = None
        weights = self.weights_fn()   # python floats, or a device tensor (no host read: the step is graph-capturable)
        for gi, graph in enumerate(self.graphs_fn()):
            w = weights[gi]
            g_u, g_i = self._row_blocks(graph)
            acc_u, acc_i = e0[self.u0:self.u1].clone(), e0[nu + self.i0:nu + self.i1].clone()
            full = e0
            for layer in range(self.n_layers):
                last_u, last_i = ops.spmm_raw(g_u, full), ops.spmm_raw(g_i, full)
                acc_u += last_u
                acc_i += last_i
                if layer + 1 < self.n_layers:
                    stream_barrier(self.token)   # every rank is done reading the previous layer
                    rows_push(last_u, self.layer_all.ptr_table, W, self.u0, d)
                    rows_push(last_i, self.layer_all.ptr_table, W, nu + self.i0, d)
                    stream_barrier(self.token)
                    full = self.layer_all.tensor
            if torch.is_tensor(w):   # same operation order as the single-GPU model: w_g * (sum / (L + 1))
                term_u, term_i = w * (acc_u / float(self.n_layers + 1)), w * (acc_i / float(self.n_layers + 1))
            else:
                scale = float(w) / float(self.n_layers + 1)
                term_u, term_i = acc_u * scale, acc_i * scale
            out_u = term_u if out_u is None else out_u + term_u
            out_i = term_i if out_i is None else out_i + term_i
            stream_barrier(self.token)           # the next graph's chain reuses the layer buffer
        return out_u, out_i

    @torch.no_grad()
    def eval_factors(self):
        out_u, out_i = self.forward()
        rows_push(out_i.contiguous(), self.items_all.ptr_table, self.world, self.i0, self.d)
        stream_barrier(self.token)
        return out_u, self.items_all.tensor

    def close(self):
        self.layer_all.close()
        self.items_all.close()


def sharded_genrecv1(model, group=None):
    """Row-sharded evaluation propagation of a GenRecV1 replica (its ``propagate()``: the content embedding)."""
    import torch.nn.functional as F

    def weights():
        return F.softmax(torch.stack([model.origin_weight.detach(), model.generation_weight.detach()]).flatten(), dim=0)

    from .models._common import as_graph
    return ShardedGCNChain(model, lambda: [as_graph(model.norm_adj), as_graph(model.image_UI_matrix)], weights,
                           lambda: (model.user_embedding.weight, model.item_id_embedding.weight), model.n_layers, group)


def sharded_lightgcn(model, group=None):
    """Row-sharded ``LightGCN.forward`` (mean of the layer outputs over the normalised adjacency)."""
    from .models._common import as_graph
    return ShardedGCNChain(model, lambda: [as_graph(model.norm_adj_matrix)], lambda: [1.0],
                           lambda: (model.embedding_dict["user_emb"], model.embedding_dict["item_emb"]), model.n_layers, group)


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


def evaluate_sharded(sharded, evaluator, shard, k=None, precision="tc", group=None):
    """One sharded evaluation pass, the multi-GPU form of ``Trainer.evaluate`` (common/trainer.py:369-388 of the
    reference): this rank's propagated factors (``sharded.eval_factors()`` of a ``ShardedDiffMM`` / ``ShardedGCNChain``)
    -> fused score + mask + top-K over its user block (``shard`` from ``shard_eval_by_user_block``) -> hit / metric SUMS
    -> all-reduce of the ``[4, K]`` float64 sums and the user count -> the reference's rounded dict, identical on every
    rank.  Returns ``(dict, unrounded [n_metrics, K], local top-K ids)``.  ``recall2`` needs the global hit matrix and
    is not available here."""
    import torch.distributed as dist

    k = int(k or max(evaluator.topk))
    ue, ie = sharded.eval_factors()
    ids, _ = ops.score_mask_topk(ue.contiguous(), ie.contiguous(), k, users=shard.eval_u, mask_rowptr=shard.mask_rowptr,
                                 mask_items=shard.mask_items, precision=precision, return_scores=False)
    sums, _ = evaluator.metric_sums(ids, shard)
    n = torch.tensor([float(ids.shape[0])], dtype=torch.float64, device=sums.device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, group=group)
        dist.all_reduce(n, group=group)
    out, raw = evaluator.finalize(sums, int(round(float(n.item()))))
    return out, raw, ids
