"""The column-sharded DiffMM propagation (dist.ColShardedDiffMM) as a dataflow, under real
world_size-2 and -4 process groups (gloo, CPU): every rank owns d / world embedding columns of every row, runs the
phases P0 .. P2 on its slice with plain float64 torch ops, exchanges exactly what the GPU path exchanges (column slices
of [Z | Z + I0], per-row partial square sums, final column slabs), and must reproduce the embeddings the UNMODIFIED
reference produced for the toy fixture (tests/golden/toy_diffmm.npz: `emb/user`, `emb/item`).  No kernel runs here: this
pins the algebra (column separability of A X, the regrouped forward_MM, the split row norm) and the block bookkeeping
the CUDA path shares (block_bounds, col_shard_supported)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from conftest import REPO


def _dense(z, name, n):
    idx, val = z["graph/%s/indices" % name], z["graph/%s/values" % name]
    m = torch.zeros((n, n), dtype=torch.float64)
    m.index_put_((torch.from_numpy(idx[0]), torch.from_numpy(idx[1])), torch.from_numpy(val).double(), accumulate=True)
    return m


def _worker(rank, world, port, out):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import load_golden, toy_arrays
        from genmmrec_b200 import dist as gd

        z, meta = load_golden("toy_diffmm")
        cfg = meta["config"]
        data = toy_arrays()
        nu, ni = data["n_users"], data["n_items"]
        n = nu + ni
        d = z["param/uEmbeds"].shape[1]
        assert gd.col_shard_supported(d, world) and not gd.col_shard_supported(d, 3) and not gd.col_shard_supported(d, 1)
        dc = d // world
        cols = slice(rank * dc, (rank + 1) * dc)
        ub, ib = gd.block_bounds(nu, world), gd.block_bounds(ni, world)
        i0_, i1_ = ib[rank], ib[rank + 1]

        t64 = lambda a: torch.from_numpy(np.asarray(a)).double()
        u0, it0 = t64(z["param/uEmbeds"]), t64(z["param/iEmbeds"])
        w = torch.softmax(t64(z["param/modal_weight"]), dim=0)
        adj = _dense(z, "norm_adj", n)
        r_hat, r_hat_t = adj[:nu, nu:], adj[nu:, :nu]
        mix = cfg["ris_adj_lambda"] * (w[0] * _dense(z, "image_UI_matrix", n) + w[1] * _dense(z, "text_UI_matrix", n))

        # ---- P0: projections of this rank's ITEM block (row-sharded), [Z | Z + I0] of the block ----
        pv = F.leaky_relu(t64(data["img"])[i0_:i1_] @ t64(z["param/image_trans"]), 0.2)
        pt = F.leaky_relu(t64(data["txt"])[i0_:i1_] @ t64(z["param/text_trans"]), 0.2)
        z_blk = w[0] * F.normalize(pv) + w[1] * F.normalize(pt)
        xi_blk = torch.cat([z_blk, z_blk + it0[i0_:i1_]], dim=1)                   # [block, 2 d]
        # exchange: every rank ends up with ITS columns of both halves for all items (the GPU path stores the slices
        # into the owners' replicas; here the blocks are gathered and sliced)
        xi_all = gd.all_gather_rows(xi_blk, [ib[g + 1] - ib[g] for g in range(world)])
        xi_c = torch.cat([xi_all[:, cols], xi_all[:, d + rank * dc:d + (rank + 1) * dc]], dim=1)   # [ni, 2 dc]
        e0_c = torch.cat([u0, it0])[:, cols]
        modal_c = mix @ e0_c
        # ---- P1: every SpMM over the WHOLE graph on the rank's columns ----
        yu = r_hat @ xi_c                                                           # [R Z | R (Z + I0)] (cols)
        xu = u0[:, cols] + yu[:, :dc]
        modal_c[:nu] += yu[:, dc:]
        modal_c[nu:] += r_hat_t @ xu
        layers = [modal_c]
        for _ in range(cfg["n_layers"]):
            layers.append(adj @ layers[-1])
        ss = (modal_c ** 2).sum(dim=1)                                              # this rank's share of |modal|^2
        dist.all_reduce(ss)                                                         # (GPU path: parts pushed, added in order)
        # ---- P2: E_c, then the column slabs to their consumers ----
        emb_c = sum(layers) + cfg["ris_lambda"] * modal_c / ss.sqrt().clamp_min(1e-12).unsqueeze(1)
        slabs = [torch.empty_like(emb_c) for _ in range(world)]
        dist.all_gather(slabs, emb_c.contiguous())
        emb = torch.cat(slabs, dim=1)                                               # slab g = columns of rank g
        users_blk, items = emb[ub[rank]:ub[rank + 1]], emb[nu:]
        ref_u, ref_i = t64(z["emb/user"]), t64(z["emb/item"])
        scale = float(max(ref_u.abs().max(), ref_i.abs().max()))
        err_u = float((users_blk - ref_u[ub[rank]:ub[rank + 1]]).abs().max()) / scale
        err_i = float((items - ref_i).abs().max()) / scale
        assert err_u < 1e-5 and err_i < 1e-5, (err_u, err_i)
        out.put((rank, "ok"))
    except Exception:  # surface the failure in the parent
        import traceback
        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_column_sharded_dataflow_matches_reference(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + (os.getpid() + 11 * world) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", "rank %d failed:\n%s" % (rank, msg)
