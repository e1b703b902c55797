"""Column-sharded propagation (dist.ColShardedDiffMM) on ONE GPU: the narrow SpMM against the wide kernel's columns, the
split row-norm against the 64-column row kernel, the column-slice pushes, and the whole dataflow with the ranks emulated
as separate buffers of one process -- everything must equal the single-GPU result BIT FOR BIT."""
import numpy as np
import pytest
import torch

from oracle import c_api

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from genmmrec_b200 import dist as gd, ops
    return gd, ops


def _graph(ops, rng, n_rows, n_cols, mean_deg, long_rows=(), dev="cuda:0"):
    deg = rng.poisson(mean_deg, size=n_rows)
    deg[rng.integers(0, n_rows, size=max(1, n_rows // 50))] = 0          # empty rows
    for r, n in long_rows:
        deg[r] = n                                                        # split rows (> 256 nonzeros)
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    col = rng.integers(0, n_cols, size=int(rowptr[-1])).astype(np.int32)
    val = rng.standard_normal(col.size).astype(np.float32)
    g = ops.GraphCSR(torch.from_numpy(rowptr).to(dev), torch.from_numpy(col).to(dev), torch.from_numpy(val).to(dev),
                     (n_rows, n_cols))
    return g, rowptr, col, val


@pytest.mark.parametrize("dc", [8, 16, 32])
@pytest.mark.parametrize("mean_deg,long_rows", [(3, ()), (40, ((5, 700), (77, 257), (300, 5000))), (130, ((0, 256),))])
def test_narrow_spmm_two_chains_equals_wide_columns(mods, dc, mean_deg, long_rows):
    gd, ops = mods
    rng = np.random.default_rng(dc * 1000 + mean_deg)
    n_rows, n_cols = 1531, 977
    g, rowptr, col, val = _graph(ops, rng, n_rows, n_cols, mean_deg, long_rows)
    x = torch.from_numpy(rng.standard_normal((n_cols, 64)).astype(np.float32)).cuda()
    wide = ops.spmm_raw(g, x)
    ref = c_api.spmm_csr_f64(rowptr, col, val, x.cpu().numpy())
    assert np.abs(wide.cpu().numpy() - ref).max() / np.abs(ref).max() < 1e-5
    for c0 in range(0, 64, dc):
        xc = x[:, c0:c0 + dc].contiguous()
        for sell in (False, True):                                        # CSR stream and sliced-ELL snapshot
            y = ops.spmm_narrow(g, xc, chains=2, sell=sell)
            assert torch.equal(y, wide[:, c0:c0 + dc]), (dc, c0, sell)
    # strided operand / output, alpha and beta
    y0 = torch.from_numpy(rng.standard_normal((n_rows, 64)).astype(np.float32)).cuda()
    wide2 = ops.spmm_raw(g, x, out=y0.clone(), alpha=0.5, beta=1.0)
    for sell in (False, True):
        buf = y0.clone()
        for c0 in range(0, 64, dc):
            ops.spmm_narrow(g, x[:, c0:c0 + dc], chains=2, out=buf[:, c0:c0 + dc], alpha=0.5, beta=1.0, sell=sell)
        assert torch.equal(buf, wide2), sell
    # in-place value change: the snapshot re-reads the values
    g.val.mul_(2.0)
    assert torch.equal(ops.spmm_narrow(g, x[:, :dc].contiguous(), chains=2, sell=True), ops.spmm_raw(g, x)[:, :dc])


@pytest.mark.parametrize("dc2", [16, 32, 64])
def test_narrow_spmm_one_chain_equals_128_wide_columns(mods, dc2):
    gd, ops = mods
    rng = np.random.default_rng(dc2)
    n_rows, n_cols = 1200, 800
    g, rowptr, col, val = _graph(ops, rng, n_rows, n_cols, 30, ((3, 900), (1100, 300)))
    x = torch.from_numpy(rng.standard_normal((n_cols, 128)).astype(np.float32)).cuda()
    wide = ops.spmm_raw(g, x)
    for c0 in range(0, 128, dc2):
        for sell in (False, True):
            y = ops.spmm_narrow(g, x[:, c0:c0 + dc2].contiguous(), chains=1, sell=sell)
            assert torch.equal(y, wide[:, c0:c0 + dc2]), (dc2, c0, sell)


def test_narrow_spmm_short_row_graph_and_nan_output(mods):
    """kNN-like graph (the wide side takes the short-row kernel) and beta = 0 over a NaN-filled output."""
    gd, ops = mods
    rng = np.random.default_rng(7)
    n = 3000
    g, rowptr, col, val = _graph(ops, rng, n, n, 5)
    x = torch.from_numpy(rng.standard_normal((n, 64)).astype(np.float32)).cuda()
    wide = ops.spmm_raw(g, x)
    for sell in (False, True):
        out = torch.full((n, 8), float("nan"), device="cuda")
        ops.spmm_narrow(g, x[:, 8:16].contiguous(), chains=2, out=out, sell=sell)
        assert torch.equal(out, wide[:, 8:16]), sell


@pytest.mark.parametrize("world", [2, 4, 8])
def test_split_row_norm_equals_row_kernel(mods, world):
    gd, ops = mods
    g = torch.Generator(device="cuda")
    g.manual_seed(world)
    n, d = 4099, 64
    dc = d // world
    x = torch.randn(n, d, device="cuda", generator=g)
    y = torch.randn(n, d, device="cuda", generator=g)
    z = torch.randn(n, d, device="cuda", generator=g) * torch.rand(n, 1, device="cuda", generator=g) * 10
    z[5] = 0.0                                                            # zero row: eps clamp
    ref = ops.rows_axpby_norm(x, y, z, a=1.0, b=1.0, c=0.3)
    n_pad = (n + 3) // 4 * 4
    parts = torch.zeros(world, n_pad, device="cuda")
    for k in range(world):
        ops.rows_sumsq(z[:, k * dc:(k + 1) * dc], out=parts[k])
    out = torch.empty_like(x)
    for k in range(world):
        sl = slice(k * dc, (k + 1) * dc)
        ops.rows_axpby_ss(x[:, sl], y[:, sl], z[:, sl], ss_parts=parts, n_parts=world, ss_stride=n_pad, a=1.0, b=1.0, c=0.3,
                          out=out[:, sl])
    assert torch.equal(out, ref)
    # no norm term: plain a x + b y
    assert torch.equal(ops.rows_axpby_ss(x[:, :dc], y[:, :dc], a=1.0, b=1.0), ops.rows_axpby_norm(x, y, None, a=1.0, b=1.0)[:, :dc])


def test_cols_push_all_to_all_gather_and_routed(mods):
    gd, ops = mods
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    world, dc, rows = 4, 8, 301
    src = torch.randn(rows, 2 * world * dc, device="cuda", generator=g)
    peers = [torch.full((1000, 2 * dc), -5.0, device="cuda") for _ in range(world)]
    table = torch.tensor([p.data_ptr() for p in peers], dtype=torch.int64, device="cuda")
    off = 123
    ops.cols_push(src, dc, table, world, 2 * dc, src_col0=0, src_col_step=dc, dst_row_offset=off, dst_col0=0)
    ops.cols_push(src, dc, table, world, 2 * dc, src_col0=world * dc, src_col_step=dc, dst_row_offset=off, dst_col0=dc)
    for p in range(world):
        assert torch.equal(peers[p][off:off + rows, :dc], src[:, p * dc:(p + 1) * dc])
        assert torch.equal(peers[p][off:off + rows, dc:], src[:, world * dc + p * dc:world * dc + (p + 1) * dc])
        assert (peers[p][:off] == -5).all() and (peers[p][off + rows:] == -5).all()
    # both halves in ONE launch (two segments): same result
    peers2 = [torch.full((1000, 2 * dc), -5.0, device="cuda") for _ in range(world)]
    table2 = torch.tensor([p.data_ptr() for p in peers2], dtype=torch.int64, device="cuda")
    ops.cols_push(src, dc, table2, world, 2 * dc, src_col0=0, src_col_step=dc, dst_row_offset=off, dst_col0=0, n_seg=2,
                  src_seg_step=world * dc, dst_seg_step=dc)
    for p in range(world):
        assert torch.equal(peers2[p], peers[p])
    # all-gather of one column slice into column block 2 of every peer
    wide = [torch.full((rows, world * dc), -1.0, device="cuda") for _ in range(world)]
    wt = torch.tensor([p.data_ptr() for p in wide], dtype=torch.int64, device="cuda")
    ops.cols_push(src[:, 16:16 + dc], dc, wt, world, world * dc, dst_col0=2 * dc)
    for p in range(world):
        assert torch.equal(wide[p][:, 2 * dc:3 * dc], src[:, 16:16 + dc]) and (wide[p][:, :2 * dc] == -1).all()
    # routed by row range
    bounds = [0, 50, 50, 200, rows]
    blk = [torch.full((200, world * dc), -2.0, device="cuda") for _ in range(world)]
    bt = torch.tensor([p.data_ptr() for p in blk], dtype=torch.int64, device="cuda")
    ops.cols_push(src[:, :dc], dc, bt, world, world * dc, row_bounds=torch.tensor(bounds, dtype=torch.int64, device="cuda"),
                  dst_col0=dc)
    for p in range(world):
        n = bounds[p + 1] - bounds[p]
        assert torch.equal(blk[p][:n, dc:2 * dc], src[bounds[p]:bounds[p + 1], :dc])
        assert (blk[p][n:] == -2).all() and (blk[p][:, :dc] == -2).all()


def test_slabs_to_rows(mods):
    gd, ops = mods
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    n_slabs, slab_rows, n_rows, dc = 8, 700, 613, 8
    slabs = torch.randn(n_slabs * slab_rows, dc, device="cuda", generator=g)
    out = torch.full((n_rows, n_slabs * dc + 8), -9.0, device="cuda")
    ops.slabs_to_rows(slabs, n_slabs, slab_rows, n_rows, dc, out)
    for p in range(n_slabs):
        assert torch.equal(out[:, p * dc:(p + 1) * dc], slabs[p * slab_rows:p * slab_rows + n_rows])
    assert (out[:, n_slabs * dc:] == -9).all()


@pytest.mark.parametrize("world,layers", [(2, 1), (4, 2), (8, 1), (8, 0)])
def test_col_sharded_diffmm_equals_single_gpu(mods, world, layers):
    """The whole column-sharded dataflow with the ranks emulated in one process: user blocks and item table bit-identical
    to the single-GPU propagation (Baby shape)."""
    gd, ops = mods
    from genmmrec_b200.workload import Workload

    wl = Workload("DiffMM", "baby", torch.device("cuda:0"), overrides={"n_layers": layers})
    model = wl.model
    with torch.no_grad():
        ue, ie = model.propagate()
        ranks = gd.ColShardedDiffMM.emulate(model, world)
        for _ in range(2):                                                # buffers are reused across steps
            res = gd.ColShardedDiffMM.emulated_eval_factors(ranks)
    torch.cuda.synchronize()
    for r, (su, items) in zip(ranks, res):
        assert torch.equal(items, ie), (world, r.rank)
        assert torch.equal(su, ue[r.u0:r.u1]), (world, r.rank)


def test_peer_barrier_single_process(mods):
    """Two 'ranks' of one process on two streams meet at the flag barrier."""
    gd, ops = mods
    dev = torch.device("cuda:0")
    reps = gd._LocalReplicas(1, 64, dev, 2)
    b = [gd.PeerBarrier(dev, replicas=reps.view(k), rank=k, world=2) for k in range(2)]
    s = [torch.cuda.Stream(), torch.cuda.Stream()]
    for _ in range(3):
        for k in range(2):
            with torch.cuda.stream(s[k]):
                b[k]()
    torch.cuda.synchronize()
    for k in range(2):
        b[k].check()
        assert int(b[k].state[0]) == 3
