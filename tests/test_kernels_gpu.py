"""Parity of the CUDA kernels (through the C ABI) against the C oracle on seeded inputs, including
the edge cases the domain has: empty rows, duplicate entries, rows longer than the chunk (split
rows), ragged widths, masks that leave fewer than K items, ties, K at the limits."""
import numpy as np
import pytest
import torch

from oracle import c_api

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from genmmrec_b200 import ops as _ops
    return _ops


def random_csr(rng, n_rows, n_cols, avg, long_rows=(), empty_rows=()):
    deg = rng.poisson(avg, size=n_rows).astype(np.int64)
    for r, n in long_rows:
        deg[r] = n
    for r in empty_rows:
        deg[r] = 0
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    col = rng.integers(0, n_cols, size=int(rowptr[-1])).astype(np.int32)  # duplicates allowed
    val = rng.standard_normal(col.size).astype(np.float32)
    return rowptr, col, val


def to_graph(ops, rowptr, col, val, shape, chunk=0):
    dev = torch.device("cuda:0")
    return ops.GraphCSR(torch.from_numpy(rowptr).to(dev), torch.from_numpy(col).to(dev), torch.from_numpy(val).to(dev),
                        shape, chunk_nnz=chunk)


def rel(a, b):
    return np.abs(a.astype(np.float64) - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("d", [64, 128, 192, 32, 20, 7, 1, 516])
def test_spmm_matches_oracle(ops, d):
    rng = np.random.default_rng(d)
    n_rows, n_cols = 3000, 2500
    rowptr, col, val = random_csr(rng, n_rows, n_cols, 20, long_rows=[(7, 5000), (1500, 700), (2999, 257)],
                                  empty_rows=[0, 11, 2998])
    x = rng.standard_normal((n_cols, d)).astype(np.float32)
    g = to_graph(ops, rowptr, col, val, (n_rows, n_cols))
    assert g.plan_stats()["split_rows"] == 3
    y = ops.spmm_raw(g, torch.from_numpy(x).cuda()).cpu().numpy()
    y64 = c_api.spmm_csr_f64(rowptr, col, val, x)
    y32 = c_api.spmm_csr(rowptr, col, val, x)
    # tolerance of the north star: 1e-5 relative (fp32), per matrix
    assert rel(y, y64) < 1e-5
    assert rel(y32, y64) < 1e-5
    assert np.all(y[0] == 0) and np.all(y[11] == 0)
    # determinism: split rows are reduced in a fixed order
    y2 = ops.spmm_raw(g, torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(y, y2)


def test_spmm_alpha_beta_and_strided_views(ops):
    rng = np.random.default_rng(3)
    n_rows, n_cols, d = 1000, 800, 64
    rowptr, col, val = random_csr(rng, n_rows, n_cols, 9, long_rows=[(5, 900)])
    g = to_graph(ops, rowptr, col, val, (n_rows, n_cols))
    xw = torch.from_numpy(rng.standard_normal((n_cols, 192)).astype(np.float32)).cuda()
    yw = torch.from_numpy(rng.standard_normal((n_rows, 192)).astype(np.float32)).cuda()
    y0 = yw.clone()
    x = xw[:, 64:128]
    ops.spmm_raw(g, x, out=yw[:, 128:192], alpha=0.5, beta=2.0)
    ref = 0.5 * c_api.spmm_csr_f64(rowptr, col, val, x.cpu().numpy()) + 2.0 * y0[:, 128:192].cpu().numpy()
    assert rel(yw[:, 128:192].cpu().numpy(), ref) < 1e-5
    assert torch.equal(yw[:, :128], y0[:, :128])  # neighbours of the slice untouched


def test_spmm_chunk_sizes_agree(ops):
    rng = np.random.default_rng(4)
    rowptr, col, val = random_csr(rng, 500, 400, 40, long_rows=[(3, 3000)])
    x = torch.from_numpy(rng.standard_normal((400, 64)).astype(np.float32)).cuda()
    ref = c_api.spmm_csr_f64(rowptr, col, val, x.cpu().numpy())
    for chunk in (32, 64, 256, 4096):
        g = to_graph(ops, rowptr, col, val, (500, 400), chunk=chunk)
        assert rel(ops.spmm_raw(g, x).cpu().numpy(), ref) < 1e-5


@pytest.mark.parametrize("shape", ["knn", "ragged_split", "tiny"])
def test_spmm_short_row_kernel_is_bit_identical(ops, shape, monkeypatch):
    """The half-warp-per-row variant (kNN / modal-mix graphs, D = 64) sums in the general kernel's order: forcing either
    kernel through GMR_SPMM_SHORT must give identical bits, for whole rows, split rows, empty rows, rows past one
    16-entry register stage, an odd row count (idle half-warp) and the alpha/beta/strided epilogue."""
    rng = np.random.default_rng(17)
    if shape == "knn":  # fixed 10 neighbours per row, identity plan
        n_rows = n_cols = 4001
        rowptr = (np.arange(n_rows + 1) * 10).astype(np.int32)
        col = rng.integers(0, n_cols, size=n_rows * 10).astype(np.int32)
        val = rng.standard_normal(col.size).astype(np.float32)
    elif shape == "ragged_split":
        n_rows, n_cols = 2777, 1900
        rowptr, col, val = random_csr(rng, n_rows, n_cols, 5, long_rows=[(3, 900), (100, 17), (101, 33), (2776, 300)],
                                      empty_rows=[0, 1, 50, 2775])
    else:
        n_rows, n_cols = 3, 5
        rowptr, col, val = random_csr(rng, n_rows, n_cols, 2, empty_rows=[1])
    g = to_graph(ops, rowptr, col, val, (n_rows, n_cols))
    xw = torch.from_numpy(rng.standard_normal((n_cols, 128)).astype(np.float32)).cuda()
    y0 = torch.from_numpy(rng.standard_normal((n_rows, 192)).astype(np.float32)).cuda()
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("GMR_SPMM_SHORT", mode)
        plain = ops.spmm_raw(g, xw[:, :64].contiguous())
        yw = y0.clone()
        ops.spmm_raw(g, xw[:, 64:], out=yw[:, 64:128], alpha=0.25, beta=-1.5)
        out[mode] = (plain, yw)
    assert torch.equal(out["0"][0], out["1"][0])
    assert torch.equal(out["0"][1], out["1"][1])
    ref = c_api.spmm_csr_f64(rowptr, col, val, xw[:, :64].cpu().numpy())
    assert rel(out["1"][0].cpu().numpy(), ref) < 1e-5
    assert torch.equal(out["1"][1][:, :64], y0[:, :64]) and torch.equal(out["1"][1][:, 128:], y0[:, 128:])
    monkeypatch.delenv("GMR_SPMM_SHORT")
    auto = ops.spmm_raw(g, xw[:, :64].contiguous())  # the heuristic's pick, whichever it is
    assert torch.equal(auto, out["0"][0])


def test_spmm_backward_is_transpose(ops):
    rng = np.random.default_rng(5)
    rowptr, col, val = random_csr(rng, 300, 200, 6)
    g = to_graph(ops, rowptr, col, val, (300, 200))
    x = torch.from_numpy(rng.standard_normal((200, 64)).astype(np.float32)).cuda().requires_grad_(True)
    w = torch.from_numpy(rng.standard_normal((300, 64)).astype(np.float32)).cuda()
    (ops.spmm(g, x) * w).sum().backward()
    dense = g.to_torch_coo().to_dense().double()
    assert rel(x.grad.cpu().numpy(), (dense.T @ w.double()).cpu().numpy()) < 1e-5


def make_mask(rng, b, n_items, avg):
    lens = rng.poisson(avg, size=b)
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    items = np.concatenate([np.sort(rng.choice(n_items, size=n, replace=False)) for n in lens] + [np.zeros(0, np.int64)])
    return rowptr, items.astype(np.int32)


@pytest.mark.parametrize("n_items,d,k,with_bias", [(7050, 64, 50, False), (1000, 128, 20, False), (333, 256, 64, True),
                                                   (130, 64, 128, False), (5000, 60, 5, True), (900, 64, 200, False)])
def test_score_mask_topk_bit_exact(ops, n_items, d, k, with_bias):
    rng = np.random.default_rng(n_items + d)
    n_users, b = 700, 389
    eu = rng.standard_normal((n_users, d)).astype(np.float32)
    ei = rng.standard_normal((n_items, d)).astype(np.float32)
    ei[17] = ei[3]  # exact score ties between two items -> id order decides
    users = rng.choice(n_users, size=b, replace=False).astype(np.int64)
    bias = rng.standard_normal(n_items).astype(np.float32) if with_bias else None
    mrp, mit = make_mask(rng, b, n_items, 30)
    ids_ref, sc_ref = c_api.score_mask_topk(eu, users, ei, bias, mrp, mit, k)
    ids, sc = ops.score_mask_topk(torch.from_numpy(eu).cuda(), torch.from_numpy(ei).cuda(), k,
                                  users=torch.from_numpy(users).cuda(),
                                  bias=None if bias is None else torch.from_numpy(bias).cuda(),
                                  mask_rowptr=torch.from_numpy(mrp).cuda(), mask_items=torch.from_numpy(mit).cuda())
    assert np.array_equal(ids.cpu().numpy(), ids_ref)
    assert np.array_equal(sc.cpu().numpy(), sc_ref)


def test_score_topk_fewer_unmasked_than_k(ops):
    """trainer.py:384 writes the finite -1e10: masked items surface once the unmasked run out."""
    rng = np.random.default_rng(9)
    n_items, d, k = 40, 64, 32
    eu = rng.standard_normal((5, d)).astype(np.float32)
    ei = rng.standard_normal((n_items, d)).astype(np.float32)
    lens = np.array([0, 10, 20, 39, 40])
    mrp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    mit = np.concatenate([np.sort(rng.choice(n_items, size=n, replace=False)) for n in lens]).astype(np.int32)
    ids_ref, sc_ref = c_api.score_mask_topk(eu, None, ei, None, mrp, mit, k)
    ids, sc = ops.score_mask_topk(torch.from_numpy(eu).cuda(), torch.from_numpy(ei).cuda(), k,
                                  mask_rowptr=torch.from_numpy(mrp).cuda(), mask_items=torch.from_numpy(mit).cuda())
    assert np.array_equal(ids.cpu().numpy(), ids_ref)
    assert np.array_equal(sc.cpu().numpy(), sc_ref)
    assert (sc.cpu().numpy()[4] == -1e10).all()


def test_score_topk_k_larger_than_items(ops):
    rng = np.random.default_rng(10)
    eu = rng.standard_normal((3, 64)).astype(np.float32)
    ei = rng.standard_normal((10, 64)).astype(np.float32)
    ids_ref, sc_ref = c_api.score_mask_topk(eu, None, ei, None, None, None, 16)
    ids, sc = ops.score_mask_topk(torch.from_numpy(eu).cuda(), torch.from_numpy(ei).cuda(), 16)
    assert np.array_equal(ids.cpu().numpy(), ids_ref)
    assert (ids.cpu().numpy()[:, 10:] == -1).all() and np.isneginf(sc.cpu().numpy()[:, 10:]).all()


def test_score_topk_adversarial_order(ops):
    """Scores increasing with the item id: every tile beats the running threshold (worst case for
    the append buffers, exercises the overflow/retry path)."""
    n_items, d, k = 6000, 64, 50
    eu = np.zeros((130, d), dtype=np.float32)
    eu[:, 0] = 1.0
    ei = np.zeros((n_items, d), dtype=np.float32)
    ei[:, 0] = np.arange(n_items, dtype=np.float32)
    ids, sc = ops.score_mask_topk(torch.from_numpy(eu).cuda(), torch.from_numpy(ei).cuda(), k)
    want = np.arange(n_items - 1, n_items - 1 - k, -1, dtype=np.int32)
    assert (ids.cpu().numpy() == want[None, :]).all()


@pytest.mark.parametrize("k", [50, 20, 1, 64, 200])
def test_hits_metrics_match_oracle(ops, k):
    rng = np.random.default_rng(k)
    u, n_items = 1000, 400
    topk = np.stack([rng.choice(n_items, size=k, replace=False) for _ in range(u)]).astype(np.int32)
    lens = rng.integers(1, 80, size=u)
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    items = np.concatenate([np.sort(rng.choice(n_items, size=n, replace=False)) for n in lens]).astype(np.int32)
    hit_ref = c_api.hits(topk, rowptr, items)
    m_ref = c_api.metrics(hit_ref, lens)
    sums, hit = ops.hits_metrics(torch.from_numpy(topk).cuda(), torch.from_numpy(rowptr).cuda(),
                                 torch.from_numpy(items).cuda(), return_hit=True)
    assert np.array_equal(hit.cpu().numpy(), hit_ref)  # integer work: bit-exact
    got = (sums / u).cpu().numpy()
    for j, m in enumerate(("recall", "ndcg", "precision", "map")):
        assert np.abs(got[j] - m_ref[m]).max() < 1e-12, m


def test_rows_axpby_norm(ops):
    rng = np.random.default_rng(12)
    x, y, z = (torch.from_numpy(rng.standard_normal((500, 64)).astype(np.float32)).cuda() for _ in range(3))
    out = ops.rows_axpby_norm(x, y, z, a=1.0, b=0.3, c=0.5)
    ref = x + 0.3 * y + 0.5 * torch.nn.functional.normalize(z)
    assert rel(out.cpu().numpy(), ref.double().cpu().numpy()) < 1e-6
    # D = 64 vector path: odd row count, x a column slice of a wider buffer, z aliasing x, y absent, in-place output
    wide = torch.from_numpy(rng.standard_normal((501, 128)).astype(np.float32)).cuda()
    xs = wide[:, 64:]
    ref2 = xs + 0.25 * torch.nn.functional.normalize(xs)
    out2 = ops.rows_axpby_norm(xs, None, xs, a=1.0, c=0.25)
    assert rel(out2.cpu().numpy(), ref2.double().cpu().numpy()) < 1e-6
    y2 = torch.from_numpy(rng.standard_normal((501, 64)).astype(np.float32)).cuda()
    ref3 = (y2 + 2.0 * xs).clone()
    ops.rows_axpby_norm(y2, xs, None, a=1.0, b=2.0, out=y2)
    assert rel(y2.cpu().numpy(), ref3.double().cpu().numpy()) < 1e-6
    # generic path (D = 96)
    x9, z9 = (torch.from_numpy(rng.standard_normal((77, 96)).astype(np.float32)).cuda() for _ in range(2))
    out9 = ops.rows_axpby_norm(x9, None, z9, a=0.5, c=1.0)
    assert rel(out9.cpu().numpy(), (0.5 * x9 + torch.nn.functional.normalize(z9)).double().cpu().numpy()) < 1e-6


def test_errors_are_loud(ops):
    rng = np.random.default_rng(13)
    rowptr, col, val = random_csr(rng, 10, 10, 3)
    g = to_graph(ops, rowptr, col, val, (10, 10))
    with pytest.raises(ValueError):
        ops.spmm_raw(g, torch.zeros((11, 64), device="cuda"))
    with pytest.raises(RuntimeError):
        ops.score_mask_topk(torch.zeros((4, 64), device="cuda"), torch.zeros((9, 64), device="cuda"), 1000)


# ---- tensor-core scoring paths: must equal the fp32 path bit for bit -----------------------------------
#   "tc"        one fp16 tcgen05 pass as a certified screen + exact fp32 re-score (score_topk_screen.cu)
#   "tc_split"  3-term split-bf16 tcgen05 product + certified exact re-rank       (score_topk_tc.cu)

TC_MODES = ["tc", "tc_split"]


def _tc_case(ops, rng, n_users, b, n_items, d, k, with_bias, mask_avg, scale=1.0, precision="tc"):
    eu = (rng.standard_normal((n_users, d)) * scale).astype(np.float32)
    ei = (rng.standard_normal((n_items, d)) * scale).astype(np.float32)
    users = rng.choice(n_users, size=b, replace=False).astype(np.int64)
    bias = rng.standard_normal(n_items).astype(np.float32) if with_bias else None
    mrp, mit = make_mask(rng, b, n_items, mask_avg)
    args = dict(users=torch.from_numpy(users).cuda(), bias=None if bias is None else torch.from_numpy(bias).cuda(),
                mask_rowptr=torch.from_numpy(mrp).cuda(), mask_items=torch.from_numpy(mit).cuda())
    eu_t, ei_t = torch.from_numpy(eu).cuda(), torch.from_numpy(ei).cuda()
    ids_f, sc_f = ops.score_mask_topk(eu_t, ei_t, k, precision="fp32", **args)
    ids_t, sc_t = ops.score_mask_topk(eu_t, ei_t, k, precision=precision, **args)
    torch.cuda.synchronize()
    return ids_f.cpu().numpy(), sc_f.cpu().numpy(), ids_t.cpu().numpy(), sc_t.cpu().numpy(), ops.last_tc_fallback_rows()


@pytest.mark.parametrize("n_items,d,k,with_bias", [(7050, 64, 50, False), (1000, 64, 20, False), (333, 64, 5, True),
                                                   (5000, 128, 50, False), (2100, 192, 50, True), (130, 64, 50, False),
                                                   (4000, 64, 100, False), (3000, 128, 10, True)])
@pytest.mark.parametrize("precision", TC_MODES)
def test_score_tc_equals_fp32(ops, n_items, d, k, with_bias, precision):
    rng = np.random.default_rng(1000 + n_items + d)
    ids_f, sc_f, ids_t, sc_t, fb = _tc_case(ops, rng, 900, 517, n_items, d, k, with_bias, 25, precision=precision)
    assert np.array_equal(ids_t, ids_f)
    assert np.array_equal(sc_t, sc_f)
    assert fb <= 0.02 * 517  # certification almost never fails on continuous random scores


@pytest.mark.parametrize("n_users,b,n_items,d,k,with_bias", [(300, 300, 3000, 256, 50, True), (2000, 1300, 9000, 64, 200, False),
                                                             (64, 7, 100, 64, 50, False), (700, 600, 20000, 64, 256, False),
                                                             (1024, 1024, 4096, 128, 64, False)])
def test_score_screen_shapes(ops, n_users, b, n_items, d, k, with_bias):
    """Shapes only the screen path takes: D = 256, K up to 256, fewer users than one CTA group, fewer items
    than one tile, exact multiples of the tile sizes."""
    rng = np.random.default_rng(4000 + n_items + d + k)
    ids_f, sc_f, ids_t, sc_t, fb = _tc_case(ops, rng, n_users, b, n_items, d, k, with_bias, 10, precision="tc")
    assert np.array_equal(ids_t, ids_f)
    assert np.array_equal(sc_t, sc_f)


def test_score_screen_all_equal_items_overflow(ops):
    """Every item identical: the screen cannot separate anything, every row overflows its candidate buffer and
    is redone on the fp32 path; order is then decided by the item id alone."""
    rng = np.random.default_rng(8)
    n_items, d, k = 3000, 64, 50
    ei = np.repeat(rng.standard_normal((1, d)).astype(np.float32), n_items, axis=0)
    eu = rng.standard_normal((300, d)).astype(np.float32)
    ids_ref, sc_ref = c_api.score_mask_topk(eu, None, ei, None, None, None, k)
    ids, sc = ops.score_mask_topk(torch.from_numpy(eu).cuda(), torch.from_numpy(ei).cuda(), k, precision="tc")
    fb = ops.last_tc_fallback_rows()
    assert np.array_equal(ids.cpu().numpy(), ids_ref)
    assert np.array_equal(sc.cpu().numpy(), sc_ref)
    assert fb == 300


def test_score_screen_zero_and_tiny_rows(ops):
    """All-zero user rows (every score ties at 0), subnormal-scale rows and items next to ordinary ones."""
    rng = np.random.default_rng(9)
    n_items, d, k = 2500, 64, 20
    eu = rng.standard_normal((260, d)).astype(np.float32)
    eu[3] = 0.0
    eu[77] *= 1e-30
    eu[78] *= 1e20
    ei = rng.standard_normal((n_items, d)).astype(np.float32)
    ei[5] = 0.0
    ei[6] *= 1e-12
    ids_ref, sc_ref = c_api.score_mask_topk(eu, None, ei, None, None, None, k)
    ids, sc = ops.score_mask_topk(torch.from_numpy(eu).cuda(), torch.from_numpy(ei).cuda(), k, precision="tc")
    assert np.array_equal(ids.cpu().numpy(), ids_ref)
    assert np.array_equal(sc.cpu().numpy(), sc_ref)


def test_score_screen_skewed_norms_early_stop(ops, monkeypatch):
    """Popularity-skewed item norms: the norm-ordered sweep must stop long before the catalogue ends (exact
    Cauchy-Schwarz pruning) and still return the oracle's result; train histories are drawn from the
    high-norm items so that some rows need more tiles than others."""
    monkeypatch.setenv("GMR_SCREEN_STATS", "1")
    rng = np.random.default_rng(11)
    n_items, d, k, b = 30000, 64, 50, 700
    scale = np.exp(rng.normal(0.0, 1.2, size=(n_items, 1)))
    ei = (rng.standard_normal((n_items, d)) * scale).astype(np.float32)
    eu = rng.standard_normal((b, d)).astype(np.float32)
    hot = np.argsort(-scale[:, 0])[:600]
    lens = rng.integers(0, 120, size=b)
    lens[::50] = 550                                  # a few users have seen almost every popular item
    rows = [np.sort(rng.choice(hot, size=n, replace=False)).astype(np.int32) for n in lens]
    mrp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    mit = np.concatenate(rows)
    ids_ref, sc_ref = c_api.score_mask_topk(eu, None, ei, None, mrp, mit, k)
    ids, sc = ops.score_mask_topk(torch.from_numpy(eu).cuda(), torch.from_numpy(ei).cuda(), k,
                                  mask_rowptr=torch.from_numpy(mrp).cuda(), mask_items=torch.from_numpy(mit).cuda(),
                                  precision="tc")
    st = ops.last_tc_stats()
    assert np.array_equal(ids.cpu().numpy(), ids_ref)
    assert np.array_equal(sc.cpu().numpy(), sc_ref)
    assert ops.last_tc_fallback_rows() == 0
    n_groups, full = (b + 255) // 256, (n_items + 127) // 128
    assert 0 < st["tiles_swept"] < 0.5 * n_groups * full, st


@pytest.mark.parametrize("n_items,k,b,with_users,d", [(30000, 50, 1500, False, 64), (30000, 100, 700, True, 64),
                                                      (200, 50, 300, False, 64), (128, 20, 260, False, 64),
                                                      (129, 64, 260, True, 64), (5000, 128, 515, False, 64),
                                                      (20000, 50, 600, True, 128), (20000, 80, 300, False, 128),
                                                      (20000, 50, 600, False, 192)])
def test_score_screen_exact_head(ops, monkeypatch, n_items, k, b, with_users, d):
    """The exact head (score_head_kernel): with popularity-skewed norms most rows are settled by the exact scores of
    the 128 / 256 highest-norm items plus the Cauchy-Schwarz bound of the rest; rows whose history covers the popular
    items, or whose K-th head score does not beat the bound, continue through the screen.  Ids AND scores must equal
    the oracle's either way, and switching the head off must not change a bit."""
    monkeypatch.setenv("GMR_SCREEN_STATS", "1")
    rng = np.random.default_rng(n_items + k + d)
    scale = np.exp(rng.normal(0.0, 1.5, size=(n_items, 1)))
    # propagated embeddings share a dominant direction (cosines near 1): only then does |u| |e| bound anything
    ei = ((1.0 + 0.3 * rng.standard_normal((n_items, d))) * scale).astype(np.float32)
    n_users = b + 37 if with_users else b
    eu = (1.0 + 0.3 * rng.standard_normal((n_users, d))).astype(np.float32)
    eu[3] = 0.0                                            # a zero row: every score ties at 0, cannot finish early
    eu[5] *= 1e-30                                         # squares underflow: the norm is no bound, the screen takes the row
    users = rng.permutation(n_users)[:b].astype(np.int64) if with_users else None
    hot = np.argsort(-scale[:, 0])[:min(400, n_items)]
    lens = rng.integers(0, min(100, len(hot)), size=b)
    lens[::40] = min(len(hot), 380)                        # these rows have seen nearly every popular item
    rows = [np.sort(rng.choice(hot, size=n, replace=False)).astype(np.int32) for n in lens]
    mrp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    mit = np.concatenate(rows)
    ids_ref, sc_ref = c_api.score_mask_topk(eu, users, ei, None, mrp, mit, k)
    args = dict(users=None if users is None else torch.from_numpy(users).cuda(), mask_rowptr=torch.from_numpy(mrp).cuda(),
                mask_items=torch.from_numpy(mit).cuda(), precision="tc")
    ids, sc = ops.score_mask_topk(torch.from_numpy(eu).cuda(), torch.from_numpy(ei).cuda(), k, **args)
    st = ops.last_tc_stats()
    assert np.array_equal(ids.cpu().numpy(), ids_ref)
    assert np.array_equal(sc.cpu().numpy(), sc_ref)
    assert 0 < st["head_rows"] < b, st                     # some rows settled early, the masked-out / zero ones not
    monkeypatch.setenv("GMR_SCREEN_HEAD", "0")
    ids0, sc0 = ops.score_mask_topk(torch.from_numpy(eu).cuda(), torch.from_numpy(ei).cuda(), k, **args)
    assert ops.last_tc_stats()["head_rows"] == 0
    assert torch.equal(ids0, ids) and torch.equal(sc0, sc)


def test_score_screen_exact_head_without_mask(ops, monkeypatch):
    """No train-history mask at all (mask_rowptr = NULL): the head's bitmap pass is skipped and every row is a candidate."""
    monkeypatch.setenv("GMR_SCREEN_STATS", "1")
    rng = np.random.default_rng(5)
    n_items, d, k, b = 12000, 64, 50, 700
    scale = np.exp(rng.normal(0.0, 1.5, size=(n_items, 1)))
    ei = ((1.0 + 0.3 * rng.standard_normal((n_items, d))) * scale).astype(np.float32)
    eu = (1.0 + 0.3 * rng.standard_normal((b, d))).astype(np.float32)
    eu[7] = -eu[7]                                          # every score negative: the bound can never be beaten
    ids_ref, sc_ref = c_api.score_mask_topk(eu, None, ei, None, None, None, k)
    ids, sc = ops.score_mask_topk(torch.from_numpy(eu).cuda(), torch.from_numpy(ei).cuda(), k, precision="tc")
    st = ops.last_tc_stats()
    assert np.array_equal(ids.cpu().numpy(), ids_ref)
    assert np.array_equal(sc.cpu().numpy(), sc_ref)
    assert 0 < st["head_rows"] < b, st


def test_score_screen_exact_head_ties(ops, monkeypatch):
    """Two of the highest-norm items are identical rows: their scores tie exactly in every row that has not seen one of
    them, and the head must order the twins by item id like the oracle (its 32-bit selection cannot: such rows switch to
    the 64-bit (score, id) network)."""
    monkeypatch.setenv("GMR_SCREEN_STATS", "1")
    rng = np.random.default_rng(77)
    n_items, d, k, b = 20000, 64, 50, 900
    scale = np.exp(rng.normal(0.0, 1.5, size=(n_items, 1)))
    ei = ((1.0 + 0.3 * rng.standard_normal((n_items, d))) * scale).astype(np.float32)
    eu = (1.0 + 0.3 * rng.standard_normal((b, d))).astype(np.float32)
    hot = np.argsort(-scale[:, 0])
    ei[hot[4]] = ei[hot[2]]                                 # twins inside every row's top K
    lens = rng.integers(0, 60, size=b)
    rows = []
    for r in range(b):
        m = set(rng.choice(hot[:300], size=lens[r], replace=False).tolist())
        if r % 3 == 0:
            m.add(int(hot[4]))                              # this row has seen one twin: no tie left
        rows.append(np.sort(np.fromiter(m, dtype=np.int32)))
    mrp = np.concatenate([[0], np.cumsum([len(x) for x in rows])]).astype(np.int64)
    mit = np.concatenate(rows).astype(np.int32)
    ids_ref, sc_ref = c_api.score_mask_topk(eu, None, ei, None, mrp, mit, k)
    ids, sc = ops.score_mask_topk(torch.from_numpy(eu).cuda(), torch.from_numpy(ei).cuda(), k,
                                  mask_rowptr=torch.from_numpy(mrp).cuda(), mask_items=torch.from_numpy(mit).cuda(), precision="tc")
    st = ops.last_tc_stats()
    assert np.array_equal(ids.cpu().numpy(), ids_ref)
    assert np.array_equal(sc.cpu().numpy(), sc_ref)
    assert st["head_rows"] > 0, st
    assert st["head_pairs_64bit"] > 0, st                   # the 32-bit selection saw the tie and handed the rows over


def test_score_screen_head_switches_off_for_flat_norms(ops, monkeypatch):
    """Flat item norms: no row can beat the outside bound, the head must not run (and nothing changes)."""
    monkeypatch.setenv("GMR_SCREEN_STATS", "1")
    rng = np.random.default_rng(2)
    n_items, d, k, b = 4000, 64, 50, 300
    ei = rng.standard_normal((n_items, d)).astype(np.float32)
    ei /= np.linalg.norm(ei, axis=1, keepdims=True)
    eu = rng.standard_normal((b, d)).astype(np.float32)
    ids_ref, sc_ref = c_api.score_mask_topk(eu, None, ei, None, None, None, k)
    ids, sc = ops.score_mask_topk(torch.from_numpy(eu).cuda(), torch.from_numpy(ei).cuda(), k, precision="tc")
    assert ops.last_tc_stats()["head_rows"] == 0
    assert np.array_equal(ids.cpu().numpy(), ids_ref)
    assert np.array_equal(sc.cpu().numpy(), sc_ref)


def test_score_screen_heavy_mask(ops):
    """Rows whose train history covers most of the catalogue (fewer than K unmasked items for some): masked
    items must surface with -1e10 exactly like trainer.py:384."""
    rng = np.random.default_rng(10)
    n_items, d, k, b = 400, 64, 50, 270
    eu = rng.standard_normal((b, d)).astype(np.float32)
    ei = rng.standard_normal((n_items, d)).astype(np.float32)
    rowptr = [0]
    items = []
    for r in range(b):
        n = n_items - (r % 80)                      # leaves 0..79 unmasked items
        items.append(np.sort(rng.choice(n_items, size=n, replace=False)).astype(np.int32))
        rowptr.append(rowptr[-1] + n)
    mrp, mit = np.asarray(rowptr, dtype=np.int64), np.concatenate(items)
    ids_ref, sc_ref = c_api.score_mask_topk(eu, None, ei, None, mrp, mit, k)
    ids, sc = ops.score_mask_topk(torch.from_numpy(eu).cuda(), torch.from_numpy(ei).cuda(), k,
                                  mask_rowptr=torch.from_numpy(mrp).cuda(), mask_items=torch.from_numpy(mit).cuda(),
                                  precision="tc")
    assert np.array_equal(ids.cpu().numpy(), ids_ref)
    assert np.array_equal(sc.cpu().numpy(), sc_ref)


@pytest.mark.parametrize("precision", TC_MODES)
def test_score_tc_matches_oracle_bit_exact(ops, precision):
    rng = np.random.default_rng(77)
    n_users, b, n_items, d, k = 400, 300, 2500, 64, 50
    eu = rng.standard_normal((n_users, d)).astype(np.float32)
    ei = rng.standard_normal((n_items, d)).astype(np.float32)
    users = rng.choice(n_users, size=b, replace=False).astype(np.int64)
    mrp, mit = make_mask(rng, b, n_items, 40)
    ids_ref, sc_ref = c_api.score_mask_topk(eu, users, ei, None, mrp, mit, k)
    ids, sc = ops.score_mask_topk(torch.from_numpy(eu).cuda(), torch.from_numpy(ei).cuda(), k,
                                  users=torch.from_numpy(users).cuda(), mask_rowptr=torch.from_numpy(mrp).cuda(),
                                  mask_items=torch.from_numpy(mit).cuda(), precision=precision)
    assert np.array_equal(ids.cpu().numpy(), ids_ref)
    assert np.array_equal(sc.cpu().numpy(), sc_ref)


@pytest.mark.parametrize("precision", TC_MODES)
def test_score_tc_ties_force_exact_fallback(ops, precision):
    """Many duplicated items: near/exact ties everywhere, so certification must fail for most rows and
    the fp32 redo must still deliver the oracle's total order (score desc, id asc)."""
    rng = np.random.default_rng(5)
    n_items, d, k = 1500, 64, 50
    base = rng.standard_normal((30, d)).astype(np.float32)
    ei = base[rng.integers(0, 30, size=n_items)]          # only 30 distinct item vectors
    eu = rng.standard_normal((260, d)).astype(np.float32)
    ids_ref, sc_ref = c_api.score_mask_topk(eu, None, ei, None, None, None, k)
    ids, sc = ops.score_mask_topk(torch.from_numpy(eu).cuda(), torch.from_numpy(ei).cuda(), k, precision=precision)
    fb = ops.last_tc_fallback_rows()
    assert np.array_equal(ids.cpu().numpy(), ids_ref)
    assert np.array_equal(sc.cpu().numpy(), sc_ref)
    if precision == "tc_split":
        assert fb > 50  # certification fails wherever K-th/KP-th candidates are (near) ties
    # the screen keeps every tied candidate (its buffer holds the ~50 copies of each leading vector)


@pytest.mark.parametrize("precision", TC_MODES)
def test_score_tc_wide_dynamic_range(ops, precision):
    """Rows and items with norms spread over 4 decades: the error bound scales with the norms."""
    rng = np.random.default_rng(6)
    n_items, d, k = 6000, 64, 50
    eu = (rng.standard_normal((300, d)) * np.exp(rng.uniform(-4, 4, size=(300, 1)))).astype(np.float32)
    ei = (rng.standard_normal((n_items, d)) * np.exp(rng.uniform(-4, 4, size=(n_items, 1)))).astype(np.float32)
    ids_ref, sc_ref = c_api.score_mask_topk(eu, None, ei, None, None, None, k)
    ids, sc = ops.score_mask_topk(torch.from_numpy(eu).cuda(), torch.from_numpy(ei).cuda(), k, precision=precision)
    assert np.array_equal(ids.cpu().numpy(), ids_ref)
    assert np.array_equal(sc.cpu().numpy(), sc_ref)


def test_score_tc_unsupported_shapes_are_loud(ops):
    with pytest.raises(RuntimeError):
        ops.score_mask_topk(torch.zeros((4, 60), device="cuda"), torch.zeros((90, 60), device="cuda"), 5, precision="tc")


# ---- row glue + size-independent properties at larger sizes --------------------------------------------


def test_rows_normalize_mix(ops):
    rng = np.random.default_rng(21)
    n, d = 777, 64
    x1, x2, y = (torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32)).cuda() for _ in range(3))
    x1[5] = 0.0                                    # zero row: F.normalize clamps the norm at eps
    out = ops.rows_normalize_mix(x1, x2, 0.3, 0.7, y=y, slope=0.2)
    lr = torch.nn.functional.leaky_relu
    z = 0.3 * torch.nn.functional.normalize(lr(x1, 0.2)) + 0.7 * torch.nn.functional.normalize(lr(x2, 0.2))
    assert out.shape == (n, 2 * d)
    assert rel(out[:, :d].cpu().numpy(), z.double().cpu().numpy()) < 1e-6
    assert rel(out[:, d:].cpu().numpy(), (z + y).double().cpu().numpy()) < 1e-6
    out1 = ops.rows_normalize_mix(x1, x2, 1.0, 0.0)  # no y, identity activation
    assert out1.shape == (n, d)
    assert rel(out1.cpu().numpy(), torch.nn.functional.normalize(x1).double().cpu().numpy()) < 1e-6


def test_spmm_linearity_and_adjoint_large(ops):
    """Properties that hold at any size: A(x + 2y) = Ax + 2Ay and <Ax, z> = <x, A'z> (power-law rows, 4M nnz)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    n_rows, n_cols, nnz, d = 150_000, 90_000, 4_000_000, 64
    w = torch.rand(n_rows, device="cuda", generator=g) ** 30         # heavy-tailed row degrees
    rows = torch.multinomial(w, nnz, replacement=True, generator=g)
    cols = torch.randint(0, n_cols, (nnz,), device="cuda", generator=g)
    val = torch.rand(nnz, device="cuda", generator=g)
    a = ops.GraphCSR.from_coo(torch.stack([rows, cols]), val, (n_rows, n_cols), torch.device("cuda"))
    assert a.plan_stats()["split_rows"] > 0                          # long rows are cut into chunks
    x = torch.randn(n_cols, d, device="cuda", generator=g)
    y = torch.randn(n_cols, d, device="cuda", generator=g)
    z = torch.randn(n_rows, d, device="cuda", generator=g)
    ax, ay, axy = ops.spmm_raw(a, x), ops.spmm_raw(a, y), ops.spmm_raw(a, x + 2 * y)
    assert rel(axy.cpu().numpy(), (ax.double() + 2 * ay.double()).cpu().numpy()) < 1e-5
    atz = ops.spmm_raw(a.t(), z)
    lhs, rhs = (ax.double() * z.double()).sum().item(), (x.double() * atz.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0)
    again = ops.spmm_raw(a, x)
    assert torch.equal(ax, again)                                    # run-to-run deterministic


def test_score_modes_agree_large(ops):
    """All three precisions return identical ids and scores on a catalogue of 300k items (popularity-skewed norms)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(4)
    b, n_items, d, k = 1500, 300_000, 64, 50
    scale = torch.exp(1.0 * torch.randn(n_items, 1, device="cuda", generator=g))
    ei = torch.randn(n_items, d, device="cuda", generator=g) * scale
    eu = torch.randn(b, d, device="cuda", generator=g)
    lens = torch.randint(0, 80, (b,), device="cuda", generator=g)
    mrp = torch.zeros(b + 1, dtype=torch.int64, device="cuda")
    mrp[1:] = torch.cumsum(lens, 0)
    total = int(mrp[-1])
    rid = torch.repeat_interleave(torch.arange(b, device="cuda"), lens)
    key = torch.unique(rid * n_items + torch.randint(0, n_items, (total,), device="cuda", generator=g))
    mrp = torch.searchsorted(key, torch.arange(b + 1, device="cuda") * n_items)
    mit = (key % n_items).to(torch.int32)
    res = {p: ops.score_mask_topk(eu, ei, k, mask_rowptr=mrp, mask_items=mit, precision=p) for p in ("fp32", "tc", "tc_split")}
    for p in ("tc", "tc_split"):
        assert torch.equal(res[p][0], res["fp32"][0]), p
        assert torch.equal(res[p][1], res["fp32"][1]), p


@pytest.mark.parametrize("m,k", [(1000, 4096), (129, 384), (128, 32), (5000, 96), (77, 1024)])
def test_dense_proj_fp32_accuracy(ops, m, k):
    """Split-TF32 tcgen05 projection vs an fp64 product: as accurate as a plain fp32 GEMM (and far from single-pass
    TF32, whose error would be ~1e-3)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(m + k)
    a = torch.randn(m, k, device="cuda", generator=g).clamp_(min=0)        # CNN-like non-negative features
    a[:, ::7] *= 37.0
    w = (torch.rand(k, 64, device="cuda", generator=g) - 0.5) * 0.06          # xavier-like
    c = ops.dense_proj(a, w)
    ref = a.double() @ w.double()
    err = (c.double() - ref).abs().max().item() / ref.abs().max().item()
    err32 = ((a @ w).double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < max(2e-6, 4 * err32), (err, err32)
    # strided output (column slice of a wider buffer) and row tail
    wide = torch.zeros(m, 128, device="cuda")
    ops.dense_proj(a, w, out=wide[:, 64:])
    assert torch.equal(wide[:, 64:], c) and float(wide[:, :64].abs().max()) == 0.0


def test_dense_proj_unsupported_is_loud(ops):
    a = torch.zeros(10, 48, device="cuda")
    assert not ops.dense_proj_supported(a, torch.zeros(48, 64, device="cuda"))
    with pytest.raises(RuntimeError):
        ops.dense_proj(a, torch.zeros(48, 64, device="cuda"))


def test_metrics_kernel_for_a_user_without_ground_truth(ops):
    """Same answer as the reference's metric code (and the oracle) for pos_len == 0: NaN recall, zero NDCG / MAP
    (tests/test_oracle_properties.py::test_metrics_for_a_user_without_ground_truth pins the oracle to the reference)."""
    rng = np.random.default_rng(5)
    u, k, n_items = 200, 20, 90
    topk = np.stack([rng.choice(n_items, size=k, replace=False) for _ in range(u)]).astype(np.int32)
    lens = rng.integers(1, 9, size=u)
    lens[[3, 77, 199]] = 0
    gts = [np.sort(rng.choice(n_items, size=int(n), replace=False)) for n in lens]
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    items = np.concatenate(gts).astype(np.int32)
    sums, hit = ops.hits_metrics(torch.from_numpy(topk).cuda(), torch.from_numpy(rowptr).cuda(), torch.from_numpy(items).cuda(),
                                 return_hit=True)
    got = (sums / u).cpu().numpy()
    ref_hit = c_api.hits(topk, rowptr, items)
    assert np.array_equal(hit.cpu().numpy(), ref_hit)
    ref = c_api.metrics(ref_hit, lens.astype(np.int64))
    assert np.all(np.isnan(got[0])) and np.all(np.isnan(ref["recall"]))
    for j, name in ((1, "ndcg"), (2, "precision"), (3, "map")):
        assert np.all(np.isfinite(got[j])) and np.abs(got[j] - ref[name]).max() < 1e-12, name
