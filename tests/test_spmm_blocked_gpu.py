"""K1b (csrc/spmm_flat.cu): the column-blocked, nonzero-centric SpMM against the C oracle, through the C ABI.

Cases: every block width from one column to the whole matrix, rows that are empty / short / longer than a piece /
longer than many tiles, duplicate entries, unsorted columns inside a row, ragged widths, alpha / beta with strided
views, value refresh, and the canonical-order property (a row's bits do not depend on which rows share the matrix)."""
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


def random_csr(rng, n_rows, n_cols, avg, long_rows=(), empty_rows=(), sort_cols=False):
    deg = rng.poisson(avg, size=n_rows).astype(np.int64)
    for r, n in long_rows:
        deg[r] = n
    for r in empty_rows:
        deg[r] = 0
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    col = rng.integers(0, n_cols, size=int(rowptr[-1])).astype(np.int32)  # duplicates allowed
    if sort_cols:
        for r in range(n_rows):
            col[rowptr[r]:rowptr[r + 1]].sort()
    val = rng.standard_normal(col.size).astype(np.float32)
    return rowptr, col, val


def to_graph(ops, rowptr, col, val, shape):
    dev = torch.device("cuda:0")
    return ops.GraphCSR(torch.from_numpy(rowptr).to(dev), torch.from_numpy(col).to(dev), torch.from_numpy(val).to(dev), shape)


def rel(a, b):
    return np.abs(a.astype(np.float64) - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("d", [64, 128, 32, 20, 4, 516])
@pytest.mark.parametrize("block_cols", [1, 37, 600, 2500, 10 ** 9])
def test_blocked_matches_oracle(ops, d, block_cols):
    rng = np.random.default_rng(d * 7919 + block_cols % 1000)
    n_rows, n_cols = 3000, 2500
    rowptr, col, val = random_csr(rng, n_rows, n_cols, 20, long_rows=[(7, 5000), (1500, 700), (2999, 257), (8, 64), (9, 65), (10, 128), (12, 129)],
                                  empty_rows=[0, 11, 2998], sort_cols=(block_cols % 2 == 1))
    x = rng.standard_normal((n_cols, d)).astype(np.float32)
    g = to_graph(ops, rowptr, col, val, (n_rows, n_cols))
    xd = torch.from_numpy(x).cuda()
    y = ops.spmm_blocked(g, xd, block_cols).cpu().numpy()
    y64 = c_api.spmm_csr_f64(rowptr, col, val, x)
    assert rel(y, y64) < 1e-5  # the north star's tolerance: 1e-5 relative (fp32), per matrix
    assert np.all(y[0] == 0) and np.all(y[11] == 0) and np.all(y[2998] == 0)
    y2 = ops.spmm_blocked(g, xd, block_cols).cpu().numpy()
    assert np.array_equal(y, y2)  # deterministic
    st = g.blocked_plan_stats(block_cols)
    assert st["blocks"] == -(-n_cols // st["block_cols"]) and st["blocks"] <= 256  # blocks are widened past 256
    # beta = 0 must not read Y (NaN-filled output buffer)
    out = torch.full((n_rows, d), float("nan"), device="cuda")
    ops.spmm_blocked(g, xd, block_cols, out=out)
    assert np.array_equal(out.cpu().numpy(), y)


def test_blocked_alpha_beta_and_strided_views(ops):
    rng = np.random.default_rng(3)
    n_rows, n_cols = 1000, 800
    rowptr, col, val = random_csr(rng, n_rows, n_cols, 9, long_rows=[(5, 900), (6, 70)], empty_rows=[17])
    g = to_graph(ops, rowptr, col, val, (n_rows, n_cols))
    xw = torch.from_numpy(rng.standard_normal((n_cols, 192)).astype(np.float32)).cuda()
    for bc in (64, 300, 800):
        yw = torch.from_numpy(rng.standard_normal((n_rows, 192)).astype(np.float32)).cuda()
        y0 = yw.clone()
        x = xw[:, 64:128]
        ops.spmm_blocked(g, x, bc, out=yw[:, 128:192], alpha=0.5, beta=2.0)
        ref = 0.5 * c_api.spmm_csr_f64(rowptr, col, val, x.cpu().numpy()) + 2.0 * y0[:, 128:192].cpu().numpy()
        assert rel(yw[:, 128:192].cpu().numpy(), ref) < 1e-5
        assert torch.equal(yw[:, :128], y0[:, :128])  # neighbours of the slice untouched
        assert torch.equal(yw[17, 128:192], 2.0 * y0[17, 128:192])  # empty row: beta * Y


def test_blocked_rows_are_canonical_under_row_sharding(ops):
    """The bits of a row depend on the row alone: a row block of the matrix (what a rank owns under row sharding)
    returns exactly the rows the whole matrix returns."""
    rng = np.random.default_rng(5)
    n_rows, n_cols, d = 4000, 3000, 64
    rowptr, col, val = random_csr(rng, n_rows, n_cols, 30, long_rows=[(1, 4000), (1999, 300), (2000, 90), (3999, 1000)],
                                  empty_rows=[2, 2001])
    g = to_graph(ops, rowptr, col, val, (n_rows, n_cols))
    x = torch.from_numpy(rng.standard_normal((n_cols, d)).astype(np.float32)).cuda()
    for bc in (500, 3000):
        full = ops.spmm_blocked(g, x, bc)
        for r0, r1 in ((0, 1333), (1333, 2001), (2001, 4000)):
            part = ops.spmm_blocked(g.row_block(r0, r1), x, bc)
            assert torch.equal(part, full[r0:r1])


def test_blocked_value_refresh(ops):
    rng = np.random.default_rng(6)
    rowptr, col, val = random_csr(rng, 700, 900, 12, long_rows=[(3, 500)])
    g = to_graph(ops, rowptr, col, val, (700, 900))
    x = torch.from_numpy(rng.standard_normal((900, 64)).astype(np.float32)).cuda()
    y1 = ops.spmm_blocked(g, x, 200)
    g.val.mul_(-2.0)  # in place: the plan's snapshot is stale until the version check re-reads the values
    y2 = ops.spmm_blocked(g, x, 200)
    assert torch.equal(y2, -2.0 * y1)


def test_blocked_degenerate_shapes(ops):
    dev = torch.device("cuda:0")
    # no nonzeros at all
    g = ops.GraphCSR(torch.zeros(6, dtype=torch.int32, device=dev), torch.zeros(0, dtype=torch.int32, device=dev),
                     torch.zeros(0, dtype=torch.float32, device=dev), (5, 9))
    x = torch.randn(9, 8, device=dev)
    out = torch.randn(5, 8, device=dev)
    o0 = out.clone()
    ops.spmm_blocked(g, x, 4, out=out, alpha=1.0, beta=3.0)
    assert torch.equal(out, 3.0 * o0)
    assert torch.count_nonzero(ops.spmm_blocked(g, x, 4)) == 0
    # one row, one entry
    g = ops.GraphCSR(torch.tensor([0, 1], dtype=torch.int32, device=dev), torch.tensor([2], dtype=torch.int32, device=dev),
                     torch.tensor([1.5], dtype=torch.float32, device=dev), (1, 3))
    x = torch.randn(3, 4, device=dev)
    assert torch.equal(ops.spmm_blocked(g, x, 1), 1.5 * x[2:3])


def test_spmm_raw_dispatch_env(ops, monkeypatch):
    """GMR_SPMM_BLOCKED=1 routes spmm_raw through K1b with the GMR_SPMM_BLOCK_MB slice size; 0 keeps K1."""
    rng = np.random.default_rng(8)
    rowptr, col, val = random_csr(rng, 2000, 70000, 25, long_rows=[(4, 3000)])
    g = to_graph(ops, rowptr, col, val, (2000, 70000))
    x = torch.from_numpy(rng.standard_normal((70000, 64)).astype(np.float32)).cuda()
    ref = c_api.spmm_csr_f64(rowptr, col, val, x.cpu().numpy())
    monkeypatch.setenv("GMR_SPMM_BLOCKED", "1")
    monkeypatch.setenv("GMR_SPMM_BLOCK_MB", "4")  # 16384 rows of 256 B per block -> 5 blocks
    y = ops.spmm_raw(g, x)
    assert rel(y.cpu().numpy(), ref) < 1e-5
    assert len(g._bplans) == 1 and g.blocked_plan_stats(next(iter(g._bplans)))["blocks"] == 5
    monkeypatch.setenv("GMR_SPMM_BLOCKED", "0")
    y0 = ops.spmm_raw(g, x)
    assert rel(y0.cpu().numpy(), ref) < 1e-5
    assert float((y - y0).abs().max()) < 1e-4
