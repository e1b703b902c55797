"""Shared parity helpers for the tests (comparison rules stated once)."""
import numpy as np

# North star: "top-K item IDs bit-exact except at score ties narrower than the stated tolerance".
# Stated tolerance: two rankings are equivalent when, position by position, the float64 scores of
# the items they name differ by at most TIE_TOL * max|score| of that row.  This admits exactly the
# swaps/boundary substitutions that fp32 summation-order differences (cuBLAS / MKL / our fmaf chain,
# ~1e-6 relative) can cause and nothing else.
TIE_TOL = 2e-5
EMB_TOL = 1e-5      # propagated embeddings: max|Y - Y_ref| / max|Y_ref| (fp32)
METRIC_TOL = 1e-6   # unrounded Recall/NDCG/Precision/MAP vectors


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def assert_topk_equivalent(ids, ref_ids, scores64_fn, tol=TIE_TOL, min_exact_rows=0.98):
    """ids/ref_ids [n, K]; scores64_fn(row) -> float64 [n_items] masked score vector of that row."""
    ids, ref_ids = np.asarray(ids), np.asarray(ref_ids)
    assert ids.shape == ref_ids.shape
    diff_rows = np.flatnonzero((ids != ref_ids).any(axis=1))
    assert 1.0 - diff_rows.size / max(ids.shape[0], 1) >= min_exact_rows, \
        "only %.4f of the rows are identical" % (1.0 - diff_rows.size / ids.shape[0])
    for r in diff_rows:
        s = scores64_fn(int(r))
        scale = np.abs(s[np.isfinite(s) & (s > -1e9)]).max()
        gap = np.abs(s[ids[r]] - s[ref_ids[r]]).max()
        assert gap <= tol * scale, "row %d: rankings differ beyond the tie tolerance (gap %.3e, scale %.3e)" % (r, gap, scale)
    return diff_rows.size
