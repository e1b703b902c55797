"""Vectorised graph builders (host logic; torch ops on CPU or CUDA, no Python loops over edges).

Each builder returns COO parts ``(indices int64 [2, nnz], values fp32 [nnz], shape)`` whose VALUES and
ENTRY ORDER equal what the reference's scipy/Python builders produce (checked bit for bit against
``tests/golden`` at the toy shape); ``ops.GraphCSR.from_coo`` turns them into device CSR.  The
reference builders are O(nnz) Python loops (2.8 s at the Baby shape, hours at 10^8 nonzeros --
SURVEY.md section 8a); these run in milliseconds on the GPU at the 1M-user shape.

Node-degree normalisers are evaluated with numpy in float64 exactly as the reference does
(``np.power(deg, -0.5)``); the per-edge product of two float64 factors and the final cast to fp32
are IEEE-exact on any device, so CPU and CUDA builds give identical bits.
"""
import numpy as np
import torch


def _t64(x, device):
    return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(device=device, dtype=torch.int64)


def _inv_sqrt(deg_f64_numpy, eps):
    """np.power(deg + eps, -0.5) with inf -> 0, as a float64 torch tensor factory input."""
    with np.errstate(divide="ignore"):
        d = np.power(deg_f64_numpy + eps, -0.5)
    d[np.isinf(d)] = 0.0
    return d


def unique_pairs(users, items, n_items, device):
    """Sorted unique (user, item) pairs (the reference's dict keys collapse duplicates,
    GenMMRec/src/models/diffmm.py:92-96)."""
    key = torch.unique(_t64(users, device) * n_items + _t64(items, device))
    return torch.div(key, n_items, rounding_mode="floor"), key % n_items


def norm_adj(users, items, n_users, n_items, device="cpu"):
    """Symmetric-normalised bipartite adjacency  D^-1/2 [[0, R], [R^T, 0]] D^-1/2,  deg = count + 1e-7
    (GenMMRec/src/models/diffmm.py:88-107 == genrecv1.py:133-152 == lightgcn.py:65-100).
    Entries in row-major order with ascending columns, as ``sp.coo_matrix(D * A * D)`` emits them."""
    device = torch.device(device)
    u, i = unique_pairs(users, items, n_items, device)
    n = n_users + n_items
    deg_u = torch.bincount(u, minlength=n_users)
    deg_i = torch.bincount(i, minlength=n_items)
    deg = torch.cat([deg_u, deg_i]).cpu().numpy().astype(np.float64)
    dinv = torch.from_numpy(_inv_sqrt(deg, 1e-7)).to(device)
    # user rows: (u, n_users + i) already sorted by (u, i); item rows: (n_users + i, u) sorted by (i, u)
    order = torch.argsort(i * n_users + u)
    rows = torch.cat([u, n_users + i[order]])
    cols = torch.cat([n_users + i, u[order]])
    vals = (dinv[rows] * dinv[cols]).to(torch.float32)
    return torch.stack([rows, cols]), vals, (n, n)


def bipartite_halves(indices, values, n_users, n_items):
    """Split a bipartite (N x N) graph whose user rows only reach item columns and vice versa into
    R_hat [U x I] (user rows) and R_hat^T-like [I x U] (item rows).  A . [X_u; X_i] == [R X_i; R' X_u]
    row for row, with the same per-row entry order, so results are bit-identical to the full SpMM."""
    r, c = indices[0], indices[1]
    top = r < n_users
    if bool((c[top] < n_users).any()) or bool((c[~top] >= n_users).any()):
        return None  # has user-user or item-item entries (e.g. self loops): not bipartite
    ui = (torch.stack([r[top], c[top] - n_users]), values[top], (n_users, n_items))
    iu = (torch.stack([r[~top] - n_users, c[~top]]), values[~top], (n_items, n_users))
    return ui, iu


def ui_matrix(u_list, i_list, n_users, n_items, device="cpu"):
    """Modality user-item graph  D^-1/2 (bin([[0, G], [G^T, 0]]) + I) D^-1/2  with D = rowsum, no eps
    (GenMMRec/src/common/trainer.py:464-485: buildUIMatrix + normalizeAdj).  The reference's
    ``(M D)^T D`` product leaves scipy in column-major order; the edge dropper draws one random number
    per stored entry, so that order is part of the contract and is reproduced here."""
    device = torch.device(device)
    u, i = unique_pairs(u_list, i_list, n_items, device)
    n = n_users + n_items
    deg = torch.cat([torch.bincount(u, minlength=n_users), torch.bincount(i, minlength=n_items)]) + 1
    dinv = torch.from_numpy(_inv_sqrt(deg.cpu().numpy().astype(np.float64), 0.0)).to(device)
    diag = torch.arange(n, device=device)
    rows = torch.cat([u, n_users + i, diag])
    cols = torch.cat([n_users + i, u, diag])
    order = torch.argsort(rows * n + cols)  # row-major of a symmetric matrix ...
    rows, cols = rows[order], cols[order]
    vals = (dinv[rows] * dinv[cols]).to(torch.float32)
    return torch.stack([cols, rows]), vals, (n, n)  # ... listed transposed == column-major


def drop_edges(indices, values, keep_rate):
    """SpAdjDropEdge (GenMMRec/src/models/diffmm.py:287-301): keep entry e iff
    floor(rand_e + keep_rate) != 0, kept values / keep_rate.  Like the reference, the random numbers
    come from torch's CPU generator even when the graph lives on the GPU."""
    mask = (torch.rand(values.size()) + keep_rate).floor().type(torch.bool).to(values.device)
    return indices[:, mask], values[mask] / keep_rate


def knn_from_topk(knn_val, knn_ind):
    """Weighted symmetric normalisation of a kNN list (GenMMRec/src/utils/utils.py:152-165,184-197):
    deg_i = sum_j val_ij,  w_ij = deg_i^-1/2 * val_ij * deg_j^-1/2, inf -> 0; exactly k entries per
    row in top-k order."""
    n, k = knn_ind.shape
    row = torch.arange(n, device=knn_ind.device).repeat_interleave(k)
    col = knn_ind.reshape(-1).to(torch.int64)
    w = knn_val.reshape(-1).to(torch.float32)
    deg = torch.zeros(n, dtype=w.dtype, device=w.device).index_add_(0, row, w)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    w = dis[row] * w * dis[col]
    return torch.stack([row, col]), w, (n, n)


def knn_graph_dense(feat, k, eps_normalize=False):
    """Reference-shaped kNN graph through a dense similarity matrix (utils.py:147-150,184-197;
    GenRecV1's variant normalises with F.normalize, common/trainer.py:682-687).  O(I^2) memory:
    for small item counts and parity checks only -- `knn_graph_fused` is the scalable path."""
    feat = feat.to(torch.float32)
    if eps_normalize:
        fn = torch.nn.functional.normalize(feat, p=2, dim=-1)
    else:
        fn = feat.div(torch.norm(feat, p=2, dim=-1, keepdim=True))
    sim = torch.mm(fn, fn.transpose(1, 0))
    val, ind = torch.topk(sim, k, dim=-1)
    return knn_from_topk(val, ind)


def knn_graph_fused(feat, k, eps_normalize=False, precision="auto"):
    """kNN graph WITHOUT the dense I x I similarity matrix: the fused score + top-K kernel (K2/K3) run
    on the row-normalised features against themselves.  Neighbour
    order is (similarity desc, item id asc); similarities are the fp32 FMA chain of the kernel, so
    weights agree with the dense builder to ~1e-6 relative and neighbour sets up to exact ties.
    ``precision='auto'`` takes the tensor cores whenever the feature width is a multiple of 64 (the 4096-d image and
    384-d text tables are): the K-chunked split-bf16 tcgen05 product with exact fp32 re-scoring of the candidates
    (csrc/score_topk_tc.cu, STREAM_A); every precision returns the same ids and similarities."""
    from . import ops

    feat = feat.to(torch.float32)
    if eps_normalize:
        fn = torch.nn.functional.normalize(feat, p=2, dim=-1)
    else:
        fn = feat.div(torch.norm(feat, p=2, dim=-1, keepdim=True))
    fn = fn.contiguous()
    ids, sc = ops.score_mask_topk(fn, fn, k, precision=precision)
    return knn_from_topk(sc, ids.to(torch.int64))


def gume_adj(users, items, n_users, n_items, img_knn_ind, txt_knn_ind, device="cpu"):
    """GUME's enhanced graph (GenMMRec/src/models/gume.py:122-201):
    II[i, j] = 1 for j in kNN_img(i) & kNN_txt(i), j != i;  A = [[0, R], [R^T, II]] with R SUMMING
    duplicate pairs (``tolil``);  A_hat = D^-1/2 A D^-1/2 with D = rowsum, inf -> 0.
    Returns (norm_adj parts, R parts) with R = A_hat[:U, U:]; both row-major, ascending columns."""
    device = torch.device(device)
    u, i = _t64(users, device), _t64(items, device)
    n = n_users + n_items
    key, cnt = torch.unique(u * n_items + i, return_counts=True)
    ru, ri = torch.div(key, n_items, rounding_mode="floor"), key % n_items
    rv = cnt.to(torch.float64)
    a, b = img_knn_ind.to(device), txt_knn_ind.to(device)
    common = (a.unsqueeze(2) == b.unsqueeze(1)).any(dim=2)          # [I, k]: a_ij appears in b_i
    src = torch.arange(n_items, device=device).unsqueeze(1).expand_as(a)
    sel = common & (a != src)
    ii_key = torch.unique(src[sel] * n_items + a[sel].to(torch.int64))
    ii_r, ii_c = torch.div(ii_key, n_items, rounding_mode="floor"), ii_key % n_items
    order = torch.argsort(ri * n_users + ru)
    rows = torch.cat([ru, n_users + ri[order], n_users + ii_r])
    cols = torch.cat([n_users + ri, ru[order], n_users + ii_c])
    vals = torch.cat([rv, rv[order], torch.ones(ii_r.numel(), dtype=torch.float64, device=device)])
    order2 = torch.argsort(rows * n + cols)
    rows, cols, vals = rows[order2], cols[order2], vals[order2]
    # rowsum in float32 like scipy's sum over the float32 LIL-derived matrix, then float64 power
    rowsum = torch.zeros(n, dtype=torch.float64, device=device).index_add_(0, rows, vals)
    dinv32 = _inv_sqrt(rowsum.cpu().numpy().astype(np.float32), np.float32(0.0))
    dinv = torch.from_numpy(dinv32.astype(np.float32)).to(device)
    v32 = vals.to(torch.float32)
    nv = (dinv[rows] * v32) * dinv[cols]
    top = (rows < n_users)
    r_parts = (torch.stack([rows[top], cols[top] - n_users]), nv[top], (n_users, n_items))
    return (torch.stack([rows, cols]), nv, (n, n)), r_parts


def ld4mrec_rnorm(users, items, n_users, n_items, device="cpu"):
    """d_u^-1/2 R d_i^-1/2 with inf -> 0 (GenMMRec/src/models/ld4mrec.py:181-203); R sums duplicates."""
    device = torch.device(device)
    u, i = _t64(users, device), _t64(items, device)
    key, cnt = torch.unique(u * n_items + i, return_counts=True)
    ru, ri = torch.div(key, n_items, rounding_mode="floor"), key % n_items
    v = cnt.to(torch.float32)
    du = torch.zeros(n_users, dtype=torch.float32, device=device).index_add_(0, ru, v)
    di = torch.zeros(n_items, dtype=torch.float32, device=device).index_add_(0, ri, v)
    du_inv = torch.from_numpy(_inv_sqrt(du.cpu().numpy(), np.float32(0.0))).to(device)
    di_inv = torch.from_numpy(_inv_sqrt(di.cpu().numpy(), np.float32(0.0))).to(device)
    vals = (du_inv[ru] * v) * di_inv[ri]  # float32 throughout, as scipy evaluates it
    return torch.stack([ru, ri]), vals, (n_users, n_items)


def binary_r(users, items, n_users, n_items, device="cpu"):
    """Raw train matrix R (GenMMRec/src/models/genrecv1.py:54,128-131), duplicates summed,
    row-major ascending columns (see oracle/ref_port.py:genrecv1_r for the scipy note)."""
    device = torch.device(device)
    key, cnt = torch.unique(_t64(users, device) * n_items + _t64(items, device), return_counts=True)
    return (torch.stack([torch.div(key, n_items, rounding_mode="floor"), key % n_items]), cnt.to(torch.float32),
            (n_users, n_items))
