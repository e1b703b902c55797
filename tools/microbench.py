"""Developer micro-benchmark of the raw kernels on one GPU (not the contract bench; see bench.py).

    python tools/microbench.py [--users 1000000 --items 500000 --nnz 50000000 --d 64]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genmmrec_b200 import ops  # noqa: E402


def powerlaw_bipartite(n_users, n_items, nnz, dev, alpha=0.8, seed=0):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    w = torch.exp(torch.randn(n_users, device=dev, generator=g))
    deg = 5 + torch.floor(w / w.sum() * (nnz - 5 * n_users)).long()
    users = torch.repeat_interleave(torch.arange(n_users, device=dev), deg)
    p = torch.arange(1, n_items + 1, device=dev, dtype=torch.float64) ** (-alpha)
    cdf = torch.cumsum(p / p.sum(), 0)
    perm = torch.randperm(n_items, device=dev, generator=g)
    r = torch.rand(users.numel(), device=dev, generator=g, dtype=torch.float64)
    items = perm[torch.searchsorted(cdf, r).clamp_(max=n_items - 1)]
    return users, items


def time_fn(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=1_000_000)
    ap.add_argument("--items", type=int, default=500_000)
    ap.add_argument("--nnz", type=int, default=50_000_000)
    ap.add_argument("--d", type=int, default=64)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--eval-users", type=int, default=16384)
    ap.add_argument("--k", type=int, default=50)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--skip-score", action="store_true")
    ap.add_argument("--torch-ref", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    out = {"args": vars(args), "device": torch.cuda.get_device_name(0)}
    u, i = powerlaw_bipartite(args.users, args.items, args.nnz, dev)
    nnz = u.numel()
    nu, ni, d = args.users, args.items, args.d
    n = nu + ni
    val = torch.rand(2 * nnz, device=dev) * 0.1
    idx = torch.stack([torch.cat([u, i + nu]), torch.cat([i + nu, u])])
    t0 = time.time()
    adj = ops.GraphCSR.from_coo(idx, val, (n, n), dev, chunk_nnz=args.chunk)
    torch.cuda.synchronize()
    out["csr_build_s"] = time.time() - t0
    t0 = time.time()
    out["plan"] = adj.plan_stats()
    out["plan_s"] = time.time() - t0
    x = torch.randn(n, d, device=dev)
    y = torch.empty(n, d, device=dev)
    med, best = time_fn(lambda: ops.spmm_raw(adj, x, out=y), args.iters)
    balg = adj.algorithmic_bytes(d)
    out["spmm_full"] = {"ms": med, "best_ms": best, "alg_GBs": balg / med / 1e6, "gather_GBs": (adj.nnz * (8 + 4 * d)) / med / 1e6,
                        "alg_MB": balg / 1e6}
    # halves
    r_u = ops.GraphCSR.from_coo(torch.stack([u, i]), val[:nnz], (nu, ni), dev, chunk_nnz=args.chunk)
    r_i = ops.GraphCSR.from_coo(torch.stack([i, u]), val[:nnz], (ni, nu), dev, chunk_nnz=args.chunk)
    xu, xi = x[:nu], x[nu:]
    yu, yi = y[:nu], y[nu:]
    med, best = time_fn(lambda: ops.spmm_raw(r_u, xi, out=yu), args.iters)
    out["spmm_user_rows"] = {"ms": med, "alg_GBs": r_u.algorithmic_bytes(d) / med / 1e6, "plan": r_u.plan_stats()}
    med, best = time_fn(lambda: ops.spmm_raw(r_i, xu, out=yi), args.iters)
    out["spmm_item_rows"] = {"ms": med, "alg_GBs": r_i.algorithmic_bytes(d) / med / 1e6, "plan": r_i.plan_stats()}
    if args.torch_ref:
        a_t = adj.to_torch_coo().coalesce().to_sparse_csr()
        med, best = time_fn(lambda: torch.sparse.mm(a_t, x), max(3, args.iters // 2))
        out["torch_csr_spmm_ms"] = med
        diff = (torch.sparse.mm(a_t, x) - ops.spmm_raw(adj, x)).abs().max().item()
        out["torch_vs_ours_maxabs"] = diff
    if not args.skip_score:
        b = args.eval_users
        users = torch.randperm(nu, device=dev)[:b].contiguous()
        ue = torch.randn(nu, d, device=dev)
        ie = torch.randn(ni, d, device=dev)
        mrp = torch.arange(0, b + 1, device=dev, dtype=torch.int64) * 8
        mit = torch.sort(torch.randint(0, ni, (b, 8), device=dev, dtype=torch.int32), dim=1).values.reshape(-1).contiguous()
        fn = lambda: ops.score_mask_topk(ue, ie, args.k, users=users, mask_rowptr=mrp, mask_items=mit, precision="fp32")
        med, best = time_fn(fn, max(2, args.iters // 3), warmup=1)
        out["score_fp32"] = {"ms": med, "users_per_s": b / med * 1e3, "TFLOPs": 2.0 * b * ni * d / med / 1e9}
        if ops.tc_supported(d, args.k):
            fn = lambda: ops.score_mask_topk(ue, ie, args.k, users=users, mask_rowptr=mrp, mask_items=mit, precision="tc")
            med, best = time_fn(fn, max(3, args.iters // 2), warmup=2)
            out["score_tc"] = {"ms": med, "users_per_s": b / med * 1e3, "TFLOPs": 2.0 * b * ni * d / med / 1e9,
                               "fallback_rows": ops.last_tc_fallback_rows()}
            a = ops.score_mask_topk(ue, ie, args.k, users=users, mask_rowptr=mrp, mask_items=mit, precision="tc")
            b2 = ops.score_mask_topk(ue, ie, args.k, users=users, mask_rowptr=mrp, mask_items=mit, precision="fp32")
            out["score_tc_equals_fp32"] = bool(torch.equal(a[0], b2[0]) and torch.equal(a[1], b2[1]))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
