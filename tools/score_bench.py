"""Timing + diagnostics of the fused score/mask/top-K kernels on the embeddings of a bench workload.

    python tools/score_bench.py --workload scaled --modes tc,tc_split --iters 3 [--check-users 4096]

Prints one JSON object: per mode the median ms, algorithmic TFLOP/s, rows redone on the fp32 path, the
screen counters (GMR_SCREEN_STATS=1 is set here) and, with --check-users N, whether the first N eval users
get ids/scores identical to the fp32 kernel.  GMR_TC_DEBUG=1/2 (drain only / filter only) are timing
experiments: results are then wrong by construction and the check is skipped.
"""
import argparse
import json
import os
import sys

os.environ.setdefault("GMR_SCREEN_STATS", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import genmmrec_b200  # noqa: F401,E402
from genmmrec_b200 import ops  # noqa: E402
from genmmrec_b200.workload import Workload  # noqa: E402


def time_fn(fn, iters, warmup=1):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2], ms[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="scaled")
    ap.add_argument("--model", default="DiffMM")
    ap.add_argument("--modes", default="tc,tc_split")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--users", type=int, default=0, help="score only the first N eval users (0 = all)")
    ap.add_argument("--check-users", type=int, default=0)
    ap.add_argument("--random", action="store_true", help="N(0,1) embeddings instead of the model's")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    wl = Workload(args.model, args.workload, dev)
    model, cfg, loader = wl.model, wl.config, wl.valid
    k = max(cfg["topk"])
    with torch.no_grad():
        eu, rows, ei, bias = model.eval_factors(loader.eval_u)
    eu, ei = eu.contiguous(), ei.contiguous()
    if args.random:
        g = torch.Generator(device=dev)
        g.manual_seed(1)
        eu = torch.randn(eu.shape, device=dev, generator=g)
        ei = torch.randn(ei.shape, device=dev, generator=g)
    mrp, mit = loader.mask_rowptr, loader.mask_items
    if args.users:
        rows = rows[:args.users].contiguous()
        mrp = mrp[:args.users + 1].contiguous()
    b, i, d = int(rows.numel()), int(ei.shape[0]), int(ei.shape[1])
    un = eu[rows].norm(dim=1)
    inorm = ei.norm(dim=1)
    out = {"workload": args.workload, "users": b, "items": i, "d": d, "k": k,
           "item_norm": {"max": float(inorm.max()), "median": float(inorm.median()), "mean": float(inorm.mean())},
           "user_norm": {"max": float(un.max()), "median": float(un.median())},
           "debug": os.environ.get("GMR_TC_DEBUG", "0"), "modes": {}}
    for mode in args.modes.split(","):
        fn = lambda: ops.score_mask_topk(eu, ei, k, users=rows, bias=bias, mask_rowptr=mrp, mask_items=mit,
                                         precision=mode, return_scores=False)
        med, best = time_fn(fn, args.iters)
        e = {"ms": med, "best_ms": best, "alg_TFLOPs": 2.0 * b * i * d / med / 1e9, "users_per_s": b / med * 1e3}
        if mode != "fp32":
            e["fallback_rows"] = ops.last_tc_fallback_rows()
        if mode == "tc":
            st = ops.last_tc_stats()
            e["stats"] = st
            n_chunks = b / 32.0 * ((i + 127) // 128) * 4
            e["slow_chunk_frac"] = st.get("slow_chunks", 0) / n_chunks if n_chunks else 0
            e["appends_per_row"] = st.get("appends", 0) / b
            e["rescored_per_row"] = st.get("rescored", 0) / b
            e["prunes_per_row"] = (st.get("cheap_prunes", 0) + st.get("exact_prunes", 0)) / b
            n_groups = (b + 255) // 256
            e["tiles_per_group"] = st.get("tiles_swept", 0) / n_groups
            e["tiles_full_sweep"] = (i + 127) // 128
        out["modes"][mode] = e
    if args.check_users and out["debug"] == "0":
        n = min(args.check_users, b)
        r2, m2 = rows[:n].contiguous(), mrp[:n + 1].contiguous()
        ref = ops.score_mask_topk(eu, ei, k, users=r2, bias=bias, mask_rowptr=m2, mask_items=mit, precision="fp32")
        for mode in args.modes.split(","):
            got = ops.score_mask_topk(eu, ei, k, users=r2, bias=bias, mask_rowptr=m2, mask_items=mit, precision=mode)
            out["modes"][mode]["equals_fp32_first_%d" % n] = bool(torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
