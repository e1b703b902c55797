"""Probe: K1b (column-blocked, nonzero-centric SpMM) against K1 (row-centric) on the bench workload's own graphs, for
a sweep of block sizes.  Diagnostics only.

    python tools/spmm_flat_probe.py --workload scaled --mb 0,32,48,64,96
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import genmmrec_b200  # noqa: F401,E402
from genmmrec_b200 import ops  # noqa: E402
from genmmrec_b200.workload import Workload  # noqa: E402


def time_fn(fn, iters=6, warmup=2):
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
    return ms[len(ms) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="scaled")
    ap.add_argument("--mb", default="0,32,48,64,96")
    ap.add_argument("--out", default=None)
    ap.add_argument("--cases", default="full,iu,ui")
    ap.add_argument("--iters", type=int, default=6)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    wl = Workload("DiffMM", args.workload, dev)
    adj = wl.model.norm_adj
    nu, ni, d = wl.n_users, wl.n_items, 64
    n = nu + ni
    res = {"workload": args.workload, "cases": {}}
    cases = {
        "full[N x N] D=64": (adj.full, torch.randn(n, d, device=dev), 64),
        "iu[I x U] D=64": (adj.iu, torch.randn(nu, d, device=dev), 64),
        "ui[U x I] D=128": (adj.ui, torch.randn(ni, 2 * d, device=dev), 128),
    }
    m = wl.model
    w0, w1 = m._modal_weights_host()
    mix = m._modal_mix_graph(m.image_UI_matrix, m.text_UI_matrix, w0, w1)
    cases["mix[N x N] D=64"] = (mix, torch.randn(n, d, device=dev), 64)   # ~5 nonzeros per row
    os.environ["GMR_SPMM_BLOCKED"] = "0"
    keep = args.cases.split(",")
    for name, (g, x, dd) in cases.items():
        if name.split("[")[0] not in keep:
            continue
        out = torch.empty(g.shape[0], dd, device=dev)
        alg = g.algorithmic_bytes(dd)
        ref = ops.spmm_raw(g, x, out=out).clone()
        ms = time_fn(lambda: ops.spmm_raw(g, x, out=out), iters=args.iters)
        entry = {"row_kernel_ms": ms, "row_kernel_alg_GBs": alg / ms / 1e6, "nnz": g.nnz, "alg_bytes": alg, "blocked": {}}
        for mb in [int(v) for v in args.mb.split(",")]:
            bc = g.shape[1] if mb == 0 else max(1, (mb << 20) // (dd * 4))
            t0 = time.time()
            g.blocked_plan(bc)
            torch.cuda.synchronize()
            plan_s = time.time() - t0
            y = ops.spmm_blocked(g, x, bc, out=out)
            err = float((y - ref).abs().max() / ref.abs().max())
            ms = time_fn(lambda: ops.spmm_blocked(g, x, bc, out=out), iters=args.iters)
            st = g.blocked_plan_stats(bc)
            entry["blocked"]["%dMB" % mb] = {"ms": ms, "alg_GBs": alg / ms / 1e6, "gather_GBs": g.nnz * (8 + 4 * dd) / ms / 1e6,
                                            "rel_diff_vs_row_kernel": err, "plan_s": plan_s, "stats": st}
            # free the plan before the next size (1.6 GB each at the scaled shape)
            h, _ = g._bplans.pop(min(bc, g.shape[1]))
            genmmrec_b200._lib.load().gmr_spmm_blocked_plan_destroy(h)
            print(name, mb, entry["blocked"]["%dMB" % mb], flush=True)
        res["cases"][name] = entry
    s = json.dumps(res, indent=1)
    print(s)
    if args.out:
        with open(args.out, "w") as f:
            f.write(s)


if __name__ == "__main__":
    main()
