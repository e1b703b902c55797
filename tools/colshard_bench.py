"""Single-GPU timing of the column-sharded propagation's local work (tools; not a bench line):

    python tools/colshard_bench.py [--shape scaled] [--world 8] [--out gpurun_out/colshard.json]

Times the narrow SpMM passes a rank of a `world`-GPU column-sharded run executes (whole graph, d / world columns) next to
the wide single-GPU passes, and the three phases of one emulated rank (peer stores land in local buffers, so the numbers
are the compute side only).  Also checks the emulated result against the single-GPU propagation bit for bit."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="scaled")
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--out", default="")
    ap.add_argument("--no-emulate", action="store_true")
    ap.add_argument("--profile", action="store_true", help="one launch of each SELL pass between cudaProfilerStart/Stop (ncu --profile-from-start off)")
    args = ap.parse_args()
    from genmmrec_b200 import dist as gd, ops
    from genmmrec_b200.workload import Workload

    dev = torch.device("cuda:0")
    wl = Workload("DiffMM", args.shape, dev)
    m = wl.model
    nu, ni, d, W = m.n_users, m.n_items, m.latdim, args.world
    dc = d // W
    adj = m.norm_adj
    res = {"shape": args.shape, "world": W, "dc": dc, "nnz_full": adj.full.nnz}
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    x64 = torch.randn(nu + ni, d, device=dev, generator=g)
    x128i = torch.randn(ni, 2 * d, device=dev, generator=g)
    with torch.no_grad():
        res["wide_full_64_ms"] = timeit(lambda: ops.spmm_raw(adj.full, x64))
        res["wide_ui_128_ms"] = timeit(lambda: ops.spmm_raw(adj.ui, x128i))
        res["wide_iu_64_ms"] = timeit(lambda: ops.spmm_raw(adj.iu, x64[:nu]))
        xc = x64[:, :dc].contiguous()
        xci = x128i[:, :2 * dc].contiguous()
        yc = torch.empty(nu + ni, dc, device=dev)
        yu = torch.empty(nu, 2 * dc, device=dev)
        yi = torch.empty(ni, dc, device=dev)
        for sell, tag in ((False, "narrow"), (True, "sell")):
            res[tag + "_full_ms"] = timeit(lambda: ops.spmm_narrow(adj.full, xc, chains=2, out=yc, sell=sell))
            res[tag + "_ui_2dc_ms"] = timeit(lambda: ops.spmm_narrow(adj.ui, xci, chains=1, out=yu, sell=sell))
            res[tag + "_iu_ms"] = timeit(lambda: ops.spmm_narrow(adj.iu, xc[:nu], chains=2, out=yi, sell=sell))
        if args.profile:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStart()
            ops.spmm_narrow(adj.full, xc, chains=2, out=yc, sell=True)
            ops.spmm_narrow(adj.ui, xci, chains=1, out=yu, sell=True)
            ops.spmm_narrow(adj.iu, xc[:nu], chains=2, out=yi, sell=True)
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
        wide = ops.spmm_raw(adj.full, x64)
        res["narrow_full_equals_wide"] = bool(torch.equal(ops.spmm_narrow(adj.full, xc, chains=2), wide[:, :dc]))
        del wide
        if not args.no_emulate:
            ue, ie = m.propagate()
            ranks = gd.ColShardedDiffMM.emulate(m, W)
            out = gd.ColShardedDiffMM.emulated_eval_factors(ranks)
            res["emulated_bit_identical"] = bool(all(torch.equal(items, ie) and torch.equal(su, ue[r.u0:r.u1])
                                                     for r, (su, items) in zip(ranks, out)))
            r0 = ranks[0]
            for ph in ("_p0", "_p1", "_p2", "_p3"):
                res["rank0" + ph + "_ms"] = timeit(getattr(r0, ph), iters=5, warm=1)
            ops.PROFILE = []
            for ph in ("_p0", "_p1", "_p2", "_p3"):
                getattr(r0, ph)()
            torch.cuda.synchronize()
            res["rank0_ops"] = [(n, meta.get("kernel", ""), meta.get("d", 0), round(a.elapsed_time(b), 4)) for n, meta, a, b in ops.PROFILE]
            ops.PROFILE = None
            res["single_gpu_propagate_ms"] = timeit(lambda: (m.invalidate_cache(), m.propagate()), iters=3, warm=1)
    print(json.dumps(res, indent=1))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
