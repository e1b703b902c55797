"""Probe: does column-blocking the CSR (so that each pass gathers from an L2-resident slice of X) pay for its extra
passes over Y?  Uses the bench workload's own graphs.  Diagnostics only.

    python tools/spmm_block_probe.py --workload scaled
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import genmmrec_b200  # noqa: F401,E402
from genmmrec_b200 import ops  # noqa: E402
from genmmrec_b200.workload import Workload  # noqa: E402


def time_fn(fn, iters=8, warmup=2):
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


def col_blocks(g, n_blocks):
    """Split a GraphCSR into n_blocks graphs over contiguous column ranges (columns rebased)."""
    dev = g.device
    counts = (g.rowptr[1:] - g.rowptr[:-1]).to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(g.shape[0], device=dev), counts)
    col = g.col.to(torch.int64)
    out = []
    for b in range(n_blocks):
        lo, hi = g.shape[1] * b // n_blocks, g.shape[1] * (b + 1) // n_blocks
        sel = (col >= lo) & (col < hi)
        sub = ops.GraphCSR.from_coo(torch.stack([rows[sel], col[sel] - lo]), g.val[sel], (g.shape[0], hi - lo), dev)
        sub.plan
        out.append((sub, lo, hi))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="scaled")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    wl = Workload("DiffMM", args.workload, dev)
    adj = wl.model.norm_adj
    nu, ni = wl.n_users, wl.n_items
    res = {}
    for name, g, d in (("iu_d64", adj.iu, 64), ("ui_d64", adj.ui, 64), ("ui_d128", adj.ui, 128), ("full_d64", adj.full, 64)):
        x = torch.randn(g.shape[1], d, device=dev)
        y = torch.empty(g.shape[0], d, device=dev)
        r = {"mono_ms": time_fn(lambda: ops.spmm_raw(g, x, out=y)), "table_MB": g.shape[1] * d * 4 / 1e6,
             "gather_GB": g.nnz * d * 4 / 1e9}
        ref = ops.spmm_raw(g, x)
        for nb in (2, 4, 8):
            blocks = col_blocks(g, nb)

            def run():
                for j, (sub, lo, hi) in enumerate(blocks):
                    ops.spmm_raw(sub, x[lo:hi], out=y, beta=0.0 if j == 0 else 1.0)
            r["blocks%d_ms" % nb] = time_fn(run)
            run()
            r["blocks%d_relerr" % nb] = float((y - ref).abs().max() / ref.abs().max())
            del blocks
        if d == 128:
            y2 = torch.empty(g.shape[0], d, device=dev)

            def run2():
                ops.spmm_raw(g, x[:, :64], out=y2[:, :64])
                ops.spmm_raw(g, x[:, 64:], out=y2[:, 64:])
            r["two_d64_slices_ms"] = time_fn(run2)
        res[name] = r
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
