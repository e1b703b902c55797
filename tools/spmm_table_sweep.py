"""Probe: SpMM gather rate of K1 (row-centric) and K1b (nonzero-centric, one column block) as a function of the size of
the gathered table, on a uniform random graph (1M rows x 48 nonzeros, D = 64).  Separates what the kernels can issue
from what the L2 can serve.  Diagnostics only.

    python tools/spmm_table_sweep.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import genmmrec_b200  # noqa: F401,E402
from genmmrec_b200 import ops  # noqa: E402


def time_fn(fn, iters=5, warmup=2):
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
    dev = torch.device("cuda", 0)
    n_rows, deg, d = 1_000_000, 48, 64
    res = {}
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    for mb in (8, 32, 48, 64, 96, 128, 256):
        n_cols = (mb << 20) // (d * 4)
        col = torch.randint(0, n_cols, (n_rows * deg,), device=dev, generator=g, dtype=torch.int32)
        val = torch.rand(n_rows * deg, device=dev, generator=g)
        rowptr = (torch.arange(n_rows + 1, device=dev) * deg).to(torch.int32)
        a = ops.GraphCSR(rowptr, col, val, (n_rows, n_cols))
        x = torch.randn(n_cols, d, device=dev)
        out = torch.empty(n_rows, d, device=dev)
        os.environ["GMR_SPMM_BLOCKED"] = "0"
        row_ms = time_fn(lambda: ops.spmm_raw(a, x, out=out))
        flat_ms = time_fn(lambda: ops.spmm_blocked(a, x, n_cols, out=out))
        gb = a.nnz * (8 + 4 * d) / 1e6
        res["%dMB" % mb] = {"row_ms": row_ms, "row_gather_GBs": gb / row_ms, "flat_ms": flat_ms, "flat_gather_GBs": gb / flat_ms}
        print(mb, res["%dMB" % mb], flush=True)
        del a
    print(json.dumps(res))


if __name__ == "__main__":
    main()
