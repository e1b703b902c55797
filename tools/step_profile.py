"""Kernel-level time breakdown of one hot-path step (propagation + fused eval) with torch.profiler.

    python tools/step_profile.py --workload scaled [--steps 3]

Prints the CUDA kernels of the profiled steps sorted by total device time (diagnostics; bench.py is the
contract benchmark).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import genmmrec_b200  # noqa: F401,E402
from genmmrec_b200.common.trainer import Trainer  # noqa: E402
from genmmrec_b200.workload import Workload  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="scaled")
    ap.add_argument("--model", default="DiffMM")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--top", type=int, default=40)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    wl = Workload(args.model, args.workload, dev)
    trainer = Trainer(wl.config, wl.model)

    def step():
        wl.model.invalidate_cache()
        ids, _ = trainer.topk_all(wl.valid)
        sums, _ = trainer.evaluator.metric_sums(ids, wl.valid)
        return sums

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = getattr(e, "cuda_time_total", 0)
        if t and e.device_type is not None and "cuda" in str(e.device_type).lower():
            rows.append((t / args.steps / 1e3, e.count / args.steps, e.key[:110]))
    rows.sort(reverse=True)
    total = sum(r[0] for r in rows)
    out = {"workload": args.workload, "steps": args.steps, "device_ms_per_step": total,
           "kernels": [{"ms_per_step": round(ms, 4), "calls_per_step": c, "name": n} for ms, c, n in rows[:args.top]]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
