"""Time of the kNN item-item graph construction (cosine similarity + top-10 + weighted normalisation,
GenMMRec/src/utils/utils.py:147-197) without the dense I x I matrix: tensor-core builder (K-chunked split-bf16 tcgen05,
exact re-score) against the CUDA-core fp32 builder.  Diagnostics only.

    python tools/knn_bench.py --items 100000,500000 --dims 4096,384 [--fp32-items 100000]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import genmmrec_b200  # noqa: F401,E402
from genmmrec_b200 import graph as gb, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", default="100000,500000")
    ap.add_argument("--dims", default="4096,384")
    ap.add_argument("--fp32-items", type=int, default=100000, help="largest item count the fp32 builder is timed at")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    res = []
    for d in [int(x) for x in args.dims.split(",")]:
        for n in [int(x) for x in args.items.split(",")]:
            g = torch.Generator(device=dev).manual_seed(7)
            feat = torch.randn((n, d), device=dev, generator=g)
            if d == 4096:
                feat.clamp_(min=0)          # CNN features are non-negative
            e = {"items": n, "d": d, "dense_flops": 2.0 * n * n * d}
            for prec in ("tc_split", "fp32"):
                if prec == "fp32" and n > args.fp32_items:
                    continue
                torch.cuda.synchronize()
                t0 = time.time()
                idx, w, _ = gb.knn_graph_fused(feat, 10, precision=prec)
                torch.cuda.synchronize()
                dt = time.time() - t0
                e[prec + "_s"] = dt
                e[prec + "_dense_TFLOPs"] = e["dense_flops"] / dt / 1e12
                if prec == "tc_split":
                    e["tc_fp32_redo_rows"] = ops.last_tc_fallback_rows()
                    ref = idx
                else:
                    e["same_edges_as_tc"] = bool(torch.equal(ref, idx))
                del idx, w
                torch.cuda.empty_cache()
            print(json.dumps(e), flush=True)
            res.append(e)
            del feat
            torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
