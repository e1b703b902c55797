"""Sanity sweep over the model families at the published Amazon shapes: every plugin evaluates through the fused
tensor-core path and through the fp32 path; top-K ids and metric dicts must agree (they are bit-identical by design).

    python tools/model_sweep.py [--shapes baby,sports,clothing] [--models DiffMM,GUME,GenRecV1,LD4MRec,VBPR,LightGCN]

Prints one JSON line per (model, shape) with timings; exits non-zero on any disagreement.  Diagnostics / soak test.
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
from genmmrec_b200.common.trainer import Trainer  # noqa: E402
from genmmrec_b200.workload import Workload  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="baby,sports,clothing")
    ap.add_argument("--models", default="DiffMM,GUME,GenRecV1,LD4MRec,VBPR,LightGCN")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    bad = 0
    for shape in args.shapes.split(","):
        for name in args.models.split(","):
            t0 = time.time()
            try:
                wl = Workload(name, shape, dev)
            except Exception as e:  # a model family that needs inputs this generator does not make
                print(json.dumps({"model": name, "shape": shape, "skipped": repr(e)[:200]}), flush=True)
                continue
            res = {}
            for prec in ("tc", "fp32"):
                wl.config["score_precision"] = prec
                wl.model.score_precision = prec
                tr = Trainer(wl.config, wl.model)
                wl.model.invalidate_cache()
                torch.cuda.synchronize()
                t1 = time.time()
                ids, _ = tr.topk_all(wl.valid)
                out = tr.evaluator.evaluate(ids, wl.valid)
                torch.cuda.synchronize()
                res[prec] = (ids, out, time.time() - t1)
            same_ids = bool(torch.equal(res["tc"][0], res["fp32"][0]))
            same_metrics = res["tc"][1] == res["fp32"][1]
            d = int(wl.model.eval_factors(wl.valid.eval_u)[2].shape[1])
            line = {"model": name, "shape": shape, "users": wl.n_eval_users, "items": wl.n_items, "d": d,
                    "tc_supported": ops.tc_supported(d, max(wl.config["topk"])), "ids_identical": same_ids,
                    "metrics_identical": same_metrics, "recall@20": res["tc"][1].get("recall@20"),
                    "eval_s_tc": round(res["tc"][2], 4), "eval_s_fp32": round(res["fp32"][2], 4), "setup_s": round(time.time() - t0, 2)}
            print(json.dumps(line), flush=True)
            bad += 0 if (same_ids and same_metrics) else 1
            del wl
            torch.cuda.empty_cache()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
