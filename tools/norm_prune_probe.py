"""Probe: how much of the catalogue can the Cauchy-Schwarz bound |u| |e_i| < s_K rule out per user?

For a sample of eval users of a bench workload: exact top-K scores by a dense torch matmul, then the
fraction of items whose norm bound |u| |e_i| still reaches the K-th best score (those are the only
items a norm-ordered sweep would have to score).  Diagnostics only; nothing here is on the product path.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import genmmrec_b200  # noqa: F401,E402
from genmmrec_b200.workload import Workload  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="scaled")
    ap.add_argument("--model", default="DiffMM")
    ap.add_argument("--sample", type=int, default=4096)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    wl = Workload(args.model, args.workload, dev)
    k = max(wl.config["topk"])
    with torch.no_grad():
        eu, rows, ei, bias = wl.model.eval_factors(wl.valid.eval_u)
    rows = rows[:args.sample]
    u = eu[rows]
    s = u @ ei.T
    # train-history mask
    mrp, mit = wl.valid.mask_rowptr, wl.valid.mask_items
    lens = (mrp[1:args.sample + 1] - mrp[:args.sample])
    rr = torch.repeat_interleave(torch.arange(rows.numel(), device=dev), lens)
    s[rr, mit[:int(mrp[args.sample])].long()] = -1e10
    top = torch.topk(s, k, dim=1).values
    sk = top[:, -1]
    un = u.norm(dim=1)
    inorm = ei.norm(dim=1)
    need = sk / un                                  # items with |e_i| < need cannot enter the top K
    sorted_norm = torch.sort(inorm).values
    n_ge = ei.shape[0] - torch.searchsorted(sorted_norm, need.clamp(min=0))
    frac = n_ge.double() / ei.shape[0]
    frac[sk <= 0] = 1.0
    # per group of 256 consecutive users the sweep must cover the max
    g = frac[: (frac.numel() // 256) * 256].reshape(-1, 256).max(dim=1).values
    q = torch.tensor([0.1, 0.5, 0.9, 0.99], device=dev, dtype=torch.float64)
    out = {"workload": args.workload, "users": int(rows.numel()), "items": int(ei.shape[0]), "k": k,
           "frac_items_needed_per_user": {"mean": float(frac.mean()), "quantiles_10_50_90_99": torch.quantile(frac, q).tolist(),
                                          "max": float(frac.max())},
           "frac_items_needed_per_256_group": {"mean": float(g.mean()), "max": float(g.max())},
           "sk_over_unorm": {"median": float(need.median()), "min": float(need.min())},
           "item_norm_quantiles_50_90_99_999": torch.quantile(inorm.double(), torch.tensor([0.5, 0.9, 0.99, 0.999], device=dev, dtype=torch.float64)).tolist()}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
