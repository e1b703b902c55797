"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_check.py [--shape baby]

Every rank builds the same model replica (same seeds; --model DiffMM | GenRecV1 | LightGCN), runs the row-sharded
propagation (row blocks stored into every replica over NVLink peer memory) + user-block sharded evaluation, and
compares with its own single-GPU result: embeddings of its blocks, top-K rows, all-reduced metric sums."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="baby")
    ap.add_argument("--model", default="DiffMM", choices=["DiffMM", "GenRecV1", "LightGCN"])
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--mode", default="rows", choices=["rows", "cols"], help="DiffMM: adjacency rows or embedding columns sharded")
    args = ap.parse_args()
    from genmmrec_b200 import dist as gd, ops
    from genmmrec_b200.common.trainer import Trainer
    from genmmrec_b200.workload import Workload

    rank, world, local = gd.world_info()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    wl = Workload(args.model, args.shape, dev, overrides={"n_layers": args.layers})
    model = wl.model
    trainer = Trainer(wl.config, model)
    k = max(wl.config["topk"])
    with torch.no_grad():
        ue, ie = model.propagate()
        ids_ref, _ = trainer.topk_all(wl.valid)
        sums_ref, _ = trainer.evaluator.metric_sums(ids_ref, wl.valid)
        sh = {"DiffMM": gd.ColShardedDiffMM if args.mode == "cols" else gd.ShardedDiffMM, "GenRecV1": gd.sharded_genrecv1, "LightGCN": gd.sharded_lightgcn}[args.model](model)
        for _ in range(2):  # twice: buffers are reused across calls
            su, items = sh.eval_factors()
        part = gd.shard_eval_by_user_block(wl.valid, sh.u0, sh.u1)
        ids, _ = ops.score_mask_topk(su.contiguous(), items.contiguous(), k, users=part.eval_u,
                                     mask_rowptr=part.mask_rowptr, mask_items=part.mask_items, precision="auto")
        sums, _ = trainer.evaluator.metric_sums(ids, part)
        dist.all_reduce(sums)
    torch.cuda.synchronize()
    # the sharded dataflow runs the same SpMM passes on row blocks, so agreement is bitwise in practice; the bound
    # checked is fp32 rounding (and top-K up to near-ties)
    tol_u = float((su - ue[sh.u0:sh.u1]).abs().max() / ue.abs().max())
    tol_i = float((items - ie).abs().max() / ie.abs().max())
    ok_u, ok_i = tol_u < 1e-5, tol_i < 1e-5
    same_rows = float((ids == ids_ref[part.positions]).all(dim=1).float().mean())
    ok_ids = same_rows > 0.98
    ok_m = bool((sums - sums_ref).abs().max() < 1e-3 * sums_ref.abs().max())
    if getattr(sh, "barrier", None) is not None:
        sh.barrier.check()
    res = {"rank": rank, "world": world, "mode": args.mode if args.model == "DiffMM" else "rows",
           "bit_identical": bool(torch.equal(su, ue[sh.u0:sh.u1]) and torch.equal(items, ie)), "users_block": [sh.u0, sh.u1], "items_block": [sh.i0, sh.i1],
           "user_rows_rel_err": tol_u, "gathered_items_rel_err": tol_i, "topk_rows_identical": same_rows,
           "metric_sums_match": ok_m}
    sys.stdout.flush()
    os.write(1, (json.dumps(res) + "\n").encode())   # one write per rank: lines of different ranks never interleave
    sh.close()
    dist.barrier()
    dist.destroy_process_group()
    if not (ok_u and ok_i and ok_ids and ok_m):
        sys.exit(1)


if __name__ == "__main__":
    main()
