"""Host-side sharding logic under real world_size-2 and -3 process groups (gloo, CPU): block bounds, the
padded uneven all-gather, user-block sharding of the eval loader and the all-reduced metric sums."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO


def _worker(rank, world, port, out):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from genmmrec_b200 import dist as gd, synth
        from genmmrec_b200.utils.configurator import Config
        from genmmrec_b200.utils.dataloader import EvalDataLoader
        from genmmrec_b200.utils.dataset import RecDataset
        from oracle import c_api

        # uneven all-gather of row blocks
        sizes = [3, 5, 2][:world]
        local = torch.full((sizes[rank], 4), float(rank + 1))
        full = gd.all_gather_rows(local, sizes)
        want_rows = torch.cat([torch.full((n, 4), float(g + 1)) for g, n in enumerate(sizes)])
        assert full.shape == want_rows.shape and torch.equal(full, want_rows)

        # nnz-balanced bounds cover everything, monotone, balanced within one row
        deg = torch.tensor(np.random.default_rng(0).zipf(1.5, size=1000).clip(max=500))
        rowptr = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(deg, 0)])
        b = gd.nnz_balanced_bounds(rowptr, world)
        assert b[0] == 0 and b[-1] == 1000 and all(b[i] <= b[i + 1] for i in range(world))
        per = [int(rowptr[b[g + 1]] - rowptr[b[g]]) for g in range(world)]
        assert max(per) - min(per) <= 2 * int(deg.max())
        assert gd.block_bounds(10, 3) == [0, 3, 6, 10]

        # user-block sharded evaluation == unsharded evaluation (metric SUMS are additive)
        cfg = Config("LightGCN", "toy", {"device": "cpu", "is_multimodal_model": False})
        u, i, lab = synth.make_interactions(300, 120, 3600)
        tr, va, _ = RecDataset.from_arrays(cfg, u, i, lab, 300, 120).split()
        loader = EvalDataLoader(cfg, va, additional_dataset=tr, batch_size=64)
        ub = gd.block_bounds(300, world)
        sh = gd.shard_eval_by_user_block(loader, ub[rank], ub[rank + 1])
        n_local = torch.tensor([sh.eval_u.numel()])
        dist.all_reduce(n_local)
        assert int(n_local) == loader.eval_u.numel()
        pos = sh.positions.numpy()
        assert np.array_equal(sh.eval_u.numpy() + ub[rank], loader.eval_u.numpy()[pos])
        rp, it = loader.mask_rowptr.numpy(), loader.mask_items.numpy()
        srp, sit = sh.mask_rowptr.numpy(), sh.mask_items.numpy()
        for j in (0, len(pos) // 2, len(pos) - 1):
            assert np.array_equal(sit[srp[j]:srp[j + 1]], it[rp[pos[j]]:rp[pos[j] + 1]])
        # fake top-K ids; per-rank oracle metric sums all-reduce to the global sums
        rng = np.random.default_rng(1)
        topk = np.stack([rng.choice(120, size=20, replace=False) for _ in range(loader.eval_u.numel())]).astype(np.int32)
        hit_all = c_api.hits(topk, loader.gt_rowptr.numpy(), loader.gt_items.numpy())
        m_all = c_api.metrics(hit_all, loader.eval_len_list)
        hit_loc = c_api.hits(topk[pos], sh.gt_rowptr.numpy(), sh.gt_items.numpy())
        m_loc = c_api.metrics(hit_loc, sh.eval_len_list)
        sums = torch.tensor(np.stack([m_loc[k] for k in ("recall", "ndcg", "precision", "map")]) * len(pos))
        dist.all_reduce(sums)
        want = np.stack([m_all[k] for k in ("recall", "ndcg", "precision", "map")])
        assert np.abs(sums.numpy() / loader.eval_u.numel() - want).max() < 1e-12
        # dist.evaluate_sharded: the same composition with the oracle standing in for the two kernels; the dict must
        # equal the unsharded one on every rank
        from genmmrec_b200 import ops
        from genmmrec_b200.utils.topk_evaluator import TopKEvaluator

        def fake_score(ue, ie, kk, users=None, mask_rowptr=None, mask_items=None, precision="tc", return_scores=False):
            ids_, sc_ = c_api.score_mask_topk(ue.numpy(), users.numpy(), ie.numpy(), None, mask_rowptr.numpy(),
                                              mask_items.numpy(), kk)
            return torch.from_numpy(ids_), None

        def fake_hits(topk_, gt_rowptr, gt_items, return_hit=False):
            h = c_api.hits(topk_.numpy(), gt_rowptr.numpy(), gt_items.numpy())
            m = c_api.metrics(h, np.diff(gt_rowptr.numpy()))
            return torch.from_numpy(np.stack([m[q] for q in ("recall", "ndcg", "precision", "map")]) * topk_.shape[0]), None

        ops.score_mask_topk, ops.hits_metrics = fake_score, fake_hits
        g = torch.Generator().manual_seed(3)
        ue_all, ie_all = torch.randn(300, 16, generator=g), torch.randn(120, 16, generator=g)

        class _Factors:
            def __init__(self, ue, ie):
                self.ue, self.ie = ue, ie

            def eval_factors(self):
                return self.ue, self.ie

        ev = TopKEvaluator({"metrics": ["Recall", "NDCG", "Precision", "MAP"], "topk": [5, 10, 20],
                            "save_recommended_topk": False})
        got, raw, ids_loc = gd.evaluate_sharded(_Factors(ue_all[ub[rank]:ub[rank + 1]], ie_all), ev, sh, precision="fp32")
        ids_all, _ = fake_score(ue_all, ie_all, 20, users=loader.eval_u, mask_rowptr=loader.mask_rowptr,
                                mask_items=loader.mask_items)
        assert torch.equal(ids_loc, ids_all[sh.positions])
        want_dict = ev.evaluate(ids_all, loader)
        assert got == want_dict and np.abs(raw - ev.last_raw).max() < 1e-12
        out.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        import traceback
        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("world", [2, 3])
def test_sharding_logic(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() + 7 * world) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", "rank %d failed:\n%s" % (rank, msg)
