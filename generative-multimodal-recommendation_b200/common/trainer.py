"""Evaluation loop of the hot path.

``Trainer.evaluate(eval_data, is_test=False, idx=0)`` keeps the signature and the result dict of
GenMMRec/src/common/trainer.py:369-388, but the per-batch pipeline
``full_sort_predict -> scores[mask] = -1e10 -> torch.topk -> .cpu() -> Python hit loop -> numpy``
becomes: one propagation (cached), ONE fused score + mask + top-K launch over all eval users (K2) and
one hit/metric launch (K4).  With ``torch.distributed`` initialised, eval users are sharded by rank
and only the [4, K] float64 metric sums are all-reduced.

Training (``fit``) is outside the hot path of this build; ``calculate_loss`` of the models is
differentiable through ``ops.spmm`` so an external loop can still train them.
"""
import itertools

import torch

from ..utils.topk_evaluator import TopKEvaluator
from .. import ops


class Trainer(object):
    def __init__(self, config, model, mg=False):
        self.config = config
        self.model = model
        self.device = config["device"]
        self.test_batch_size = config["eval_batch_size"]
        self.evaluator = TopKEvaluator(config)
        self.eval_mode = config["eval_mode"] or "fused"  # fused | batched (reference-shaped loop)
        keys = ["{}@{}".format(m.lower(), k) for m, k in itertools.product(config["metrics"], config["topk"])]
        self.best_valid_score = -1
        self.best_valid_result = dict.fromkeys(keys, 0.0)
        self.best_test_upon_valid = dict.fromkeys(keys, 0.0)

    @torch.no_grad()
    def topk_all(self, eval_data, return_scores=False):
        """Top-max(topk) ids of every eval user: int32 [n_eval, K] (and fp32 scores)."""
        self.model.eval()
        k = max(self.config["topk"])
        if self.eval_mode == "batched":
            outs = [self.model.full_sort_topk(batch, k) for batch in eval_data]
            return torch.cat(outs, dim=0).to(torch.int32), None
        try:
            self.model.cached_propagate()  # before the loader's tensors are touched: an in-flight upload overlaps it
        except NotImplementedError:
            pass                           # models that score through eval_factors only (LD4MRec)
        eu, rows, ei, bias = self.model.eval_factors(eval_data.eval_u)
        ids, sc = ops.score_mask_topk(eu.contiguous(), ei.contiguous(), k, users=rows, bias=bias,
                                      mask_rowptr=eval_data.mask_rowptr, mask_items=eval_data.mask_items,
                                      precision=self.model.score_precision, return_scores=return_scores)
        return ids, sc

    @torch.no_grad()
    def evaluate(self, eval_data, is_test=False, idx=0):
        ids, _ = self.topk_all(eval_data)
        return self.evaluator.evaluate(ids, eval_data, is_test=is_test, idx=idx)

    def graphed(self, fn):
        """Capture ``fn`` (a no-argument callable that only enqueues GPU work, e.g. one evaluation pass) into a CUDA
        graph and return ``replay() -> whatever fn returned at capture`` (static tensors, refreshed by every replay).
        Evaluation at the Amazon shapes is launch-bound (~60 launches for <1 ms of GPU work): one graph launch
        replaces them.  ``fn`` must have run at least once before (workspaces, plans and caches are created
        eagerly) and must not synchronise."""
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()                                   # one more eager run on the capture side stream
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = fn()

        def replay():
            graph.replay()
            return out

        replay.graph = graph
        ops.note_graph_captured()   # workspaces this graph points into must outlive it (ops._ws)
        return replay

    def fit(self, *args, **kwargs):
        raise NotImplementedError("training loops are outside the hot path this package accelerates; use "
                                  "model.calculate_loss with your own optimiser loop")
