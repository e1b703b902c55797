"""Evaluation loop of the hot path.

``Trainer.evaluate(eval_data, is_test=False, idx=0)`` keeps the signature and the result dict of
GenMMRec/src/common/trainer.py:369-388, but the per-batch pipeline
``full_sort_predict -> scores[mask] = -1e10 -> torch.topk -> .cpu() -> Python hit loop -> numpy``
becomes: one propagation (cached), ONE fused score + mask + top-K launch over all eval users (K2) and
one hit/metric launch (K4).  With ``torch.distributed`` initialised, eval users are sharded by rank
and only the [4, K] float64 metric sums are all-reduced.

``fit`` is the reference's generic training loop (common/trainer.py:144-332: one optimiser over
``model.calculate_loss``, LambdaLR schedule, evaluation every ``eval_step`` epochs, early stopping on the
validation metric) riding the same kernels: the SpMM backward is K1 on the transposed graph, the BPR
gathers are one fused kernel each way (csrc/train_ops.cu).  The diffusion / denoising pre-steps of the
DiffMM and GenRecV1 trainers stay out of scope: their graphs are attributes the caller sets.
"""
import itertools
import time

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
        # training loop state (common/trainer.py:74-121 of the reference)
        self.learner = config["learner"] or "adam"
        self.learning_rate = config["learning_rate"] or 1e-3
        self.epochs = config["epochs"] or 1
        self.eval_step = min(config["eval_step"] or 1, self.epochs)
        self.stopping_step = config["stopping_step"] or 0
        self.clip_grad_norm = config["clip_grad_norm"]
        self.valid_metric = (config["valid_metric"] or "Recall@20").lower()
        self.valid_metric_bigger = config["valid_metric_bigger"] is not False
        wd = config["weight_decay"]
        self.weight_decay = float(eval(wd) if isinstance(wd, str) else (wd or 0.0))
        self.req_training = config["req_training"] is not False
        self.start_epoch, self.cur_step = 0, 0
        self.train_loss_dict = {}
        self.optimizer = None
        self.lr_scheduler = None

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

    # ---- training (common/trainer.py:123-332) --------------------------------------------------------
    def _build_optimizer(self):
        params = [p for p in self.model.parameters() if p.requires_grad]
        name = self.learner.lower()
        opt = {"adam": torch.optim.Adam, "sgd": torch.optim.SGD, "adagrad": torch.optim.Adagrad,
               "rmsprop": torch.optim.RMSprop}.get(name, torch.optim.Adam)
        return opt(params, lr=self.learning_rate, weight_decay=self.weight_decay)

    def _train_epoch(self, train_data, epoch_idx, loss_func=None):
        """One pass over the train loader; returns (summed loss -- a tuple if the model returns loss parts --, per-batch
        losses).  A NaN loss returns the loss TENSOR, which ``fit`` takes as the signal to stop."""
        if not self.req_training:
            return 0.0, []
        self.model.train()
        loss_func = loss_func or self.model.calculate_loss
        total, batches = None, []
        for batch_idx, interaction in enumerate(train_data):
            self.optimizer.zero_grad()
            losses = loss_func(interaction)
            if isinstance(losses, tuple):
                loss = sum(losses)
                parts = tuple(x.item() for x in losses)
                total = parts if total is None else tuple(map(sum, zip(total, parts)))
            else:
                loss = losses
                total = loss.item() if total is None else total + loss.item()
            if torch.isnan(loss):
                return loss, torch.tensor(0.0)
            loss.backward()
            if self.clip_grad_norm:
                torch.nn.utils.clip_grad_norm_(self.model.parameters(), **self.clip_grad_norm)
            self.optimizer.step()
            batches.append(loss.detach())
        return total, batches

    def _valid_epoch(self, valid_data, is_test=False):
        result = self.evaluate(valid_data, is_test=is_test)
        return result[self.valid_metric], result

    @staticmethod
    def early_stopping(value, best, cur_step, max_step, bigger=True):
        """utils/utils.py:70-113 of the reference: (best, steps without improvement, stop?, improved?)."""
        improved = value > best if bigger else value < best
        if improved:
            return value, 0, False, True
        cur_step += 1
        return best, cur_step, cur_step > max_step, False

    def fit(self, train_data, valid_data=None, test_data=None, saved=False, verbose=False):
        """Train on ``train_data``; every ``eval_step`` epochs evaluate on ``valid_data`` (and ``test_data``) and stop
        early after ``stopping_step`` evaluations without improvement of ``valid_metric``.  Returns
        ``(best_valid_score, best_valid_result, best_test_upon_valid)`` like the reference."""
        if self.optimizer is None:
            self.optimizer = self._build_optimizer()
            sch = self.config["learning_rate_scheduler"] or [1.0, 50]
            self.lr_scheduler = torch.optim.lr_scheduler.LambdaLR(self.optimizer, lr_lambda=lambda e: sch[0] ** (e / sch[1]))
        for epoch_idx in range(self.start_epoch, self.epochs):
            t0 = time.time()
            self.model.pre_epoch_processing()
            train_loss, _ = self._train_epoch(train_data, epoch_idx)
            if torch.is_tensor(train_loss):   # NaN loss
                break
            self.lr_scheduler.step()
            self.train_loss_dict[epoch_idx] = sum(train_loss) if isinstance(train_loss, tuple) else train_loss
            self.model.post_epoch_processing()
            if verbose:
                print("epoch %d training [time: %.2fs, train loss: %.4f]" % (epoch_idx, time.time() - t0,
                                                                               self.train_loss_dict[epoch_idx]))
            if valid_data is not None and (epoch_idx + 1) % self.eval_step == 0:
                self.model.eval()
                valid_score, valid_result = self._valid_epoch(valid_data)
                self.best_valid_score, self.cur_step, stop, update = self.early_stopping(
                    valid_score, self.best_valid_score, self.cur_step, self.stopping_step, self.valid_metric_bigger)
                test_result = self._valid_epoch(test_data, is_test=True)[1] if test_data is not None else None
                if verbose:
                    print("epoch %d evaluating [valid_score: %f]" % (epoch_idx, valid_score))
                if update:
                    self.best_valid_result = valid_result
                    if test_result is not None:
                        self.best_test_upon_valid = test_result
                    if saved:
                        self._save_checkpoint(epoch_idx)
                if stop:
                    break
        return self.best_valid_score, self.best_valid_result, self.best_test_upon_valid

    def _save_checkpoint(self, epoch):
        import os

        d = self.config["checkpoint_dir"] or "saved"
        os.makedirs(d, exist_ok=True)
        torch.save({"epoch": epoch, "state_dict": self.model.state_dict(), "optimizer": self.optimizer.state_dict(),
                    "best_valid_score": self.best_valid_score},
                   os.path.join(d, "{}-{}.pth".format(self.config["model"], self.config["dataset"])))
