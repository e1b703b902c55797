"""Top-K evaluator on the device.

Same contract as ``TopKEvaluator`` of GenMMRec/src/utils/topk_evaluator.py:36-120,299-313 --
``evaluate(batch_matrix_list, eval_data, is_test, idx) -> {'recall@5': ...}`` with keys
``'<metric>@<k>'`` and values rounded to 4 decimals -- but the hit matrix and the four metric
prefix sums come from one kernel (csrc/metrics.cu) instead of a Python membership loop plus numpy
(4.4 s of the reference's 5.1 s evaluation pass at the Baby shape).  The unrounded metric matrix of
the last call is kept in ``last_raw`` for parity checks.
"""
import os

import numpy as np
import torch

from .. import ops

topk_metrics = {m.lower(): m for m in ["Recall", "Recall2", "Precision", "NDCG", "MAP"]}
_DEVICE_ROWS = {"recall": 0, "ndcg": 1, "precision": 2, "map": 3}


class TopKEvaluator(object):
    def __init__(self, config):
        self.config = config
        self.metrics = config["metrics"]
        self.topk = config["topk"]
        self.save_recom_result = config["save_recommended_topk"]
        self._check_args()
        self.last_raw = None
        self.last_hit = None
        self.last_topk = None

    def _check_args(self):
        if isinstance(self.metrics, (str, list)):
            if isinstance(self.metrics, str):
                self.metrics = [self.metrics]
        else:
            raise TypeError("metrics must be str or list")
        for m in self.metrics:
            if m.lower() not in topk_metrics:
                raise ValueError("There is no user grouped topk metric named {}!".format(m))
        self.metrics = [m.lower() for m in self.metrics]
        if isinstance(self.topk, int):
            self.topk = [self.topk]
        if not isinstance(self.topk, list):
            raise TypeError("The topk must be a integer, list")
        for k in self.topk:
            if k <= 0:
                raise ValueError("topk must be a positive integer or a list of positive integers, but get `{}`".format(k))

    def metric_sums(self, topk_index, eval_data, keep_hit=False):
        """Device part: per-position SUMS over this loader's users of [recall, ndcg, precision, map]
        (float64 [4, K]) -- the quantity that is all-reduced when users are sharded across GPUs."""
        topk_index = topk_index.to(torch.int32).contiguous()
        sums, hit = ops.hits_metrics(topk_index, eval_data.gt_rowptr, eval_data.gt_items, return_hit=keep_hit)
        return sums, hit

    def finalize(self, sums, n_users, hit=None, pos_len_sum=None):
        """[4, K] sums + user count -> (rounded dict, raw [n_metrics, K])."""
        means = (sums / float(n_users)).cpu().numpy()
        rows = []
        for m in self.metrics:
            if m == "recall2":
                if hit is None:
                    raise ValueError("recall2 needs the hit matrix (keep_hit=True)")
                cum = torch.cumsum(hit.to(torch.float64), dim=1).sum(dim=0)
                rows.append((cum / float(pos_len_sum)).cpu().numpy())
            else:
                rows.append(means[_DEVICE_ROWS[m]])
        raw = np.stack(rows, axis=0)
        out = {}
        for m, value in zip(self.metrics, raw):
            for k in self.topk:
                out["{}@{}".format(m, k)] = round(float(value[k - 1]), 4)
        return out, raw

    def evaluate(self, batch_matrix_list, eval_data, is_test=False, idx=0):
        topk_index = batch_matrix_list if torch.is_tensor(batch_matrix_list) else torch.cat(batch_matrix_list, dim=0)
        assert topk_index.shape[0] == len(eval_data.get_eval_len_list())
        if self.save_recom_result and is_test:
            self._dump(topk_index, eval_data, idx)
        need_hit = "recall2" in self.metrics
        sums, hit = self.metric_sums(topk_index, eval_data, keep_hit=need_hit)
        out, raw = self.finalize(sums, topk_index.shape[0], hit, int(np.sum(eval_data.get_eval_len_list())))
        self.last_raw, self.last_hit, self.last_topk = raw, hit, topk_index
        return out

    def _dump(self, topk_index, eval_data, idx):
        """Tab-separated dump of the recommended ids (topk_evaluator.py:93-106)."""
        import time

        max_k = max(self.topk)
        dir_name = os.path.abspath(self.config["recommend_topk"])
        os.makedirs(dir_name, exist_ok=True)
        path = os.path.join(dir_name, "{}-{}-idx{}-top{}-{}.csv".format(
            self.config["model"], self.config["dataset"], idx, max_k, time.strftime("%b-%d-%Y-%H-%M-%S")))
        ids = topk_index.cpu().numpy().astype(np.int64)
        users = eval_data.get_eval_users().numpy().astype(np.int64)
        header = "\t".join(["id"] + ["top_" + str(i) for i in range(max_k)])
        np.savetxt(path, np.concatenate([users[:, None], ids], axis=1), fmt="%d", delimiter="\t", header=header,
                   comments="")

    def __str__(self):
        return "The TopK Evaluator Info:\n\tMetrics:[" + ", ".join(topk_metrics[m] for m in self.metrics) + \
               "], TopK:[" + ", ".join(map(str, self.topk)) + "]"
