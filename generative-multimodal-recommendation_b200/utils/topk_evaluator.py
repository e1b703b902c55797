"""Top-K evaluator on the device.

Same contract as ``TopKEvaluator`` of GenMMRec/src/utils/topk_evaluator.py:36-120,299-313 --
``evaluate(batch_matrix_list, eval_data, is_test, idx) -> {'recall@5': ...}`` with keys
``'<metric>@<k>'`` and values rounded to 4 decimals -- but the hit matrix and the four metric
prefix sums come from one kernel (csrc/metrics.cu) instead of a Python membership loop plus numpy
(4.4 s of the reference's 5.1 s evaluation pass at the Baby shape).  The unrounded metric matrix of
the last call is kept in ``last_raw`` for parity checks.

With ``is_test=True`` the reference adds (topk_evaluator.py:123-270) Pop/Niche and Cold/Warm group metrics and the
Coverage / Gini / Gini2 / Coverage2 / Tail% diversity numbers.  Here the group metrics are the same metrics kernel run on
a filtered ground-truth CSR (popular or niche items only; warm or cold users only) built with device index ops instead of
per-user Python list comprehensions, and the diversity numbers come from one device ``bincount`` per cut-off whose
[n_items] result is reduced on the host with the reference's exact numpy expressions (integer counts, so the 4-decimal
dict is identical).
"""
import os

import numpy as np
import torch

from .. import ops

topk_metrics = {m.lower(): m for m in ["Recall", "Recall2", "Precision", "NDCG", "MAP"]}
_DEVICE_ROWS = {"recall": 0, "ndcg": 1, "precision": 2, "map": 3}


def _cfg_get(config, key):
    """``config[key] if key in config else None`` for both the Config class and plain dicts."""
    try:
        return config[key]
    except KeyError:
        return None


def _gini_active(counts):
    """``cal_gini`` of topk_evaluator.py:19-33 (Lorenz-curve Gini over the items that were recommended at least once,
    with one zero appended), same operation order; ``np.trapz`` written out so it does not depend on the numpy version."""
    cum = np.cumsum(np.sort(np.append(counts, 0)))
    x = np.array(range(0, len(cum))) / (len(cum) - 1)
    y = cum / cum[-1]
    b = (np.diff(x) * (y[1:] + y[:-1]) / 2.0).sum(-1)
    a = 0.5 - b
    return a / (a + b)


class TopKEvaluator(object):
    def __init__(self, config):
        self.config = config
        self.metrics = config["metrics"]
        self.topk = config["topk"]
        self.save_recom_result = config["save_recommended_topk"]
        self.pop_items = _cfg_get(config, "pop_items")    # popular group, a set of item ids (quick_start.py:62-81)
        self.warm_users = _cfg_get(config, "warm_users")  # users with > 5 train interactions (quick_start.py:85-92)
        self._pop_mask = None
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
        if is_test:
            self._group_metrics(out, topk_index, eval_data)
            self._diversity_metrics(out, topk_index, eval_data)
        return out

    # ---- is_test extras (topk_evaluator.py:123-270) -------------------------------------------------
    def _pop_mask_np(self, item_num):
        """bool [item_num], True for popular items (ids outside the catalogue are ignored, topk_evaluator.py:217-222)."""
        if self._pop_mask is None or self._pop_mask.shape[0] != item_num:
            m = np.zeros(item_num, dtype=bool)
            m[np.fromiter((i for i in self.pop_items if 0 <= i < item_num), dtype=np.int64)] = True
            self._pop_mask = m
        return self._pop_mask

    def _subset_metrics(self, topk_i32, eval_data, prefix, out, item_keep=None, user_sel=None):
        """Metrics over a sub-population: ground truth restricted to the entries flagged in ``item_keep`` (users left
        without any are dropped, as in the reference's ``if len(gt_pop) > 0``) or to the users flagged in ``user_sel``."""
        rowptr, items = eval_data.gt_rowptr, eval_data.gt_items
        dev = rowptr.device
        n_all = rowptr.numel() - 1
        lens = rowptr[1:] - rowptr[:-1]
        rows = torch.repeat_interleave(torch.arange(n_all, device=dev), lens)
        keep = item_keep if item_keep is not None else torch.ones(items.numel(), dtype=torch.bool, device=dev)
        if user_sel is not None:
            keep = keep & user_sel[rows]
        new_lens = torch.bincount(rows[keep], minlength=n_all)
        users = user_sel if user_sel is not None else new_lens > 0
        n = int(users.sum())
        if n == 0:
            return
        sub_lens = new_lens[users]
        sub_ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        sub_ptr[1:] = torch.cumsum(sub_lens, dim=0)
        need_hit = "recall2" in self.metrics
        sums, hit = ops.hits_metrics(topk_i32[users].contiguous(), sub_ptr, items[keep].contiguous(), return_hit=need_hit)
        res, _ = self.finalize(sums, n, hit, int(sub_lens.sum()))
        for m in self.metrics:
            for k in self.topk:
                out["{}_{}@{}".format(prefix, topk_metrics.get(m, m), k)] = res["{}@{}".format(m, k)]

    def _group_metrics(self, out, topk_index, eval_data):
        if self.pop_items is None and self.warm_users is None:
            return
        dev = eval_data.gt_rowptr.device
        topk_i32 = topk_index.to(device=dev, dtype=torch.int32)
        if self.pop_items is not None:
            pop_mask = torch.from_numpy(self._pop_mask_np(int(eval_data.dataset.item_num))).to(dev)
            is_pop = pop_mask[eval_data.gt_items.long()]
            self._subset_metrics(topk_i32, eval_data, "Pop", out, item_keep=is_pop)
            self._subset_metrics(topk_i32, eval_data, "Niche", out, item_keep=~is_pop)
        if self.warm_users is not None:
            eval_users = eval_data.get_eval_users()
            eval_users = eval_users.cpu().numpy() if torch.is_tensor(eval_users) else np.asarray(eval_users)
            warm_ids = np.fromiter(self.warm_users, dtype=np.int64, count=len(self.warm_users))
            is_warm = torch.from_numpy(np.isin(eval_users, warm_ids)).to(dev)
            if bool((~is_warm).any()):
                self._subset_metrics(topk_i32, eval_data, "Cold", out, user_sel=~is_warm)
            if bool(is_warm.any()):
                self._subset_metrics(topk_i32, eval_data, "Warm", out, user_sel=is_warm)

    def _diversity_metrics(self, out, topk_index, eval_data):
        item_num = int(eval_data.dataset.item_num)
        pop_mask = self._pop_mask_np(item_num) if self.pop_items is not None else None
        ids = topk_index.long()
        for k in self.topk:
            # the only pass over the [U, k] ids runs on the device; everything below is O(n_items) host arithmetic on
            # exact integer counts, written as the reference writes it
            rec_count = torch.bincount(ids[:, :k].reshape(-1), minlength=item_num).cpu().numpy()
            n_rec = int(ids.shape[0]) * k
            out["Coverage@{}".format(k)] = round(np.count_nonzero(rec_count) / item_num, 4)
            sorted_counts = np.sort(rec_count)
            n = item_num
            sum_counts = np.cumsum(sorted_counts)[-1]
            if sum_counts > 0:
                index = np.arange(1, n + 1)
                gini = (2 * np.sum(index * sorted_counts)) / (n * sum_counts) - (n + 1) / n
                out["Gini@{}".format(k)] = round(gini, 4)
            else:
                out["Gini@{}".format(k)] = 0.0
            active = rec_count[rec_count > 0]
            if len(active) > 0:
                out["Gini2@{}".format(k)] = round(_gini_active(active), 4)
                out["Coverage2@{}".format(k)] = round(len(active) / item_num, 4)
            else:
                out["Gini2@{}".format(k)] = 0.0
                out["Coverage2@{}".format(k)] = 0.0
            if pop_mask is not None:
                tail_count = rec_count[:item_num][~pop_mask].sum()
                out["Tail%@{}".format(k)] = round(tail_count / n_rec, 4)

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
