"""``TopKEvaluator.evaluate(..., is_test=True)``: Pop/Niche, Cold/Warm group metrics and the Coverage / Gini / Gini2 /
Coverage2 / Tail% numbers against dicts the reference's own evaluator produced (tests/golden/evaluator_extras.npz,
written by tests/golden/make_golden.py::evaluator_fixtures from GenMMRec/src/utils/topk_evaluator.py:77-270).

The CPU test exercises the host logic (ground-truth filtering, key names, diversity arithmetic) with the oracle standing
in for the metrics kernel; the GPU test runs the same comparison through the real kernel."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import c_api


class _Dataset:
    def __init__(self, n):
        self.item_num = n


class _EvalData:
    """The slice of EvalDataLoader the evaluator touches (utils/dataloader.py: gt CSR, eval users, lengths)."""

    def __init__(self, z, device):
        self.gt_rowptr = torch.from_numpy(z["gt_rowptr"]).to(device)
        self.gt_items = torch.from_numpy(z["gt_items"]).to(device)
        self._users = torch.from_numpy(z["eval_users"])
        self._lens = np.diff(z["gt_rowptr"])
        self.dataset = _Dataset(int(z["n_items"]))

    def get_eval_len_list(self):
        return self._lens

    def get_eval_users(self):
        return self._users


def _config(z, case):
    cfg = {"metrics": [str(m) for m in z[case + "/metrics"]], "topk": [int(k) for k in z[case + "/topk_list"]],
           "save_recommended_topk": False}
    if case == "groups":
        cfg["pop_items"] = set(int(i) for i in z["pop_items"])
        cfg["warm_users"] = set(int(u) for u in z["warm_users"])
    return cfg


def _oracle_hits_metrics(topk, gt_rowptr, gt_items, return_hit=False):
    """CPU stand-in with the contract of ops.hits_metrics: per-position SUMS over users, optional hit matrix."""
    assert topk.dtype == torch.int32 and gt_rowptr.dtype == torch.int64 and gt_items.dtype == torch.int32
    assert gt_rowptr.numel() == topk.shape[0] + 1
    rp = gt_rowptr.numpy()
    for r in range(len(rp) - 1):  # the kernel binary-searches each row: rows must be ascending
        seg = gt_items.numpy()[rp[r]:rp[r + 1]]
        assert np.all(seg[1:] > seg[:-1])
    hit = c_api.hits(topk.numpy(), rp, gt_items.numpy())
    m = c_api.metrics(hit, np.diff(rp))
    sums = np.stack([m["recall"], m["ndcg"], m["precision"], m["map"]]) * float(topk.shape[0])
    return torch.from_numpy(sums), (torch.from_numpy(hit) if return_hit else None)


def _check(out, z, case, split):
    keys = [str(k) for k in z["%s/%s_keys" % (case, split)]]
    vals = z["%s/%s_values" % (case, split)]
    assert sorted(out.keys()) == sorted(keys)
    for k, v in zip(keys, vals):
        assert abs(float(out[k]) - float(v)) < 1e-12, (k, out[k], v)


@pytest.mark.parametrize("case", ["groups", "plain"])
def test_is_test_extras_match_reference_dicts_host_logic(case, monkeypatch):
    from genmmrec_b200 import ops
    from genmmrec_b200.utils.topk_evaluator import TopKEvaluator

    z, _ = load_golden("evaluator_extras")
    monkeypatch.setattr(ops, "hits_metrics", _oracle_hits_metrics)
    data = _EvalData(z, "cpu")
    topk = torch.from_numpy(z["topk"])
    ev = TopKEvaluator(_config(z, case))
    _check(ev.evaluate(topk, data, is_test=True), z, case, "test")
    _check(ev.evaluate([topk[:100], topk[100:]], data, is_test=False), z, case, "valid")
    if case == "groups":  # the fixture holds every group and the users the groups drop
        keys = set(str(k) for k in z["groups/test_keys"])
        assert {"Pop_Recall@5", "Niche_NDCG@50", "Cold_MAP@10", "Warm_Precision@20", "Pop_Recall2@5", "Tail%@50",
                "Gini2@5", "Coverage2@20"} <= keys


def test_recommendation_dump_has_the_reference_file_format(tmp_path, monkeypatch):
    """topk_evaluator.py:93-106: tab-separated, header ``id top_0 .. top_{K-1}``, one integer row per eval user, written
    only when ``save_recommended_topk`` is set and ``is_test`` is true."""
    import glob
    import os

    import pandas as pd

    from genmmrec_b200 import ops
    from genmmrec_b200.utils.topk_evaluator import TopKEvaluator

    z, _ = load_golden("evaluator_extras")
    monkeypatch.setattr(ops, "hits_metrics", _oracle_hits_metrics)
    cfg = dict(_config(z, "plain"), save_recommended_topk=True, recommend_topk=str(tmp_path / "rec"), model="DiffMM",
               dataset="toy")
    data = _EvalData(z, "cpu")
    ev = TopKEvaluator(cfg)
    ev.evaluate(torch.from_numpy(z["topk"]), data, is_test=False)
    assert not os.path.exists(cfg["recommend_topk"])
    ev.evaluate(torch.from_numpy(z["topk"]), data, is_test=True, idx=3)
    files = glob.glob(os.path.join(cfg["recommend_topk"], "DiffMM-toy-idx3-top50-*.csv"))
    assert len(files) == 1
    want = pd.DataFrame(z["topk"])  # the reference's own construction
    want.insert(0, "id", z["eval_users"])
    want.columns = ["id"] + ["top_" + str(i) for i in range(50)]
    want_path = tmp_path / "want.csv"
    want.astype(int).to_csv(want_path, sep="\t", index=False)
    assert open(files[0]).read() == open(want_path).read()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["groups", "plain"])
def test_is_test_extras_match_reference_dicts_on_device(case):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from genmmrec_b200.utils.topk_evaluator import TopKEvaluator

    z, _ = load_golden("evaluator_extras")
    data = _EvalData(z, "cuda:0")
    ev = TopKEvaluator(_config(z, case))
    _check(ev.evaluate(torch.from_numpy(z["topk"]).cuda(), data, is_test=True), z, case, "test")
