"""Generate the golden fixtures under ``tests/golden/`` by running the UNMODIFIED reference
(``/root/reference``) on seeded synthetic data.  Run here (CPU container), commit the outputs:

    python tests/golden/make_golden.py            # toy fixtures for all six models + op fixtures
    python tests/golden/make_golden.py --baby     # additionally the Baby-shape samples
    python tests/golden/make_golden.py --skip-toy --sports     # GenRecV1 at the Sports shape (BASELINE config 3)
    python tests/golden/make_golden.py --skip-toy --clothing   # LD4MRec at the Clothing shape (BASELINE config 4)
    python tests/golden/make_golden.py --evaluator-only   # just the is_test evaluator dicts

The reference holds no golden vectors of its own (five smoke scripts asserting a
shape), so these files ARE the pin: the oracle (``oracle/``) is checked against them on CPU, and the
CUDA path is checked against the oracle and against them on the GPU box, where the reference tree
does not exist.
"""
import argparse
import json
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

import ref_harness as rh  # noqa: E402
from genmmrec_b200 import synth  # noqa: E402

TOY = dict(n_users=300, n_items=120, n_inter=3600, image_dim=64, text_dim=32)

MODEL_OVERRIDES = {
    "DiffMM": dict(n_layers=2, keep_rate=0.5, rebuild_k=1, ris_lambda=0.5, ris_adj_lambda=0.2),
    "GUME": dict(),
    "GenRecV1": dict(OpenInterestDebiase=False, num_layers=1, n_layers=2, rebuild_k=3, keep_rate=0.5),
    "LD4MRec": dict(svd_k=16, cnet_hidden_size=64),
    "VBPR": dict(),
    "LightGCN": dict(n_layers=3, reg_weight=1e-4, is_multimodal_model=False),
}


def loader_layout(loader):
    items = loader.get_eval_items()
    return {
        "eval_u": loader.eval_u.cpu().numpy().astype(np.int64),
        "pos_items_per_u": loader.pos_items_per_u.cpu().numpy().astype(np.int64),
        "train_pos_len_list": np.asarray(loader.train_pos_len_list, dtype=np.int64),
        "eval_len_list": np.asarray(loader.get_eval_len_list(), dtype=np.int64),
        "eval_items_flat": np.concatenate([np.asarray(x, dtype=np.int64) for x in items]),
    }


def set_generated_graphs(name, model, trainer, n_users, n_items, out):
    """DiffMM / GenRecV1 modality graphs: synthetic generated edges pushed through the reference's
    own buildUIMatrix + edgeDropper (common/trainer.py:471-485, models/diffmm.py:287-301)."""
    if name == "DiffMM":
        for attr, seed in (("image_UI_matrix", 11), ("text_UI_matrix", 12)):
            u, i = synth.generated_edges(n_users, n_items, model.rebuild_k, seed=seed)
            g = trainer.buildUIMatrix(u, i, np.ones(u.size))
            out["pre/%s" % attr] = rh.coo_parts(g)
            torch.manual_seed(seed)
            setattr(model, attr, model.edgeDropper(g))
    elif name == "GenRecV1":
        u, i = synth.generated_edges(n_users, n_items, model.rebuild_k, seed=11)
        g = trainer.buildUIMatrix(u, i, np.ones(u.size))
        out["pre/image_UI_matrix"] = rh.coo_parts(g)
        torch.manual_seed(11)
        model.image_UI_matrix = model.edgeDropper(g)
        trainer._build_item_item_matrix()


GRAPH_ATTRS = {
    "DiffMM": ["norm_adj", "image_UI_matrix", "text_UI_matrix"],
    "GUME": ["norm_adj", "R", "image_original_adj", "text_original_adj"],
    "GenRecV1": ["norm_adj", "R", "image_UI_matrix", "image_II_matrix", "text_II_matrix"],
    "LD4MRec": [],
    "VBPR": [],
    "LightGCN": ["norm_adj_matrix"],
}


def propagated(name, model):
    """The (user, item) tensors each model's full_sort_predict contracts."""
    with torch.no_grad():
        if name == "DiffMM":
            return model.forward_MM(model.norm_adj, model.image_UI_matrix, model.text_UI_matrix)
        if name == "GUME":
            e = model.forward(model.norm_adj)
            return e[:model.n_users], e[model.n_users:]
        if name == "GenRecV1":
            c, s = model.forward(model.R, model.norm_adj, model.image_UI_matrix, model.image_II_matrix,
                                 model.text_II_matrix)
            return c[:model.n_users], c[model.n_users:], s
        if name in ("VBPR", "LightGCN"):
            return model.forward()
    return None


def run_model(name, data_root, dataset, shape, out_path, sample=None, store_params=True):
    n_users, n_items = shape
    over = dict(MODEL_OVERRIDES[name])
    over["eval_batch_size"] = 128 if sample is None else 4096
    cfg = rh.ref_config(name, dataset, data_root, over)
    train, valid, test = rh.build_loaders(cfg)
    model, trainer = rh.build_model(name, cfg, train)
    model.eval()
    names = rh.override_params(model)
    if name == "LD4MRec":
        # the wide SpMM of models/ld4mrec.py:206 runs in __init__, before the override; it depends
        # only on the features and the graph, not on parameters.
        pass
    out = {}
    graphs = {}
    set_generated_graphs(name, model, trainer, n_users, n_items, graphs)
    for attr in GRAPH_ATTRS[name]:
        graphs[attr] = rh.coo_parts(getattr(model, attr))
    emb = propagated(name, model)
    res = {}
    for split, loader in (("valid", valid), ("test", test)):
        result, topk, hit, raw = rh.evaluate_capture(trainer, loader)
        res[split] = (result, topk, hit, raw)
    # scores of the first valid batch, straight from full_sort_predict
    with torch.no_grad():
        first = next(iter(valid))
        valid.pr = 0
        valid.inter_pr = 0
        scores0 = model.full_sort_predict(first).cpu().numpy()

    out["meta"] = np.frombuffer(json.dumps({
        "model": name, "n_users": n_users, "n_items": n_items, "overrides": over,
        "param_names": [[k, list(s)] for k, s in names],
        "config": {k: cfg[k] for k in ("embedding_size", "n_layers", "n_ui_layers", "knn_k", "keep_rate",
                                       "rebuild_k", "ris_lambda", "ris_adj_lambda", "trans_type", "topk",
                                       "metrics", "eval_batch_size", "svd_k", "cnet_hidden_size",
                                       "cnet_n_layers")},
    }).encode(), dtype=np.uint8)
    sd = model.state_dict()
    if store_params:
        for k, v in sd.items():
            if v.is_floating_point() and "denoise" not in k and "diffusion" not in k:
                out["param/" + k] = v.cpu().numpy()
    if name == "LD4MRec":
        out["buf/user_svd_emb"] = model.user_svd_emb.cpu().numpy()
        mm = model.user_mm_emb.cpu().numpy()
        if sample is None:
            out["buf/user_mm_emb"] = mm
        else:  # 32 sampled rows of the [n_users, 4480] wide-SpMM output + its scale
            out["buf/user_mm_emb/rows"] = sample["users"][:32].astype(np.int64)
            out["buf/user_mm_emb/values"] = mm[sample["users"][:32]]
            out["buf/user_mm_emb/maxabs"] = np.asarray([np.abs(mm[sample["users"]]).max()], dtype=np.float64)
    rng = np.random.default_rng(5)
    for attr, (idx, val) in graphs.items():
        if sample is None:
            out["graph/%s/indices" % attr] = idx
            out["graph/%s/values" % attr] = val
        else:
            pick = np.sort(rng.choice(val.size, size=min(4096, val.size), replace=False))
            out["graph/%s/nnz" % attr] = np.asarray([val.size], dtype=np.int64)
            out["graph/%s/sum" % attr] = np.asarray([val.astype(np.float64).sum()])
            out["graph/%s/pick" % attr] = pick
            out["graph/%s/indices" % attr] = idx[:, pick]
            out["graph/%s/values" % attr] = val[pick]
    if emb is not None:
        labels = ["user", "item", "side"]
        for lab, e in zip(labels, emb):
            e = e.detach().cpu().numpy()
            if sample is None:
                out["emb/" + lab] = e
            else:
                rows = sample["users"] if lab == "user" else sample["items"] if lab == "item" else sample["users"]
                out["emb/%s/rows" % lab] = rows
                out["emb/%s/values" % lab] = e[rows]
                out["emb/%s/maxabs" % lab] = np.asarray([np.abs(e).max()], dtype=np.float64)
                out["emb/%s/sum" % lab] = np.asarray([e.astype(np.float64).sum()])
    for split, (result, topk, hit, raw) in res.items():
        out["eval/%s/raw" % split] = raw.astype(np.float64)
        out["eval/%s/result_keys" % split] = np.asarray(list(result.keys()))
        out["eval/%s/result_vals" % split] = np.asarray([result[k] for k in result], dtype=np.float64)
        if sample is None:
            out["eval/%s/topk" % split] = topk.astype(np.int32)
            out["eval/%s/hit" % split] = hit
        else:
            pos = sample["eval_pos"][sample["eval_pos"] < topk.shape[0]]
            out["eval/%s/pos" % split] = pos
            out["eval/%s/topk" % split] = topk[pos].astype(np.int32)
            out["eval/%s/hit_colsum" % split] = hit.sum(axis=0).astype(np.int64)
    if sample is None:
        out["scores0"] = scores0
        for split, loader in (("valid", valid), ("test", test)):
            for k, v in loader_layout(loader).items():
                out["loader/%s/%s" % (split, k)] = v
    else:
        out["scores0/rows"] = np.arange(0, 8)
        out["scores0/values"] = scores0[:8]
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, "%.1f KB" % (os.path.getsize(out_path) / 1024),
          {k: res["valid"][0][k] for k in list(res["valid"][0])[:4]})


def op_fixtures(out_path):
    """Known-answer vectors for the stand-alone reference functions on the path."""
    rh.install_shims()
    from utils import metrics as ref_metrics
    from utils.utils import build_sim, build_knn_normalized_graph

    out = {}
    rng = np.random.default_rng(21)
    # utils/metrics.py:12-105 on a random hit matrix with pos_len both below and above K
    hit = rng.random((500, 50)) < 0.08
    pos_len = rng.integers(1, 80, size=500)
    out["metrics/hit"] = hit
    out["metrics/pos_len"] = pos_len.astype(np.int64)
    for m in ("recall", "ndcg", "precision", "map", "recall2"):
        out["metrics/" + m] = ref_metrics.metrics_dict[m](hit, pos_len).astype(np.float64)
    # utils/utils.py:147-197 kNN graph, both normalisers used on the path
    feat = np.maximum(rng.standard_normal((90, 24)).astype(np.float32), 0) + 0.01
    sim = build_sim(torch.from_numpy(feat))
    g = build_knn_normalized_graph(sim, topk=10, is_sparse=True, norm_type="sym")
    out["knn/feat"] = feat
    out["knn/sim"] = sim.numpy()
    out["knn/indices"], out["knn/values"] = rh.coo_parts(g)
    # torch.sparse.mm on an uncoalesced COO with duplicate entries and an empty row
    n_r, n_c, d = 37, 29, 64
    r = rng.integers(0, n_r, size=400)
    r[r == 5] = 6  # row 5 stays empty
    c = rng.integers(0, n_c, size=400)
    v = rng.standard_normal(400).astype(np.float32)
    x = rng.standard_normal((n_c, d)).astype(np.float32)
    a = torch.sparse_coo_tensor(torch.from_numpy(np.vstack([r, c])), torch.from_numpy(v), (n_r, n_c))
    out["spmm/rows"], out["spmm/cols"], out["spmm/vals"], out["spmm/x"] = r, c, v, x
    out["spmm/y"] = torch.sparse.mm(a, torch.from_numpy(x)).numpy()
    # mask + topk of common/trainer.py:384-386
    s = rng.standard_normal((16, 200)).astype(np.float32)
    mr = rng.integers(0, 16, size=300)
    mc = rng.integers(0, 200, size=300)
    st = torch.from_numpy(s.copy())
    st[torch.from_numpy(mr), torch.from_numpy(mc)] = -1e10
    val, idx = torch.topk(st, 50, dim=-1)
    out["topk/scores"], out["topk/mask_rows"], out["topk/mask_cols"] = s, mr, mc
    out["topk/values"], out["topk/indices"] = val.numpy(), idx.numpy().astype(np.int32)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, "%.1f KB" % (os.path.getsize(out_path) / 1024))


def evaluator_fixtures(out_path):
    """Known-answer dicts of the reference's own ``TopKEvaluator.evaluate(..., is_test=True)``
    (utils/topk_evaluator.py:77-270): base metrics, Pop/Niche and Cold/Warm group metrics, Coverage / Gini / Gini2 /
    Coverage2 / Tail%.  Two cases: with and without the popularity / warm-user groups quick_start.py:62-92 derives."""
    rh.install_shims()
    from utils.topk_evaluator import TopKEvaluator as RefEvaluator

    class _Dataset:
        def __init__(self, n):
            self.item_num = n

    class _EvalData:
        def __init__(self, users, items_per_u, n_items):
            self._u, self._items, self.dataset = users, items_per_u, _Dataset(n_items)

        def get_eval_items(self):
            return self._items

        def get_eval_len_list(self):
            return np.array([len(x) for x in self._items])

        def get_eval_users(self):
            return torch.from_numpy(self._u)

    rng = np.random.default_rng(31)
    n_users, n_items, k = 400, 150, 50
    eval_users = rng.permutation(1000)[:n_users].astype(np.int64)
    pop = np.sort(rng.choice(n_items, 30, replace=False)).astype(np.int64)
    items_per_u = []
    for u in range(n_users):
        n = int(rng.integers(1, 9))
        if u % 7 == 0:      # ground truth made of popular items only
            g = rng.choice(pop, min(n, 5), replace=False)
        elif u % 7 == 1:    # niche items only
            g = rng.choice(np.setdiff1d(np.arange(n_items), pop), n, replace=False)
        else:
            g = rng.choice(n_items, n, replace=False)
        items_per_u.append(np.sort(g).astype(np.int64))
    # recommendations: popularity-skewed permutations so that hits, coverage gaps and never-recommended items all occur
    w = 1.0 / (1.0 + np.arange(n_items)) ** 0.9
    w = w[rng.permutation(n_items)]
    topk = np.stack([rng.choice(n_items - 10, k, replace=False, p=w[:n_items - 10] / w[:n_items - 10].sum())
                     for _ in range(n_users)]).astype(np.int64)
    warm = np.sort(eval_users[rng.random(n_users) < 0.6])
    warm = np.concatenate([warm, np.array([5000, 5001])])  # ids outside the eval set are legal in the set
    out = {"topk": topk.astype(np.int32), "eval_users": eval_users, "n_items": np.int64(n_items),
           "gt_rowptr": np.concatenate([[0], np.cumsum([len(x) for x in items_per_u])]).astype(np.int64),
           "gt_items": np.concatenate(items_per_u).astype(np.int32), "pop_items": pop, "warm_users": warm}
    cases = {
        "groups": dict(metrics=["Recall", "NDCG", "Precision", "MAP", "Recall2"], topk=[5, 10, 20, 50],
                       pop_items=set(pop.tolist()), warm_users=set(warm.tolist())),
        "plain": dict(metrics=["Recall", "NDCG", "Precision", "MAP"], topk=[10, 50]),
    }
    for name, cfg in cases.items():
        cfg = dict(cfg, save_recommended_topk=False)
        ev = RefEvaluator(cfg)
        data = _EvalData(eval_users, items_per_u, n_items)
        res_test = ev.evaluate([torch.from_numpy(topk)], data, is_test=True)
        res_valid = RefEvaluator(cfg).evaluate([torch.from_numpy(topk)], data, is_test=False)
        out[name + "/metrics"] = np.array(cfg["metrics"])
        out[name + "/topk_list"] = np.array(cfg["topk"], dtype=np.int64)
        out[name + "/test_keys"] = np.array(list(res_test.keys()))
        out[name + "/test_values"] = np.array([float(v) for v in res_test.values()], dtype=np.float64)
        out[name + "/valid_keys"] = np.array(list(res_valid.keys()))
        out[name + "/valid_values"] = np.array([float(v) for v in res_valid.values()], dtype=np.float64)
        print(name, len(res_test), "keys, e.g.", {q: res_test[q] for q in list(res_test)[-6:]})
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, "%.1f KB" % (os.path.getsize(out_path) / 1024))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--baby", action="store_true")
    ap.add_argument("--sports", action="store_true", help="GenRecV1 at the Sports shape (sampled fixture)")
    ap.add_argument("--clothing", action="store_true", help="LD4MRec at the Clothing shape (sampled fixture)")
    ap.add_argument("--models", default="DiffMM,GUME,GenRecV1,LD4MRec,VBPR,LightGCN")
    ap.add_argument("--skip-toy", action="store_true")
    ap.add_argument("--evaluator-only", action="store_true", help="only (re)write evaluator_extras.npz")
    args = ap.parse_args()
    assert rh.reference_available(), "reference tree not found at %s" % rh.REF_ROOT
    torch.set_num_threads(os.cpu_count())
    if args.evaluator_only:
        evaluator_fixtures(os.path.join(HERE, "evaluator_extras.npz"))
        return
    models = args.models.split(",")
    tmp = tempfile.mkdtemp(prefix="gmr_golden_")
    try:
        if not args.skip_toy:
            op_fixtures(os.path.join(HERE, "ops.npz"))
            evaluator_fixtures(os.path.join(HERE, "evaluator_extras.npz"))
            synth.write_dataset(tmp, "toy", TOY["n_users"], TOY["n_items"], TOY["n_inter"],
                                image_dim=TOY["image_dim"], text_dim=TOY["text_dim"])
            for name in models:
                # GUME caches graphs in the dataset dir (models/gume.py:52-62,123-149): fresh dir per model
                d = os.path.join(tmp, "toy_" + name)
                shutil.copytree(os.path.join(tmp, "toy"), os.path.join(d, "toy"))
                run_model(name, d, "toy", (TOY["n_users"], TOY["n_items"]),
                          os.path.join(HERE, "toy_%s.npz" % name.lower()))
        if args.baby:
            nu, ni, nn, split = synth.SHAPES["baby"]
            synth.write_dataset(tmp, "baby", nu, ni, nn, split=split)
            rng = np.random.default_rng(77)
            sample = {
                "users": np.sort(rng.choice(nu, 256, replace=False)),
                "items": np.sort(rng.choice(ni, 256, replace=False)),
                "eval_pos": np.sort(np.concatenate([np.arange(128), rng.choice(np.arange(128, 19000), 384,
                                                                                 replace=False)])),
            }
            for name in models:
                if name not in ("DiffMM", "VBPR", "GUME"):
                    continue
                d = os.path.join(tmp, "baby_" + name)
                os.makedirs(os.path.join(d, "baby"))
                for f in os.listdir(os.path.join(tmp, "baby")):
                    os.symlink(os.path.join(tmp, "baby", f), os.path.join(d, "baby", f))
                run_model(name, d, "baby", (nu, ni), os.path.join(HERE, "baby_%s.npz" % name.lower()),
                          sample=sample, store_params=False)
        # BASELINE configs 3 and 4: one model each at its own shape, sampled like the Baby fixtures
        for flag, shape_name, name in ((args.sports, "sports", "GenRecV1"), (args.clothing, "clothing", "LD4MRec")):
            if not flag:
                continue
            nu, ni, nn, split = synth.SHAPES[shape_name]
            synth.write_dataset(tmp, shape_name, nu, ni, nn, split=split)
            rng = np.random.default_rng(78)
            sample = {
                "users": np.sort(rng.choice(nu, 256, replace=False)),
                "items": np.sort(rng.choice(ni, 256, replace=False)),
                "eval_pos": np.sort(np.concatenate([np.arange(128), rng.choice(np.arange(128, int(nu * 0.95)), 384,
                                                                                 replace=False)])),
            }
            run_model(name, tmp, shape_name, (nu, ni), os.path.join(HERE, "%s_%s.npz" % (shape_name, name.lower())),
                      sample=sample, store_params=False)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
