"""The vectorised graph builders of the package (host logic) against the reference's own graphs
(tests/golden/toy_*.npz): entry order and values bit for bit.  CPU only."""
import numpy as np
import torch

from conftest import golden_params, load_golden
from genmmrec_b200 import graph, synth


def train_pairs(d):
    m = d["label"] == 0
    return d["users"][m], d["items"][m]


def same(parts, z, attr, exact=True):
    idx, val = parts[0].numpy(), parts[1].numpy()
    assert np.array_equal(idx, z["graph/%s/indices" % attr]), attr
    if exact:
        assert np.array_equal(val, z["graph/%s/values" % attr]), attr
    else:
        assert np.abs(val - z["graph/%s/values" % attr]).max() <= 1e-7 * np.abs(val).max(), attr


def test_norm_adj(toy_data):
    z, _ = load_golden("toy_diffmm")
    tu, ti = train_pairs(toy_data)
    parts = graph.norm_adj(tu, ti, 300, 120)
    same(parts, z, "norm_adj")
    z2, _ = load_golden("toy_lightgcn")
    same(parts, z2, "norm_adj_matrix")
    # duplicates collapse (dict keys in the reference)
    parts2 = graph.norm_adj(np.concatenate([tu, tu[:50]]), np.concatenate([ti, ti[:50]]), 300, 120)
    assert torch.equal(parts2[0], parts[0]) and torch.equal(parts2[1], parts[1])


def test_bipartite_halves(toy_data):
    tu, ti = train_pairs(toy_data)
    idx, val, _ = graph.norm_adj(tu, ti, 300, 120)
    ui, iu = graph.bipartite_halves(idx, val, 300, 120)
    assert ui[0].shape[1] == iu[0].shape[1] == idx.shape[1] // 2
    dense = torch.sparse_coo_tensor(idx, val, (420, 420)).to_dense()
    assert torch.equal(torch.sparse_coo_tensor(*ui).to_dense(), dense[:300, 300:])
    assert torch.equal(torch.sparse_coo_tensor(*iu).to_dense(), dense[300:, :300])
    # a graph with self loops is not bipartite
    u, i = synth.generated_edges(300, 120, 1, seed=11)
    idx2, val2, _ = graph.ui_matrix(u, i, 300, 120)
    assert graph.bipartite_halves(idx2, val2, 300, 120) is None


def test_ui_matrix_and_edge_drop(toy_data):
    z, meta = load_golden("toy_diffmm")
    cfg = meta["config"]
    for attr, seed in (("image_UI_matrix", 11), ("text_UI_matrix", 12)):
        u, i = synth.generated_edges(300, 120, cfg["rebuild_k"], seed=seed)
        parts = graph.ui_matrix(u, i, 300, 120)
        same(parts, z, "pre/" + attr)
        torch.manual_seed(seed)
        idx, val = graph.drop_edges(parts[0], parts[1], cfg["keep_rate"])
        same((idx, val), z, attr)
    z, meta = load_golden("toy_genrecv1")
    u, i = synth.generated_edges(300, 120, meta["config"]["rebuild_k"], seed=11)
    parts = graph.ui_matrix(u, i, 300, 120)
    same(parts, z, "pre/image_UI_matrix")
    torch.manual_seed(11)
    same(graph.drop_edges(parts[0], parts[1], meta["config"]["keep_rate"]), z, "image_UI_matrix")


def test_knn_graphs(toy_data):
    z, meta = load_golden("toy_gume")
    p = golden_params(z)
    k = meta["config"]["knn_k"]
    img = graph.knn_graph_dense(torch.from_numpy(p["image_embedding.weight"]), k)
    txt = graph.knn_graph_dense(torch.from_numpy(p["text_embedding.weight"]), k)
    same(img, z, "image_original_adj", exact=False)
    same(txt, z, "text_original_adj", exact=False)
    z2, meta2 = load_golden("toy_genrecv1")
    same(graph.knn_graph_dense(torch.from_numpy(toy_data["img"]), k, eps_normalize=True), z2, "image_II_matrix", exact=False)
    same(graph.knn_graph_dense(torch.from_numpy(toy_data["txt"]), k, eps_normalize=True), z2, "text_II_matrix", exact=False)
    zo, _ = load_golden("ops")
    g = graph.knn_graph_dense(torch.from_numpy(zo["knn/feat"]), 10)
    assert np.array_equal(g[0].numpy(), zo["knn/indices"])
    assert np.abs(g[1].numpy() - zo["knn/values"]).max() < 1e-7


def test_gume_adj(toy_data):
    z, meta = load_golden("toy_gume")
    tu, ti = train_pairs(toy_data)
    k = meta["config"]["knn_k"]
    img_idx = torch.from_numpy(z["graph/image_original_adj/indices"][1].reshape(-1, k))
    txt_idx = torch.from_numpy(z["graph/text_original_adj/indices"][1].reshape(-1, k))
    na, r = graph.gume_adj(tu, ti, 300, 120, img_idx, txt_idx)
    same(na, z, "norm_adj", exact=False)
    same(r, z, "R", exact=False)


def test_ld4mrec_rnorm_and_r(toy_data):
    tu, ti = train_pairs(toy_data)
    from oracle import ref_port as rp
    idx, val, shape = graph.ld4mrec_rnorm(tu, ti, 300, 120)
    o_idx, o_val, _ = rp.ld4mrec_rnorm(tu, ti, 300, 120)
    a = torch.sparse_coo_tensor(idx, val, shape).to_dense().numpy()
    b = torch.sparse_coo_tensor(torch.from_numpy(o_idx), torch.from_numpy(o_val), shape).to_dense().numpy()
    assert np.abs(a - b).max() <= 1e-7
    z, _ = load_golden("toy_genrecv1")
    same(graph.binary_r(tu, ti, 300, 120), z, "R")
