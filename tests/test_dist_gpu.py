"""Fused SpMM + all-gather (push) kernel and the sharded DiffMM dataflow on ONE GPU: the ranks are
emulated as separate destination buffers of one kernel launch (no kernels waiting on each other)."""
import numpy as np
import pytest
import torch

from conftest import golden_params, load_golden, toy_arrays
from oracle import c_api

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from genmmrec_b200 import dist as gd, ops
    return gd, ops


def test_spmm_push_writes_every_peer(mods):
    gd, ops = mods
    rng = np.random.default_rng(0)
    n_rows, n_cols, d = 700, 900, 128
    deg = rng.poisson(12, size=n_rows)
    deg[3] = 2000  # split row: goes through the partial-sum reduction, which pushes too
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    col = rng.integers(0, n_cols, size=int(rowptr[-1])).astype(np.int32)
    val = rng.standard_normal(col.size).astype(np.float32)
    x = rng.standard_normal((n_cols, d)).astype(np.float32)
    dev = torch.device("cuda:0")
    g = ops.GraphCSR(torch.from_numpy(rowptr).to(dev), torch.from_numpy(col).to(dev), torch.from_numpy(val).to(dev),
                     (n_rows, n_cols))
    total_rows, ld, off = 2000, 192, 1100
    peers = [torch.full((total_rows, ld), -7.0, device=dev) for _ in range(3)]
    table = torch.tensor([p.data_ptr() for p in peers], dtype=torch.int64, device=dev)
    gd.spmm_push(g, torch.from_numpy(x).to(dev), table, 3, off, ld, alpha=0.5)
    ref = 0.5 * c_api.spmm_csr_f64(rowptr, col, val, x)
    for p in peers:
        got = p[off:off + n_rows, :d].cpu().numpy()
        assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-5
        assert (p[:off] == -7).all() and (p[off + n_rows:] == -7).all() and (p[off:off + n_rows, d:] == -7).all()
    assert torch.equal(peers[0], peers[1]) and torch.equal(peers[0], peers[2])
    # same rows as the plain kernel, bit for bit
    y = ops.spmm_raw(g, torch.from_numpy(x).to(dev), alpha=0.5)
    assert torch.equal(y, peers[0][off:off + n_rows, :d])


def test_rows_push_copies_block_to_every_peer(mods):
    """Dense all-gather by peer stores: a strided source block lands at the right rows/columns of every replica and
    nothing else is touched."""
    gd, ops = mods
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda")
    g.manual_seed(2)
    wide = torch.randn(333, 192, device=dev, generator=g)
    src = wide[:, 64:128]                                    # 64 columns of a wider buffer (ld = 192)
    total_rows, ld, off = 1000, 128, 417
    peers = [torch.full((total_rows, ld), -3.0, device=dev) for _ in range(4)]
    table = torch.tensor([p.data_ptr() + 64 * 4 for p in peers], dtype=torch.int64, device=dev)  # column offset 64 folded in
    gd.rows_push(src, table, 4, off, ld)
    torch.cuda.synchronize()
    for p in peers:
        assert torch.equal(p[off:off + 333, 64:128], src)
        assert (p[:off] == -3).all() and (p[off + 333:] == -3).all() and (p[off:off + 333, :64] == -3).all()


def test_modal_mix_graph_equals_two_products(mods):
    """lambda (w0 A_v + w1 A_t) as one CSR: same product as the two accumulate passes it replaces, and its values follow
    the modality weights without rebuilding the structure."""
    gd, ops = mods
    from test_models_gpu import build, load_params, set_graphs
    from conftest import golden_params, load_golden, toy_arrays

    class Env:
        pass

    from genmmrec_b200 import synth
    from genmmrec_b200.common.trainer import Trainer
    from genmmrec_b200.utils.configurator import Config
    from genmmrec_b200.utils.dataloader import EvalDataLoader, TrainDataLoader
    from genmmrec_b200.utils.dataset import RecDataset
    from genmmrec_b200.utils.utils import get_model
    env = Env()
    env.synth, env.Trainer, env.Config, env.EvalDataLoader, env.TrainDataLoader, env.RecDataset, env.get_model = \
        synth, Trainer, Config, EvalDataLoader, TrainDataLoader, RecDataset, get_model
    z, meta = load_golden("toy_diffmm")
    data = toy_arrays()
    cfg, model, loaders = build(env, "DiffMM", meta, data, "toy")
    load_params(model, golden_params(z))
    set_graphs(env, "DiffMM", model, meta, data)
    e0 = model._packed_e0()
    for w0, w1 in ((0.5, 0.5), (0.2, 0.8)):
        mix = model._modal_mix_graph(model.image_UI_matrix, model.text_UI_matrix, w0, w1)
        got = ops.spmm_raw(mix, e0)
        lam = model.ris_adj_lambda
        ref = ops.spmm_raw(model.image_UI_matrix, e0, alpha=lam * w0)
        ops.spmm_raw(model.text_UI_matrix, e0, out=ref, alpha=lam * w1, beta=1.0)
        assert (got - ref).abs().max() <= 2e-6 * ref.abs().max()
    assert model._modal_mix_graph(model.image_UI_matrix, model.text_UI_matrix, 0.5, 0.5) is mix   # structure cached


def test_sharded_diffmm_world1_equals_model(mods):
    gd, ops = mods
    from test_models_gpu import build, load_params, set_graphs

    class Env:
        pass

    from genmmrec_b200 import synth
    from genmmrec_b200.common.trainer import Trainer
    from genmmrec_b200.utils.configurator import Config
    from genmmrec_b200.utils.dataloader import EvalDataLoader, TrainDataLoader
    from genmmrec_b200.utils.dataset import RecDataset
    from genmmrec_b200.utils.utils import get_model
    env = Env()
    env.synth, env.Trainer, env.Config, env.EvalDataLoader, env.TrainDataLoader, env.RecDataset, env.get_model = \
        synth, Trainer, Config, EvalDataLoader, TrainDataLoader, RecDataset, get_model
    z, meta = load_golden("toy_diffmm")
    data = toy_arrays()
    cfg, model, loaders = build(env, "DiffMM", meta, data, "toy")
    load_params(model, golden_params(z))
    set_graphs(env, "DiffMM", model, meta, data)
    sh = gd.ShardedDiffMM(model)
    with torch.no_grad():
        ue, ie = model.forward_MM(model.norm_adj, model.image_UI_matrix, model.text_UI_matrix)
        su, si = sh.forward_MM()
    # same kernels; the 64-wide and 192-wide SpMM variants add a row's nonzeros in different orders
    # (two interleaved partial sums vs one), so agreement is to rounding, not bitwise
    assert (su - ue).abs().max() <= 2e-6 * ue.abs().max() and (si - ie).abs().max() <= 2e-6 * ie.abs().max()
    assert np.abs(su.cpu().numpy() - z["emb/user"]).max() / np.abs(z["emb/user"]).max() < 1e-5
    part = gd.shard_eval_by_user_block(loaders["valid"], 100, 200)
    assert part.eval_u.numel() > 0 and int(part.eval_u.max()) < 100
    sh.close()


@pytest.mark.parametrize("name", ["GenRecV1", "LightGCN"])
def test_sharded_gcn_chain_world1_equals_model(mods, name):
    """The row-sharded LightGCN-style chain (GenRecV1 content embedding / LightGCN.forward) reproduces the model's own
    propagate() on one rank (every exchange is a store into the rank's own replica), for 1 and 3 layers."""
    gd, ops = mods
    from genmmrec_b200.workload import Workload
    for layers in (1, 3):
        wl = Workload(name, "toy", torch.device("cuda:0"), overrides={"n_layers": layers})
        model = wl.model
        sh = (gd.sharded_genrecv1 if name == "GenRecV1" else gd.sharded_lightgcn)(model)
        with torch.no_grad():
            ue, ie = model.propagate()
            su, items = sh.eval_factors()
        assert (su - ue).abs().max() <= 2e-6 * ue.abs().max()
        assert (items - ie).abs().max() <= 2e-6 * ie.abs().max()
        sh.close()
