"""Training-side fusion on cuda:0, through the C ABI: the fused BPR gather + dot kernels
against plain PyTorch autograd (fp32 reference of the same op), the gradients of a whole ``calculate_loss`` against the
CPU restatement of the reference under autograd, the multi-RHS ``forward_cl_MM``, the device edge extraction that
replaces the reference trainer's per-element loop, and ``Trainer.fit``."""
import numpy as np
import pytest
import torch

from conftest import golden_params, load_golden, toy_arrays
from oracle import ref_port as rp
from test_models_gpu import build, env, load_params, set_graphs  # noqa: F401  (env is a fixture)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d", [64, 128, 20, 7])
def test_bpr_scores_match_torch_autograd(d):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from genmmrec_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(d)
    nu, ni, b = 500, 300, 4000   # b > nu: every user and most items repeat inside the batch (atomic accumulation)
    eu = torch.randn(nu, d, device="cuda", generator=g, requires_grad=True)
    ei = torch.randn(ni, d, device="cuda", generator=g, requires_grad=True)
    users = torch.randint(0, nu, (b,), device="cuda", generator=g)
    pos = torch.randint(0, ni, (b,), device="cuda", generator=g)
    neg = torch.randint(0, ni, (b,), device="cuda", generator=g)
    w = torch.randn(b, device="cuda", generator=g)
    ps, ns = ops.bpr_scores(eu, ei, users, pos, neg)
    loss = (w * torch.nn.functional.logsigmoid(ps - ns)).sum() + 0.1 * (ps * ns).sum()
    loss.backward()
    g_u, g_i = eu.grad.clone(), ei.grad.clone()
    eu.grad = ei.grad = None
    rps, rns = (eu[users] * ei[pos]).sum(-1), (eu[users] * ei[neg]).sum(-1)
    rloss = (w * torch.nn.functional.logsigmoid(rps - rns)).sum() + 0.1 * (rps * rns).sum()
    rloss.backward()
    assert float((ps - rps).abs().max()) <= 1e-5 * float(rps.abs().max())
    assert float((ns - rns).abs().max()) <= 1e-5 * float(rns.abs().max())
    assert float((g_u - eu.grad).abs().max()) <= 2e-5 * float(eu.grad.abs().max())
    assert float((g_i - ei.grad).abs().max()) <= 2e-5 * float(ei.grad.abs().max())
    # column slices of wider buffers (the propagated tables are slices) go through the same kernels
    wide = torch.randn(nu, 2 * d + 3, device="cuda", generator=g)
    ps2, _ = ops.bpr_scores(wide[:, 1:d + 1], ei.detach(), users, pos, neg)
    assert float((ps2 - (wide[:, 1:d + 1][users] * ei.detach()[pos]).sum(-1)).abs().max()) <= 1e-4


def test_diffmm_loss_gradients_match_reference_port(env):  # noqa: F811
    """calculate_loss (BPR + L2; the contrastive weight is set to 0 so that the CPU restatement covers every term) through
    the differentiable SpMM (backward = K1 on the transposed graph) and the fused BPR kernels, against the restated
    reference under plain torch autograd on the CPU."""
    z, meta = load_golden("toy_diffmm")
    data = toy_arrays()
    cfg, model, loaders = build(env, "DiffMM", meta, data, "toy")
    load_params(model, golden_params(z))
    set_graphs(env, "DiffMM", model, meta, data)
    model.ssl_reg = 0.0
    model.train()
    rng = np.random.default_rng(0)
    b = 512
    inter = torch.from_numpy(np.stack([rng.integers(0, data["n_users"], b), rng.integers(0, data["n_items"], b),
                                       rng.integers(0, data["n_items"], b)])).to(model.device)
    loss = model.calculate_loss(inter)
    loss.backward()
    names = ["uEmbeds", "iEmbeds", "image_trans", "text_trans", "modal_weight"]
    got = {n: getattr(model, n).grad.detach().cpu() for n in names}

    p = {k: v.detach().cpu().clone().requires_grad_(k in names) for k, v in model.state_dict().items()}

    def coo(gr):
        t = gr.to_torch_coo()
        return torch.sparse_coo_tensor(t._indices().cpu(), t._values().cpu(), t.shape)

    c = meta["config"]
    ue, ie = rp.diffmm_forward_mm(p, coo(model.norm_adj.full), coo(model.image_UI_matrix), coo(model.text_UI_matrix),
                                  model.v_feat.cpu(), model.t_feat.cpu(), data["n_users"], c["n_layers"], c["ris_lambda"],
                                  c["ris_adj_lambda"])
    u, pi, ni = inter.cpu()
    bpr = -torch.log(1e-10 + torch.sigmoid((ue[u] * ie[pi]).sum(1) - (ue[u] * ie[ni]).sum(1))).mean()
    ref = bpr + (p["uEmbeds"].norm(2).square() + p["iEmbeds"].norm(2).square()) * model.reg_weight   # diffmm.py:203-221
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    for n in names:
        scale = float(p[n].grad.abs().max())
        assert float((got[n] - p[n].grad).abs().max()) <= 2e-4 * scale + 1e-9, n


def test_genrecv1_loss_matches_reference_port(env):  # noqa: F811
    """GenRecV1.calculate_loss (BPR + L2 + four in-batch InfoNCE terms over the content / side embeddings,
    genrecv1.py:355-415) against the CPU restatement under autograd: loss value and embedding gradients."""
    z, meta = load_golden("toy_genrecv1")
    data = toy_arrays()
    cfg, model, loaders = build(env, "GenRecV1", meta, data, "toy")
    load_params(model, golden_params(z))
    set_graphs(env, "GenRecV1", model, meta, data)
    model.eval()   # BatchNorm / Dropout in inference mode on both sides (the port restates eval-mode layers)
    rng = np.random.default_rng(1)
    b = 256
    inter = torch.from_numpy(np.stack([rng.integers(0, data["n_users"], b), rng.integers(0, data["n_items"], b),
                                       rng.integers(0, data["n_items"], b)])).to(model.device)
    loss = model.calculate_loss(inter)
    loss.backward()
    names = ["user_embedding.weight", "item_id_embedding.weight"]
    got = {"user_embedding.weight": model.user_embedding.weight.grad.detach().cpu(),
           "item_id_embedding.weight": model.item_id_embedding.weight.grad.detach().cpu()}
    p = {k: v.detach().cpu().clone().requires_grad_(k in names) for k, v in model.state_dict().items()}

    def coo(gr):
        t = gr.to_torch_coo()
        return torch.sparse_coo_tensor(t._indices().cpu(), t._values().cpu(), t.shape)

    from genmmrec_b200.models._common import as_graph
    c = meta["config"]
    nu = data["n_users"]
    ue, ie = rp.genrecv1_content(p, coo(as_graph(model.norm_adj)), coo(model.image_UI_matrix), nu, c["n_layers"])
    side = rp.genrecv1_side(p, torch.cat([ue, ie]), coo(model.R), coo(model.image_II_matrix), coo(model.text_II_matrix),
                            model.image_embedding.cpu(), model.text_embedding.cpu(), c["n_layers"])
    su, si = side[:nu], side[nu:]
    u, pi, ni = inter.cpu()
    F = torch.nn.functional

    def nce(a, bb, t):
        a, bb = F.normalize(a, dim=1), F.normalize(bb, dim=1)
        return -torch.log(torch.exp((a * bb).sum(-1) / t) / torch.exp(a @ bb.t() / t).sum(1)).mean()

    ref = -torch.mean(F.logsigmoid((ue[u] * ie[pi]).sum(-1) - (ue[u] * ie[ni]).sum(-1))) \
        + (p[names[0]].norm(2).square() + p[names[1]].norm(2).square()) * model.reg_weight \
        + (nce(si[pi], ie[pi], model.temp) + nce(su[u], ue[u], model.temp)) * model.ssl_reg1 \
        + (nce(ue[u], ie[pi], model.temp) + nce(ue[u], si[pi], model.temp)) * model.ssl_reg2
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 2e-5 * abs(float(ref))
    for n in names:
        scale = float(p[n].grad.abs().max())
        assert float((got[n] - p[n].grad).abs().max()) <= 5e-4 * scale + 1e-9, n


def test_ld4mrec_diffusion_loss_matches_reference_port(env):  # noqa: F811
    """LD4MRec.calculate_loss (ld4mrec.py:265-344) with the random draws fixed: device-built history rows, q-sample,
    C-Net x0 prediction, label-smoothed MSE, and the closed-form loss-history update, against the CPU restatement
    (which updates the history one sample at a time like the reference)."""
    z, meta = load_golden("toy_ld4mrec")
    data = toy_arrays()
    cfg, model, loaders = build(env, "LD4MRec", meta, data, "toy")
    load_params(model, golden_params(z))
    model.user_svd_emb = torch.from_numpy(z["buf/user_svd_emb"]).to(model.device)
    model.eval()                       # dropout off on both sides
    nu, ni = data["n_users"], data["n_items"]
    assert float((model.alpha_bar.cpu() - rp.ld4mrec_noise_schedule(model.steps, 0.001)).abs().max()) <= 1e-6
    rng = np.random.default_rng(3)
    b = 96
    user = torch.from_numpy(rng.integers(0, nu, b)).to(model.device)
    t = torch.from_numpy(rng.integers(0, 7, b)).to(model.device)     # 7 distinct steps: every step repeats in the batch
    noise = torch.from_numpy(rng.standard_normal((b, ni)).astype(np.float32)).to(model.device)
    # history rows == the dense slice of the train matrix the reference takes on the host
    tr = data["label"] == 0
    dense = np.zeros((nu, ni), dtype=np.float32)
    dense[data["users"][tr], data["items"][tr]] = 1.0
    x_in = model.history_rows(user)
    assert np.array_equal(x_in.cpu().numpy(), dense[user.cpu().numpy()])
    h0 = model.loss_history.clone()
    h0 += torch.linspace(0, 1, model.steps, device=h0.device)          # a non-trivial starting history
    model.loss_history.copy_(h0)
    loss = model.calculate_loss((user,), t=t, noise=noise)
    loss.backward()
    p = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    rows = rp.ld4mrec_loss(p, dense[user.cpu().numpy()], t.cpu(), noise.cpu(), model.alpha_bar.cpu(),
                           model.user_svd_emb[user].cpu(), model.user_mm_emb[user].cpu(), model.cnet_layers,
                           model.smoothing_gamma)
    assert abs(float(loss) - float(rows.mean())) <= 2e-5 * abs(float(rows.mean()))
    want_h = rp.ld4mrec_loss_history(h0.cpu().numpy(), t.cpu().numpy(), rows.numpy())
    assert float(np.abs(model.loss_history.cpu().numpy() - want_h).max()) <= 1e-5 * float(np.abs(want_h).max())
    assert model.cnet.output_proj.weight.grad is not None and bool(torch.isfinite(model.cnet.output_proj.weight.grad).all())
    # without fixed draws: steps come from the device-resident history (torch.multinomial), the loss is finite
    assert bool(torch.isfinite(model.calculate_loss((user,))))


def test_forward_cl_mm_multi_rhs_equals_per_view_chains(env):  # noqa: F811
    z, meta = load_golden("toy_diffmm")
    data = toy_arrays()
    cfg, model, loaders = build(env, "DiffMM", meta, data, "toy")
    load_params(model, golden_params(z))
    set_graphs(env, "DiffMM", model, meta, data)
    from genmmrec_b200.models._common import as_graph
    from genmmrec_b200.ops import spmm
    import torch.nn.functional as F
    adj, ia, ta = as_graph(model.norm_adj), as_graph(model.image_UI_matrix), as_graph(model.text_UI_matrix)
    with torch.no_grad():
        u1, i1, u2, i2 = model.forward_cl_MM(model.norm_adj, model.image_UI_matrix, model.text_UI_matrix)

        def view(m_adj, feats):   # diffmm.py:171-195, one view at a time
            e = spmm(m_adj, torch.concat([model.uEmbeds, F.normalize(feats)]))
            lst = [e]
            for _ in range(model.gnn_layer):
                lst.append(spmm(adj, lst[-1]))
            return sum(lst)

        e1, e2 = view(ia, model.getImageFeats()), view(ta, model.getTextFeats())
    nu = data["n_users"]
    for got, want in ((u1, e1[:nu]), (i1, e1[nu:]), (u2, e2[:nu]), (i2, e2[nu:])):
        assert float((got - want).abs().max()) <= 2e-6 * float(want.abs().max())


def test_edges_from_denoised_matches_the_reference_loop(env):  # noqa: F811
    from genmmrec_b200.models.diffmm import DiffMM
    g = torch.Generator(device="cuda").manual_seed(3)
    batch_index = torch.randperm(500, device="cuda", generator=g)[:64]
    den = torch.randn(64, 120, device="cuda", generator=g)
    k = 3
    u, i = DiffMM.edges_from_denoised(batch_index, den, k)
    _, idx = torch.topk(den, k=k)
    u_list, i_list = [], []
    for a in range(batch_index.shape[0]):           # common/trainer.py:548-553, literally
        for b in range(idx[a].shape[0]):
            u_list.append(int(batch_index[a].cpu().numpy()))
            i_list.append(int(idx[a][b].cpu().numpy()))
    assert u.cpu().tolist() == u_list and i.cpu().tolist() == i_list


@pytest.mark.parametrize("name", ["LightGCN", "VBPR"])
def test_trainer_fit_trains_and_early_stops(env, name):  # noqa: F811
    z, meta = load_golden("toy_" + name.lower())
    data = toy_arrays()
    cfg, model, loaders = build(env, name, meta, data, "toy",
                                extra={"epochs": 6, "stopping_step": 2, "learning_rate": 0.01, "train_batch_size": 512,
                                       "eval_step": 1, "learning_rate_scheduler": [0.96, 50]})
    from genmmrec_b200.utils.dataloader import TrainDataLoader
    load_params(model, golden_params(z))
    trainer = env.Trainer(cfg, model)
    ds = env.RecDataset.from_arrays(cfg, data["users"], data["items"], data["label"], data["n_users"], data["n_items"])
    tr, _, _ = ds.split()
    train = TrainDataLoader(cfg, tr, batch_size=512, shuffle=True)
    before = trainer.evaluate(loaders["valid"])
    best, best_valid, best_test = trainer.fit(train, loaders["valid"], loaders["test"])
    assert len(trainer.train_loss_dict) >= 1
    losses = [trainer.train_loss_dict[e] for e in sorted(trainer.train_loss_dict)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0]          # the objective goes down
    assert best >= before[trainer.valid_metric] - 1e-12                  # best validation score is tracked
    assert set(best_valid) == set(before)
