"""Synthetic workloads of the published shapes, assembled into ready-to-run models + loaders.

Small shapes (toy / baby / sports / clothing) come from the bit-reproducible numpy generator
(``synth.py``) so that tests, golden fixtures and bench lines see the same data; the 1M-user shape is
generated directly on the GPU with the same recipe (heavy-tailed user degrees, rank^-0.8 item
popularity, unique pairs, per-user leave-two-out split) because a host-side build would dominate
the run.
"""
import torch

from . import synth
from .utils.configurator import Config
from .utils.dataloader import EvalDataLoader, TrainDataLoader
from .utils.dataset import RecDataset
from .utils.utils import get_model


def make_interactions_torch(n_users, n_items, n_inter, device, seed=999, split="loo", alpha=0.8):
    """GPU twin of ``synth.make_interactions`` (same recipe, torch RNG).  Returns int64 tensors
    (users, items, label) grouped by user in random per-user time order."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    extra = 2 if split == "loo" else 0
    min_deg = 5 + extra
    total = n_inter + extra * n_users
    w = torch.exp(torch.randn(n_users, device=device, generator=g, dtype=torch.float64))
    deg = min_deg + torch.floor(w / w.sum() * (total - min_deg * n_users)).long()
    deg.clamp_(max=max(8, n_items // 4))
    users = torch.repeat_interleave(torch.arange(n_users, device=device), deg)
    p = torch.arange(1, n_items + 1, device=device, dtype=torch.float64) ** (-alpha)
    cdf = torch.cumsum(p / p.sum(), 0)
    perm = torch.randperm(n_items, device=device, generator=g)

    def draw(n):
        r = torch.rand(n, device=device, generator=g, dtype=torch.float64)
        return perm[torch.searchsorted(cdf, r).clamp_(max=n_items - 1)]

    items = draw(users.numel())
    for it in range(6):  # redraw duplicated (user, item) pairs
        key = users * n_items + items
        order = torch.argsort(key)
        sk = key[order]
        dup = torch.zeros_like(key, dtype=torch.bool)
        dup[order[1:]] = sk[1:] == sk[:-1]
        n_dup = int(dup.sum())
        if n_dup == 0:
            break
        items[dup] = draw(n_dup) if it < 4 else torch.randint(0, n_items, (n_dup,), device=device, generator=g)
    key = torch.unique(users * n_items + items)  # whatever is still duplicated is dropped
    users, items = torch.div(key, n_items, rounding_mode="floor"), key % n_items
    # random time order inside each user
    order = torch.argsort(users.double() + torch.rand(users.numel(), device=device, generator=g, dtype=torch.float64) * 0.999)
    users, items = users[order], items[order]
    cnt = torch.bincount(users, minlength=n_users)
    start = torch.cumsum(cnt, 0) - cnt
    pos = torch.arange(users.numel(), device=device) - start[users]
    n_u = cnt[users]
    if split == "loo":
        n_eval = torch.ones_like(n_u)
    else:
        n_eval = torch.where(n_u < 10, torch.ones_like(n_u), torch.clamp(n_u // 10, min=1))
    label = torch.zeros_like(users)
    label[pos >= n_u - 2 * n_eval] = 1
    label[pos >= n_u - n_eval] = 2
    label[n_u < 3] = 0  # degenerate users (lost draws to de-duplication) stay train-only
    return users, items, label


def make_features_torch(n_items, device, seed=999, image_dim=4096, text_dim=384):
    g = torch.Generator(device=device)
    g.manual_seed(seed + 1)
    img = torch.randn((n_items, image_dim), device=device, generator=g).clamp_(min=0)
    txt = torch.randn((n_items, text_dim), device=device, generator=g)
    txt = txt / txt.norm(dim=1, keepdim=True)
    return img, txt


def generated_edges_torch(n_users, n_items, rebuild_k, device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    u = torch.arange(n_users, device=device).repeat_interleave(rebuild_k)
    i = torch.randint(0, n_items, (n_users * rebuild_k,), device=device, generator=g)
    return u, i


class Workload(object):
    """A model of the named family at a named synthetic shape, with its eval loaders."""

    def __init__(self, model_name, shape, device, overrides=None, seed=999, image_dim=None, text_dim=None,
                 on_gpu=None):
        nu, ni, nn, split = synth.SHAPES[shape]
        self.shape, self.n_users, self.n_items, self.n_inter = shape, nu, ni, nn
        device = torch.device(device)
        big = (nn >= 5_000_000) if on_gpu is None else on_gpu
        image_dim = image_dim or synth.FEAT_DIMS["image"]
        text_dim = text_dim or synth.FEAT_DIMS["text"]
        if big:
            users, items, label = make_interactions_torch(nu, ni, nn, device, seed, split)
            img, txt = make_features_torch(ni, device, seed, image_dim, text_dim)
        else:
            u, i, l = synth.make_interactions(nu, ni, nn, seed, split)
            users, items, label = (torch.from_numpy(a).to(device) for a in (u, i, l))
            im, tx = synth.make_features(ni, seed, image_dim, text_dim)
            img, txt = torch.from_numpy(im).to(device), torch.from_numpy(tx).to(device)
        cfg_over = {"device": device, "preloaded_features": (img, txt), "skip_svd": True}
        cfg_over.update(overrides or {})
        self.config = Config(model_name, shape, cfg_over)
        ds = RecDataset.from_arrays(self.config, users, items, label, nu, ni)
        self.train_ds, self.valid_ds, self.test_ds = ds.split()
        self.train = TrainDataLoader(self.config, self.train_ds, batch_size=self.config["train_batch_size"])
        self.valid = EvalDataLoader(self.config, self.valid_ds, additional_dataset=self.train_ds,
                                    batch_size=self.config["eval_batch_size"])
        torch.manual_seed(seed)
        self.model = get_model(model_name)(self.config, self.train).to(device)
        self.model.eval()
        if model_name == "DiffMM":
            k = self.config["rebuild_k"]
            mk = generated_edges_torch if big else None
            for attr, s in (("image", 11), ("text", 12)):
                if big:
                    e = mk(nu, ni, k, device, seed + s)
                else:
                    e = synth.generated_edges(nu, ni, k, seed=s)
                setattr(self, attr + "_edges", e)
            torch.manual_seed(seed)
            self.model.set_generated_edges(self.image_edges, self.text_edges)
        elif model_name == "GenRecV1":
            e = generated_edges_torch(nu, ni, self.config["rebuild_k"], device, seed + 11) if big else \
                synth.generated_edges(nu, ni, self.config["rebuild_k"], seed=11)
            torch.manual_seed(seed)
            self.model.set_generated_edges(e)

    @property
    def n_eval_users(self):
        return int(self.valid.eval_u.numel())

    @property
    def nnz_train(self):
        return len(self.train_ds)
