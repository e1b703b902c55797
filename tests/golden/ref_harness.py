"""Bring the UNMODIFIED reference (``/root/reference``) up inside this container so that golden
vectors can be generated from its own classes.  Test infrastructure only: nothing in the product
package, the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports this module, because the
reference tree does not exist on the GPU box.

The shims below do not touch hot-path arithmetic:
  * stub ``matplotlib`` / ``lmdb`` modules (imported, never used on this path);
  * ``numpy.float`` alias (removed in numpy >= 1.24, used by utils/metrics.py:51,57,81,84);
  * a ``torch_scatter.scatter_add`` stand-in (``index_add_``) for utils/utils.py:153-155;
  * ``str(dataset)`` on every split because ``inter_num`` is only set in ``RecDataset.__str__``
    (utils/dataset.py:123, read at utils/dataloader.py:55).
"""
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("GMR_REFERENCE", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "GenMMRec", "src")


def reference_available():
    return os.path.isdir(REF_SRC)


def install_shims():
    for m in ("matplotlib", "matplotlib.pyplot", "lmdb"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(np, "float"):
        np.float = float
    if "torch_scatter" not in sys.modules:
        ts = types.ModuleType("torch_scatter")

        def scatter_add(src, index, dim=0, dim_size=None):
            return torch.zeros(dim_size, dtype=src.dtype, device=src.device).index_add_(0, index, src)

        ts.scatter_add = scatter_add
        sys.modules["torch_scatter"] = ts
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)


def ref_config(model, dataset, data_root, overrides=None):
    """The reference's own Config (YAML merge order of utils/configurator.py:68-90)."""
    install_shims()
    from utils.configurator import Config

    cfg_dict = {
        # keys of configs/dataset/baby.yaml for a dataset that has no yaml of its own
        "USER_ID_FIELD": "userID", "ITEM_ID_FIELD": "itemID", "RATING_FIELD": "rating",
        "TIME_FIELD": "timestamp", "filter_out_cod_start_users": True,
        "inter_file_name": dataset + ".inter", "vision_feature_file": "image_feat.npy",
        "text_feature_file": "text_feat.npy", "field_separator": "\t",
        "data_path": os.path.join(data_root, ""), "use_gpu": False, "gpu_id": 0,
        "use_wandb": False, "save_recommended_topk": False, "use_neighborhood_loss": False,
        "seed": 999,
    }
    cfg_dict.update(overrides or {})
    cwd = os.getcwd()
    os.chdir(REF_SRC)  # configurator.py:72-73 resolves ``configs/`` against the cwd
    try:
        cfg = Config(model, dataset, cfg_dict)
    finally:
        os.chdir(cwd)
    cfg["device"] = torch.device("cpu")
    return cfg


def build_loaders(cfg):
    install_shims()
    from utils.dataset import RecDataset
    from utils.dataloader import TrainDataLoader, EvalDataLoader

    ds = RecDataset(cfg)
    str(ds)
    tr, va, te = ds.split()
    for x in (tr, va, te):
        str(x)
    train = TrainDataLoader(cfg, tr, batch_size=cfg["train_batch_size"], shuffle=True)
    valid = EvalDataLoader(cfg, va, additional_dataset=tr, batch_size=cfg["eval_batch_size"])
    test = EvalDataLoader(cfg, te, additional_dataset=tr, batch_size=cfg["eval_batch_size"])
    return train, valid, test


def build_model(name, cfg, train):
    install_shims()
    from utils.utils import init_seed, get_model, get_trainer

    if name == "LightGCN":
        _patch_lightgcn()
    init_seed(999)
    train.pretrain_setup()
    model = get_model(name)(cfg, train).to(cfg["device"])
    trainer = get_trainer(name)(cfg, model, False)
    return model, trainer


def _patch_lightgcn():
    """models/lightgcn.py:86 calls the private ``dok_matrix._update`` that modern scipy removed.
    DiffMM/GenRecV1 carry the same builder with an explicit assignment loop
    (models/diffmm.py:88-107) that yields the identical matrix; route LightGCN through it."""
    import scipy.sparse as sp
    from models.lightgcn import LightGCN

    def get_norm_adj_mat(self):
        A = sp.dok_matrix((self.n_users + self.n_items, self.n_users + self.n_items), dtype=np.float32)
        inter_M = self.interaction_matrix
        inter_M_t = self.interaction_matrix.transpose()
        for r, c in zip(inter_M.row, inter_M.col + self.n_users):
            A[r, c] = 1
        for r, c in zip(inter_M_t.row + self.n_users, inter_M_t.col):
            A[r, c] = 1
        sumArr = (A > 0).sum(axis=1)
        diag = np.power(np.array(sumArr.flatten())[0] + 1e-7, -0.5)
        D = sp.diags(diag)
        L = sp.coo_matrix(D * A * D)
        i = torch.LongTensor(np.array([L.row, L.col]))
        return torch.sparse_coo_tensor(i, torch.FloatTensor(L.data), torch.Size(L.shape))

    LightGCN.get_norm_adj_mat = get_norm_adj_mat


def override_params(model, skip=("running_", "num_batches", "loss_history", "image_embedding.weight",
                                 "text_embedding.weight")):
    """Overwrite every float parameter/buffer with ``synth.make_params`` values keyed by its
    (first, state_dict-order) name, so the other side can regenerate them without the reference.
    Returns the ordered list of (name, shape)."""
    from genmmrec_b200 import synth

    sd = model.state_dict()
    seen = set()
    names = []
    with torch.no_grad():
        for k, v in sd.items():
            if not v.is_floating_point() or any(s in k for s in skip):
                continue
            if v.data_ptr() in seen:
                continue
            seen.add(v.data_ptr())
            val = synth.make_params({k: tuple(v.shape)})[k]
            v.copy_(torch.from_numpy(val))
            names.append((k, tuple(v.shape)))
    return names


def evaluate_capture(trainer, loader):
    """Run the reference's Trainer.evaluate (common/trainer.py:369-388) and also return what it
    hides: the concatenated top-K ids and the unrounded metric matrix."""
    captured = {}
    ev = trainer.evaluator
    orig = ev.evaluate

    def spy(batch_matrix_list, eval_data, is_test=False, idx=0):
        captured["topk"] = torch.cat(batch_matrix_list, dim=0).cpu().numpy()
        return orig(batch_matrix_list, eval_data, is_test=is_test, idx=idx)

    ev.evaluate = spy
    try:
        result = trainer.evaluate(loader)
    finally:
        ev.evaluate = orig
    topk = captured["topk"]
    pos_items = loader.get_eval_items()
    pos_len = loader.get_eval_len_list()
    hit = np.asarray([[i in set(m.tolist()) for i in n] for m, n in zip(pos_items, topk)])
    raw = ev._calculate_metrics(pos_len, hit)
    return result, topk, hit, raw


def coo_parts(t):
    """(indices int64 [2,nnz], values fp32 [nnz]) of a torch sparse COO tensor, as stored."""
    return t._indices().cpu().numpy().astype(np.int64), t._values().cpu().numpy().astype(np.float32)
