"""Host-side logic of the package on CPU: loaders, config merge, dataset files, ABI surface.
No CUDA computation here (the hot path has no CPU fallback by design)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import REPO, load_golden
from genmmrec_b200 import synth
from genmmrec_b200.utils.configurator import Config
from genmmrec_b200.utils.dataset import RecDataset
from genmmrec_b200.utils.dataloader import EvalDataLoader, TrainDataLoader


def toy_config(**over):
    cfg = Config("DiffMM", "toy", dict({"device": "cpu", "eval_batch_size": 128}, **over))
    return cfg


def toy_splits(cfg, d):
    ds = RecDataset.from_arrays(cfg, d["users"], d["items"], d["label"], 300, 120)
    return ds.split()


def test_eval_loader_layout_matches_reference(toy_data):
    z, _ = load_golden("toy_diffmm")
    cfg = toy_config()
    tr, va, te = toy_splits(cfg, toy_data)
    for split, ds in (("valid", va), ("test", te)):
        ld = EvalDataLoader(cfg, ds, additional_dataset=tr, batch_size=128)
        assert np.array_equal(ld.eval_u.numpy(), z["loader/%s/eval_u" % split])
        assert np.array_equal(ld.pos_items_per_u.numpy(), z["loader/%s/pos_items_per_u" % split])
        assert np.array_equal(np.asarray(ld.train_pos_len_list), z["loader/%s/train_pos_len_list" % split])
        assert np.array_equal(ld.get_eval_len_list(), z["loader/%s/eval_len_list" % split])
        assert np.array_equal(np.concatenate(ld.get_eval_items()), z["loader/%s/eval_items_flat" % split])
        # batches: users slice + mask rebased to the batch (dataloader.py:359-368), twice (cursor reset)
        for _ in range(2):
            pr = 0
            n_batches = 0
            for users, mask in ld:
                ref_mask = z["loader/%s/pos_items_per_u" % split]
                sel = (ref_mask[0] >= pr) & (ref_mask[0] < pr + 128)
                assert np.array_equal(users.numpy(), z["loader/%s/eval_u" % split][pr:pr + 128])
                assert np.array_equal(mask.numpy(), np.stack([ref_mask[0][sel] - pr, ref_mask[1][sel]]))
                pr += 128
                n_batches += 1
            assert n_batches == len(ld)
        # CSR forms consumed by the kernels: same sets, ascending inside a row
        rp, it = ld.mask_rowptr.numpy(), ld.mask_items.numpy()
        ref = z["loader/%s/pos_items_per_u" % split]
        for p in (0, 1, len(rp) - 2):
            assert np.array_equal(it[rp[p]:rp[p + 1]], np.sort(ref[1][ref[0] == p]))
        g_rp, g_it = ld.gt_rowptr.numpy(), ld.gt_items.numpy()
        items = ld.get_eval_items()
        for p in (0, 7, len(g_rp) - 2):
            assert np.array_equal(g_it[g_rp[p]:g_rp[p + 1]], np.sort(items[p]))
        b_rp, b_it = ld.batch_mask_csr(128, 128)
        assert b_rp[0] == 0 and b_rp.numel() == min(128, len(rp) - 1 - 128) + 1
        assert np.array_equal(b_it.numpy(), it[rp[128]:rp[128 + b_rp.numel() - 1]])


def test_eval_loader_requires_train_history(toy_data):
    cfg = toy_config()
    tr, va, _ = toy_splits(cfg, toy_data)
    keep = tr.users != va.users[0]
    tr2 = tr.copy_arrays(tr.users[keep], tr.items[keep])
    with pytest.raises(KeyError):
        EvalDataLoader(cfg, va, additional_dataset=tr2, batch_size=64)
    with pytest.raises(ValueError):
        EvalDataLoader(cfg, va, additional_dataset=None, batch_size=64)


def test_dataset_file_round_trip(tmp_path, toy_data):
    synth.write_dataset(str(tmp_path), "toy", 300, 120, 3600, image_dim=8, text_dim=4)
    cfg = toy_config(data_path=str(tmp_path) + os.sep)
    ds = RecDataset(cfg)
    assert ds.get_user_num() == 300 and ds.get_item_num() == 120 and len(ds) == 3600
    tr, va, te = ds.split()
    m = toy_data["label"] == 0
    assert np.array_equal(tr.users.numpy(), toy_data["users"][m]) and np.array_equal(tr.items.numpy(), toy_data["items"][m])
    loader = TrainDataLoader(cfg, tr, batch_size=256)
    mat = loader.inter_matrix("scipy")
    assert mat.shape == (300, 120) and mat.nnz == m.sum() and (mat.data == 1.0).all()
    assert loader.inter_matrix("coo").nnz == mat.nnz and loader.inter_matrix("csr").shape == (300, 120)
    batch = next(iter(loader))
    assert batch.shape == (3, 256)
    hist = set(zip(tr.users.tolist(), tr.items.tolist()))
    assert all((int(u), int(n)) not in hist for u, n in zip(batch[0], batch[2]))  # negatives are unseen
    with pytest.raises(ValueError):
        RecDataset(toy_config(data_path=str(tmp_path / "missing") + os.sep))


def test_config_merge_order_and_missing_keys(tmp_path):
    cfg = Config("DiffMM", "baby", {"device": "cpu", "n_layers": 3})
    assert cfg["n_layers"] == 3                 # argument dict wins over the model yaml
    assert cfg["embedding_size"] == 64 and cfg["topk"] == [5, 10, 20, 50] and cfg["eval_batch_size"] == 4096
    assert cfg["reg_weight"] == 1e-6 and isinstance(cfg["reg_weight"], float)   # "1.0e-6" parsed as float
    assert cfg["inter_file_name"] == "baby.inter"
    assert cfg["definitely_not_a_key"] is None  # configurator.py:125-129
    assert "seed" in cfg["hyper_parameters"] and cfg["valid_metric_bigger"] is True
    # a foreign configs/ directory in the reference's layout is consumed unchanged
    (tmp_path / "model").mkdir()
    (tmp_path / "dataset").mkdir()
    (tmp_path / "overall.yaml").write_text("embedding_size: 32\ntopk: [10]\nvalid_metric: Recall@10\nreg: 1e-05\n")
    (tmp_path / "model" / "DiffMM.yaml").write_text("embedding_size: 16\nhyper_parameters: ['reg']\n")
    cfg2 = Config("DiffMM", "baby", {"device": "cpu"}, config_dir=str(tmp_path))
    assert cfg2["embedding_size"] == 16 and cfg2["topk"] == [10] and cfg2["reg"] == 1e-5
    assert cfg2["hyper_parameters"] == ["reg", "seed"]


def test_abi_surface():
    """libgmr.so loads and exports every symbol include/gmr.h declares; the ctypes table covers them."""
    from genmmrec_b200 import _lib
    header = open(os.path.join(REPO, "include", "gmr.h")).read()
    declared = set(re.findall(r"\b(gmr_[a-z0-9_]+)\s*\(", header))
    declared -= {"gmr_spmm_plan"}
    assert declared, "no declarations parsed"
    assert os.path.exists(_lib.LIB_PATH), "libgmr.so is not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libgmr.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().gmr_abi_version() == 1


def test_every_exported_symbol_is_documented_with_its_reference_call_site():
    """INTEGRATION.md names every entry point of include/gmr.h (its table maps each one to the reference call site it
    stands in for) and every GMR_* environment switch the package or bench.py reads."""
    header = open(os.path.join(REPO, "include", "gmr.h")).read()
    doc = open(os.path.join(REPO, "INTEGRATION.md")).read()
    declared = set(re.findall(r"\b(gmr_[a-z0-9_]+)\s*\(", header)) - {"gmr_spmm_plan"}

    def covered(sym):
        # families are written as `gmr_peer_alloc / _free / _export`: accept the suffix form after the family's first member
        if sym in doc:
            return True
        parts = sym.split("_")
        return any("_".join(parts[:i]) in doc and ("_" + "_".join(parts[i:])) in doc for i in range(2, len(parts)))

    missing = sorted(s for s in declared if not covered(s))
    assert not missing, missing
    switches = set()
    for root, _, files in os.walk(os.path.join(REPO, "generative-multimodal-recommendation_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                switches |= set(re.findall(r"(?:getenv\(|environ\.get\(|environ\[)\s*\"(GMR_[A-Z0-9_]+)\"", open(os.path.join(root, f)).read()))
    switches |= set(re.findall(r"environ\.get\(\s*\"(GMR_[A-Z0-9_]+)\"", open(os.path.join(REPO, "bench.py")).read()))
    undocumented = sorted(s for s in switches if s not in doc)
    assert not undocumented, undocumented


def test_models_refuse_cpu(toy_data):
    """No CPU fallback: building a model on a CPU device fails loudly."""
    from genmmrec_b200.models.lightgcn import LightGCN
    cfg = Config("LightGCN", "toy", {"device": "cpu", "n_layers": 2, "is_multimodal_model": False})
    tr, _, _ = toy_splits(cfg, toy_data)
    with pytest.raises(RuntimeError, match="CUDA only"):
        LightGCN(cfg, TrainDataLoader(cfg, tr, batch_size=64))


def test_registry():
    from genmmrec_b200.utils.utils import get_model, get_trainer
    for name in ("DiffMM", "GUME", "GenRecV1", "LD4MRec", "VBPR", "LightGCN"):
        assert get_model(name).__name__ == name
    assert get_trainer("DiffMM").__name__ == "Trainer"


def test_popularity_groups_follow_quick_start():
    """quick_start.py:46-92: top 20 % of the train items by count are 'popular', users with > 5 train rows are 'warm'."""
    from genmmrec_b200.utils.utils import set_popularity_groups

    class _Train:
        pass

    rng = np.random.default_rng(8)
    tr = _Train()
    tr.items = torch.from_numpy(rng.zipf(1.6, size=4000).clip(max=300).astype(np.int64) - 1)
    tr.users = torch.from_numpy(rng.integers(0, 500, size=4000))
    cfg = set_popularity_groups({}, tr)
    counts = np.bincount(tr.items.numpy())
    seen = np.flatnonzero(counts)
    pop = np.array(sorted(cfg["pop_items"]))
    assert len(pop) == int(len(seen) * 0.2) and set(pop) <= set(seen)
    rest = np.setdiff1d(seen, pop)
    assert counts[pop].min() >= counts[rest].max()
    ucounts = np.bincount(tr.users.numpy(), minlength=500)
    assert cfg["warm_users"] == set(np.flatnonzero(ucounts > 5).tolist())


def test_early_stopping_and_train_loader_negatives(toy_data):
    """Host logic of the training loop (GenMMRec/src/utils/utils.py:70-113, utils/dataloader.py:226-275): the early-stopping
    rule, and one uniform negative per row that is never in the user's train history."""
    from genmmrec_b200.common.trainer import Trainer
    es = Trainer.early_stopping
    assert es(0.5, 0.4, 3, 2) == (0.5, 0, False, True)            # improved: counter resets
    assert es(0.3, 0.4, 1, 2) == (0.4, 2, False, False)           # not improved, still within patience
    assert es(0.3, 0.4, 2, 2) == (0.4, 3, True, False)            # patience exceeded
    assert es(0.3, 0.4, 0, 2, bigger=False) == (0.3, 0, False, True)
    cfg = Config("LightGCN", "toy", {"device": "cpu", "is_multimodal_model": False})
    tr, _, _ = toy_splits(cfg, toy_data)
    loader = TrainDataLoader(cfg, tr, batch_size=256, shuffle=True)
    hist = set(zip(tr.users.tolist(), tr.items.tolist()))
    seen = 0
    for inter in loader:
        assert inter.shape[0] == 3 and inter.dtype == torch.int64
        u, p, n = inter.tolist()
        assert all((a, b) in hist for a, b in zip(u, p))          # positives are train interactions
        assert not any((a, b) in hist for a, b in zip(u, n))      # negatives never are
        seen += len(u)
    assert seen == len(tr)


def test_device_edge_extraction_equals_reference_loop():
    """DiffMM.edges_from_denoised (vectorised) against the per-element loop of common/trainer.py:548-553, on CPU tensors."""
    from genmmrec_b200.models.diffmm import DiffMM
    g = torch.Generator().manual_seed(5)
    batch_index = torch.randperm(300, generator=g)[:17]
    den = torch.randn(17, 40, generator=g)
    u, i = DiffMM.edges_from_denoised(batch_index, den, 4)
    _, idx = torch.topk(den, k=4)
    u_ref = [int(batch_index[a]) for a in range(17) for _ in range(4)]
    i_ref = [int(idx[a][b]) for a in range(17) for b in range(4)]
    assert u.tolist() == u_ref and i.tolist() == i_ref


def test_blocked_spmm_dispatch_policy(monkeypatch):
    """ops._block_cols_for: K1b is opt-in; `auto` blocks only tables beyond GMR_SPMM_BLOCK_MIN_MB; the slice size follows
    GMR_SPMM_BLOCK_MB; unaligned operands and widths that are not multiples of 4 stay on K1."""
    from genmmrec_b200 import ops

    class G:
        shape = (1000, 2_000_000)

    class T:
        def __init__(self, rows, stride, ptr=256):
            self.shape, self._s, self._p = (rows, 64), stride, ptr

        def stride(self, k):
            return self._s

        def data_ptr(self):
            return self._p

    x, y = T(2_000_000, 64), T(1000, 64)
    monkeypatch.delenv("GMR_SPMM_BLOCKED", raising=False)
    assert ops._block_cols_for(G, 64, x, y) is None                      # default: row-centric K1
    monkeypatch.setenv("GMR_SPMM_BLOCKED", "auto")
    assert ops._block_cols_for(G, 64, x, y) == (48 << 20) // 256          # 512 MB table: 48 MB slices
    monkeypatch.setenv("GMR_SPMM_BLOCK_MB", "32")
    assert ops._block_cols_for(G, 64, x, y) == (32 << 20) // 256
    assert ops._block_cols_for(G, 62, x, y) is None                       # width not a multiple of 4
    assert ops._block_cols_for(G, 64, T(2_000_000, 64, ptr=260), y) is None   # misaligned X
    G.shape = (1000, 100_000)                                             # 25.6 MB table: below the 96 MB threshold
    assert ops._block_cols_for(G, 64, x, y) is None
    monkeypatch.setenv("GMR_SPMM_BLOCKED", "1")
    assert ops._block_cols_for(G, 64, x, y) == 100_000                    # forced: one block covers the table
