"""Model plugin base classes -- the drop-in boundary of the hot path.

Same contract as GenMMRec/src/common/abstract_recommender.py:10-103: a model is an ``nn.Module`` built
as ``Model(config, dataloader)`` that implements ``calculate_loss`` and ``full_sort_predict``;
``GeneralRecommender.__init__`` reads the table sizes from ``dataloader.dataset`` and loads
``image_feat.npy`` / ``text_feat.npy`` onto the device.

Additions used by the fused evaluation (the reference API keeps working without them):
  * ``eval_factors(users)``: the factorised form ``(Eu, user_rows, Ei, bias)`` of the score matrix
    ``full_sort_predict`` would return, so that scoring, masking and top-K run in ONE kernel and the
    ``[B, n_items]`` matrix is never written;
  * ``full_sort_topk(interaction, k)``: batch-level fused path with the reference's batch layout;
  * a propagation cache: the reference re-runs the whole graph propagation for every evaluation
    batch (GenMMRec/src/models/diffmm.py:276); here the propagated embeddings are reused while the
    model is in eval mode under ``no_grad`` and no parameter/graph changed (version counters).
"""
import os

import numpy as np
import torch
import torch.nn as nn

from .. import ops


class AbstractRecommender(nn.Module):
    def pre_epoch_processing(self):
        pass

    def post_epoch_processing(self):
        pass

    def calculate_loss(self, interaction):
        raise NotImplementedError

    def predict(self, interaction):
        raise NotImplementedError

    def full_sort_predict(self, interaction):
        raise NotImplementedError

    def __str__(self):
        params = sum(int(np.prod(p.size())) for p in self.parameters())
        return super().__str__() + "\nTrainable parameters: {}".format(params)


class GeneralRecommender(AbstractRecommender):
    def __init__(self, config, dataloader):
        super(GeneralRecommender, self).__init__()
        self.config = config
        self.USER_ID = config["USER_ID_FIELD"]
        self.ITEM_ID = config["ITEM_ID_FIELD"]
        self.NEG_ITEM_ID = (config["NEG_PREFIX"] or "neg__") + str(self.ITEM_ID)
        self.n_users = dataloader.dataset.get_user_num()
        self.n_items = dataloader.dataset.get_item_num()
        self.batch_size = config["train_batch_size"]
        self.device = torch.device(config["device"])
        if self.device.type != "cuda":
            raise RuntimeError("genmmrec_b200 models run on CUDA only (device=%s): the hot path has no CPU "
                               "fallback" % self.device)
        self.score_precision = config["score_precision"] or "auto"
        self.cache_propagation = config["cache_propagation"] is not False

        self.v_feat, self.t_feat = None, None
        if not config["end2end"] and config["is_multimodal_model"]:
            feats = config["preloaded_features"]  # optional (v_feat, t_feat) tensors, skips the .npy files
            if feats is not None:
                self.v_feat, self.t_feat = (None if f is None else f.to(self.device, torch.float32) for f in feats)
            else:
                dataset_path = os.path.abspath(config["data_path"] + config["dataset"])
                v_path = os.path.join(dataset_path, config["vision_feature_file"])
                t_path = os.path.join(dataset_path, config["text_feature_file"])
                if os.path.isfile(v_path):
                    self.v_feat = torch.from_numpy(np.load(v_path, allow_pickle=True)).type(torch.FloatTensor).to(self.device)
                if os.path.isfile(t_path):
                    self.t_feat = torch.from_numpy(np.load(t_path, allow_pickle=True)).type(torch.FloatTensor).to(self.device)
            assert self.v_feat is not None or self.t_feat is not None, "Features all NONE"
        self._prop_cache = None
        self._graph_version = 0

    # ---- propagation cache -----------------------------------------------------------------------
    def _state_signature(self):
        return (self._graph_version,) + tuple((id(p), p._version) for p in self.parameters())

    def invalidate_cache(self):
        self._graph_version += 1
        self._prop_cache = None

    def cached_propagate(self):
        """``self.propagate()`` -> (user_emb, item_emb), cached across eval batches."""
        usable = self.cache_propagation and not self.training and not torch.is_grad_enabled()
        if not usable:
            return self.propagate()
        sig = self._state_signature()
        if self._prop_cache is None or self._prop_cache[0] != sig:
            self._prop_cache = (sig, self.propagate())
        return self._prop_cache[1]

    def propagate(self):
        """(user embeddings [n_users, D], item embeddings [n_items, D]) that full_sort_predict contracts."""
        raise NotImplementedError

    # ---- scoring ---------------------------------------------------------------------------------
    def eval_factors(self, users):
        """(Eu, user_rows, Ei, bias): scores[b, i] = bias[i] + <Eu[user_rows[b]], Ei[i]>."""
        ue, ie = self.cached_propagate()
        return ue, users, ie, None

    def full_sort_predict(self, interaction):
        """fp32 [B, n_items] score matrix (API parity with the reference; the fused evaluation
        never calls this)."""
        eu, rows, ei, bias = self.eval_factors(interaction[0])
        return ops.scores_dense(eu.contiguous(), ei.contiguous(), users=rows, bias=bias)

    def full_sort_topk(self, interaction, k, return_scores=False):
        """Fused ``full_sort_predict`` -> mask -> ``topk`` for one reference-style batch
        ``[users, mask[2, nnz]]`` (GenMMRec/src/common/trainer.py:381-386).  Returns int64 [B, k]."""
        users, mask = interaction[0], interaction[1]
        eu, rows, ei, bias = self.eval_factors(users)
        b = int(users.numel())
        key = torch.sort(mask[0] * self.n_items + mask[1]).values
        rowptr = torch.searchsorted(key, torch.arange(b + 1, device=key.device, dtype=torch.int64) * self.n_items)
        items = (key % self.n_items).to(torch.int32)
        ids, sc = ops.score_mask_topk(eu.contiguous(), ei.contiguous(), k, users=rows, bias=bias, mask_rowptr=rowptr,
                                      mask_items=items, precision=self.score_precision, return_scores=return_scores)
        return (ids.to(torch.int64), sc) if return_scores else ids.to(torch.int64)
