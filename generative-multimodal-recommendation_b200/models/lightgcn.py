"""LightGCN -- the canonical propagation layer of the hot path (GenMMRec/src/models/lightgcn.py).

``forward()`` = mean over l = 0..n_layers of A_hat^l [U0; I0] (lightgcn.py:115-127), with every
``torch.sparse.mm`` replaced by the CSR SpMM kernel K1 (csrc/spmm.cu); ``full_sort_predict`` contracts
the propagated user rows with all item rows (lightgcn.py:156-164) -- fused with masking and top-K in
``Trainer.evaluate``.
"""
import torch
import torch.nn as nn

from ..common.abstract_recommender import GeneralRecommender
from .. import graph as gb
from ..ops import bpr_scores
from ._common import BipartiteAdj, bpr_loss, emb_loss


class LightGCN(GeneralRecommender):
    def __init__(self, config, dataset):
        super(LightGCN, self).__init__(config, dataset)
        self.interaction_matrix = dataset.inter_matrix(form="coo")
        self.latent_dim = config["embedding_size"]
        n_layers = config["n_layers"]
        self.n_layers = n_layers[0] if isinstance(n_layers, list) else n_layers
        rw = config["reg_weight"]
        self.reg_weight = rw[0] if isinstance(rw, list) else rw
        init = nn.init.xavier_uniform_
        self.embedding_dict = nn.ParameterDict({
            "user_emb": nn.Parameter(init(torch.empty(self.n_users, self.latent_dim))),
            "item_emb": nn.Parameter(init(torch.empty(self.n_items, self.latent_dim))),
        })
        self.norm_adj_matrix = self.get_norm_adj_mat()

    def get_norm_adj_mat(self):
        m = self.interaction_matrix
        idx, val, _ = gb.norm_adj(m.row, m.col, self.n_users, self.n_items, device=self.device)
        return BipartiteAdj(idx, val, self.n_users, self.n_items, self.device)

    def get_ego_embeddings(self):
        return torch.cat([self.embedding_dict["user_emb"], self.embedding_dict["item_emb"]], 0)

    def forward(self):
        e = self.get_ego_embeddings()
        acc = e
        for _ in range(self.n_layers):
            e = self.norm_adj_matrix.mm(e)
            acc = acc + e
        out = acc / float(self.n_layers + 1)
        return out[:self.n_users, :], out[self.n_users:, :]

    def propagate(self):
        return self.forward()

    def calculate_loss(self, interaction):
        user, pos_item, neg_item = interaction[0], interaction[1], interaction[2]
        ua, ia = self.forward()
        mf = bpr_loss(*bpr_scores(ua, ia, user, pos_item, neg_item))   # fused gather + dot (csrc/train_ops.cu)
        reg = emb_loss(self.embedding_dict["user_emb"][user, :], self.embedding_dict["item_emb"][pos_item, :],
                       self.embedding_dict["item_emb"][neg_item, :])
        return mf + self.reg_weight * reg
