"""VBPR (GenMMRec/src/models/vbpr.py): no graph; the item side is [id embedding | Linear(cat(t, v))]
(vbpr.py:69-75) and ``full_sort_predict`` is the [B, 128] x [128, n_items] contraction (:100-106),
here fused with masking and top-K (K2)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..common.abstract_recommender import GeneralRecommender
from ..ops import bpr_scores, linear_proj
from ._common import bpr_loss, emb_loss


class VBPR(GeneralRecommender):
    def __init__(self, config, dataloader):
        super(VBPR, self).__init__(config, dataloader)
        self.u_embedding_size = self.i_embedding_size = config["embedding_size"]
        self.reg_weight = config["reg_weight"]
        self.u_embedding = nn.Parameter(nn.init.xavier_uniform_(torch.empty(self.n_users, self.u_embedding_size * 2)))
        self.i_embedding = nn.Parameter(nn.init.xavier_uniform_(torch.empty(self.n_items, self.i_embedding_size)))
        if self.v_feat is not None and self.t_feat is not None:
            self.item_raw_features = torch.cat((self.t_feat, self.v_feat), -1)
        elif self.v_feat is not None:
            self.item_raw_features = self.v_feat
        else:
            self.item_raw_features = self.t_feat
        self.item_linear = nn.Linear(self.item_raw_features.shape[1], self.i_embedding_size)
        nn.init.xavier_normal_(self.item_linear.weight)
        nn.init.constant_(self.item_linear.bias, 0)

    def forward(self, dropout=0.0):
        # [I, 4480] x Linear(4480, 64): the tcgen05 projection kernel under no_grad (ops.linear_proj)
        item_embeddings = torch.cat((self.i_embedding, linear_proj(self.item_raw_features, self.item_linear)), -1)
        return F.dropout(self.u_embedding, dropout), F.dropout(item_embeddings, dropout)

    def propagate(self):
        return self.forward()

    def calculate_loss(self, interaction):
        user, pos_item, neg_item = interaction[0], interaction[1], interaction[2]
        ue, ie = self.forward()
        mf = bpr_loss(*bpr_scores(ue, ie, user, pos_item, neg_item))   # fused gather + dot (csrc/train_ops.cu)
        return mf + self.reg_weight * emb_loss(ue[user, :], ie[pos_item, :], ie[neg_item, :])
