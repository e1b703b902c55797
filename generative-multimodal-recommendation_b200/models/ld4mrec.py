"""LD4MRec -- the sparse/dense hot spots of GenMMRec/src/models/ld4mrec.py on the B200 kernels.

Three pieces of this model sit on the accelerated path:
  * the one-off wide SpMM ``user_mm_emb = R_norm . cat(v_feat, t_feat)`` (ld4mrec.py:161-206,
    D = 4096 + 384) -> K1 with column blocking;
  * ``CNet.item_proj`` applied to the dense history rows ``x_in`` (ld4mrec.py:22,42,357-359): the
    reference materialises ``[B, n_items]`` from a scipy CSR slice on the HOST and multiplies it
    densely; a binary row times ``W^T`` is exactly a sparse gather-sum, so it is K1 on the binary
    train matrix R against ``W^T`` ([n_items, hidden]);
  * ``CNet.output_proj`` (ld4mrec.py:36,54), the score contraction ``[B, hidden] x [hidden, n_items]
    + bias`` -> fused with masking and top-K (K2 with a bias vector).
The small conditional residual blocks in between stay in torch.  The SVD user encoder
(ld4mrec.py:138-159) and the diffusion training loss are outside the hot path: ``user_svd_emb`` is a
buffer the caller may overwrite (it is computed with the same scipy call by default).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..common.abstract_recommender import GeneralRecommender
from .. import graph as gb
from ..ops import GraphCSR, spmm, spmm_raw


class ConditionalBlock(nn.Module):
    """Pre-norm residual MLP block with FiLM-style (1 + scale, shift) conditioning (ld4mrec.py:56-87)."""

    def __init__(self, hidden_size, dropout):
        super(ConditionalBlock, self).__init__()
        self.norm1 = nn.LayerNorm(hidden_size)
        self.linear1 = nn.Linear(hidden_size, hidden_size)
        self.act = nn.GELU()
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(hidden_size, hidden_size)
        self.cond_scale = nn.Linear(hidden_size, hidden_size)
        self.cond_shift = nn.Linear(hidden_size, hidden_size)

    def forward(self, x, cond):
        h = self.norm1(x) * (1 + self.cond_scale(cond)) + self.cond_shift(cond)
        return x + self.linear2(self.dropout(self.act(self.linear1(h))))


class CNet(nn.Module):
    def __init__(self, n_items, hidden_size, cond_dim, n_layers=3, dropout=0.1):
        super(CNet, self).__init__()
        self.n_items, self.hidden_size = n_items, hidden_size
        self.item_proj = nn.Linear(n_items, hidden_size)
        self.cond_proj = nn.Linear(cond_dim, hidden_size)
        self.time_proj = nn.Linear(hidden_size, hidden_size)
        self.layers = nn.ModuleList([ConditionalBlock(hidden_size, dropout) for _ in range(n_layers)])
        self.output_proj = nn.Linear(hidden_size, n_items)

    def hidden(self, h_items, t_emb, condition):
        """Everything between item_proj and output_proj; ``h_items`` = item_proj(x_t) computed sparsely."""
        g = self.cond_proj(condition) + self.time_proj(t_emb)
        h = h_items
        for layer in self.layers:
            h = layer(h, g)
        return h

    def forward(self, x_t, t_emb, condition):
        return self.output_proj(self.hidden(self.item_proj(x_t), t_emb, condition))


class LD4MRec(GeneralRecommender):
    def __init__(self, config, dataset):
        super(LD4MRec, self).__init__(config, dataset)
        self.embedding_size = config["embedding_size"]
        self.svd_k = config["svd_k"]
        self.cnet_hidden = config["cnet_hidden_size"]
        self.cnet_layers = config["cnet_n_layers"]
        self.interaction_matrix = dataset.inter_matrix(form="coo")
        m = self.interaction_matrix
        self.R = GraphCSR.from_coo(*gb.binary_r(m.row, m.col, self.n_users, self.n_items, device=self.device), self.device)
        self._init_svd()
        self._init_multimodal()
        self.mm_dim = (self.v_feat.shape[1] if self.v_feat is not None else 0) + \
                      (self.t_feat.shape[1] if self.t_feat is not None else 0)
        self.mm_project = nn.Linear(self.mm_dim, self.embedding_size) if self.mm_dim > 0 else None
        cond_dim = self.svd_k + (self.embedding_size if self.mm_dim > 0 else 0)
        self.cnet = CNet(self.n_items, self.cnet_hidden, cond_dim, self.cnet_layers, config["dropout"] or 0.1)
        self.time_emb_dim = self.cnet_hidden
        self.t_in = nn.Parameter(torch.zeros(1))

    def _init_svd(self):
        if self.config["skip_svd"]:
            self.user_svd_emb = torch.zeros(self.n_users, self.svd_k, device=self.device)
            return
        from scipy.sparse.linalg import svds
        u, s, _ = svds(self.interaction_matrix.to_scipy().astype(np.float32), k=self.svd_k)
        u, s = u[:, ::-1], s[::-1]
        self.user_svd_emb = torch.from_numpy(np.ascontiguousarray(u * np.sqrt(s))).float().to(self.device)

    def _init_multimodal(self):
        feats = [f for f in (self.v_feat, self.t_feat) if f is not None]
        self.user_mm_emb = None
        if feats:
            m = self.interaction_matrix
            rn = gb.ld4mrec_rnorm(m.row, m.col, self.n_users, self.n_items, device=self.device)
            self.R_norm = GraphCSR.from_coo(*rn, self.device)
            self.user_mm_emb = spmm_raw(self.R_norm, torch.cat(feats, dim=1).contiguous())

    def get_time_embedding(self, timesteps):
        half = self.time_emb_dim // 2
        freq = torch.exp(torch.arange(half, dtype=torch.float32, device=self.device) * -np.log(10000.0) / (half - 1))
        ang = timesteps[:, None].float() * freq[None, :]
        emb = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)
        if self.time_emb_dim % 2 == 1:
            emb = F.pad(emb, (0, 1))
        return emb

    def item_proj_all(self):
        """item_proj applied to EVERY user's binary history row: R . W^T + b  ([n_users, hidden])."""
        w_t = self.cnet.item_proj.weight.t().contiguous()
        return spmm(self.R, w_t) + self.cnet.item_proj.bias

    def hidden_states(self, user):
        h_items = self.item_proj_all()[user]
        t_emb = self.get_time_embedding(torch.abs(self.t_in).expand(len(user)))
        u_mm = self.mm_project(self.user_mm_emb[user]) if self.mm_project is not None else None
        cond = torch.cat([self.user_svd_emb[user], u_mm], dim=1) if u_mm is not None else self.user_svd_emb[user]
        return self.cnet.hidden(h_items, t_emb, cond)

    def eval_factors(self, users):
        h = self.hidden_states(users)
        return h, None, self.cnet.output_proj.weight, self.cnet.output_proj.bias

    def propagate(self):
        raise NotImplementedError("LD4MRec scores through eval_factors (per-user hidden states), not a user table")
