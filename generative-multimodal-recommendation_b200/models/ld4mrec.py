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
(ld4mrec.py:138-159) is outside the hot path: ``user_svd_emb`` is a buffer the caller may overwrite (it is
computed with the same scipy call by default).  ``calculate_loss`` is the reference's diffusion objective
(ld4mrec.py:265-344) kept on the device end to end: history rows expanded from the CSR with one scatter instead
of a host scipy slice + upload, steps drawn with ``torch.multinomial`` from the device-resident loss history,
and the history's per-sample moving average applied in closed form instead of a ``.item()`` loop.
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
        self.steps = int(config["steps"] or 100)
        self.smoothing_gamma = float(config["smoothing_gamma"] if config["smoothing_gamma"] is not None else 0.1)
        self.register_buffer("loss_history", torch.ones(self.steps))   # importance sampling of the diffusion step
        self._init_noise_schedule(float(config["min_noise_level"] if config["min_noise_level"] is not None else 0.001))
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

    def _init_noise_schedule(self, alpha_min):
        """ld4mrec.py:208-239: 1 - alpha_bar_t linear in t, betas clamped to [1e-4, 0.9999], alpha_bar = cumprod."""
        t = torch.arange(1, self.steps + 1, dtype=torch.float32, device=self.device)
        alpha_bar = 1 - (alpha_min + (t - 1) / (self.steps - 1) * (1 - alpha_min))
        ratio = alpha_bar / torch.cat([torch.ones(1, device=self.device), alpha_bar[:-1]])
        self.betas = torch.clamp(1 - ratio, min=0.0001, max=0.9999)
        self.alphas = 1 - self.betas
        self.alpha_bar = torch.cumprod(self.alphas, dim=0)

    def history_rows(self, user):
        """Dense binary train rows [B, n_items] of ``user`` expanded from the device CSR (ld4mrec.py:283-292 slices a
        scipy CSR on the host and uploads the result)."""
        rp, col = self.R.rowptr.to(torch.int64), self.R.col.to(torch.int64)
        user = user.to(torch.int64)
        lens = rp[user + 1] - rp[user]
        starts = torch.cumsum(lens, 0) - lens
        rows = torch.repeat_interleave(torch.arange(user.numel(), device=user.device), lens)
        src = torch.repeat_interleave(rp[user] - starts, lens) + torch.arange(rows.numel(), device=user.device)
        x = torch.zeros(user.numel(), self.n_items, device=user.device)
        x[rows, col[src]] = 1.0
        return x

    def q_sample(self, x_start, t, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        ab = self.alpha_bar[t].unsqueeze(1)
        return torch.sqrt(ab) * x_start + torch.sqrt(1 - ab) * noise, noise

    @torch.no_grad()
    def _update_loss_history(self, t, losses):
        """h[s] <- 0.9 h[s] + 0.1 loss_i applied once per sample in batch order (ld4mrec.py:338-342), in closed form:
        a step drawn m times ends at 0.9^m h + 0.1 * sum_j 0.9^(m-1-j) loss_j (j = order of appearance)."""
        order = torch.argsort(t, stable=True)
        ts, ls = t[order], losses[order].to(torch.float32)
        counts = torch.bincount(ts, minlength=self.steps)
        first = torch.cumsum(counts, 0) - counts
        rank = torch.arange(ts.numel(), device=t.device) - first[ts]
        w = 0.1 * torch.pow(torch.full_like(ls, 0.9), (counts[ts] - 1 - rank).to(torch.float32))
        add = torch.zeros_like(self.loss_history).index_add_(0, ts, w * ls)
        self.loss_history.mul_(torch.pow(torch.full_like(self.loss_history, 0.9), counts.to(torch.float32))).add_(add)

    def calculate_loss(self, interaction, t=None, noise=None):
        """Diffusion objective of ld4mrec.py:265-344: importance-sampled step, q-sample of the history row, x0 prediction
        by the C-Net, MSE against the label-smoothed history.  ``t`` / ``noise`` may be supplied for reproducibility."""
        user = interaction[0]
        x_in = self.history_rows(user)
        gamma = self.smoothing_gamma
        target = x_in * (1 - gamma) + (1 - x_in) * gamma
        if t is None:
            probs = torch.sqrt(self.loss_history ** 2)
            t = torch.multinomial(probs / probs.sum(), user.numel(), replacement=True)
        x_t, _ = self.q_sample(x_in, t, noise)
        u_mm = self.mm_project(self.user_mm_emb[user]) if self.mm_project is not None else None
        cond = torch.cat([self.user_svd_emb[user], u_mm], dim=1) if u_mm is not None else self.user_svd_emb[user]
        pred = self.cnet(x_t, self.get_time_embedding(t), cond)
        loss = F.mse_loss(pred, target, reduction="none").mean(dim=1)
        self._update_loss_history(t, loss.detach())
        return loss.mean()

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
