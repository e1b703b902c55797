"""DiffMM -- propagation + full-sort scoring of GenMMRec/src/models/diffmm.py on the B200 kernels.

Hot path (diffmm.py:129-169, ``forward_MM``): 6 + n_layers ``torch.spmm`` calls over the normalised
user-item adjacency ``A`` and the two diffusion-generated modality graphs, then
``matmul(usr[user], itm.T)`` (:276-278).  Under ``no_grad`` the propagation is restructured around the
bipartite blocks of ``A``:

    [H_v | H_t | G ]_users = R_hat  . [F_v | F_t | I0]      one 192-wide pass over the user rows
    [G_v | G_t | H ]_items = R_hat' . [H_v | H_t | U0]      one 192-wide pass over the item rows

which yields every H_m / G_m block of the reference's four 64-wide full-matrix SpMMs (two of whose
halves it computes twice) with bit-identical rows, in 2 kernel launches that read the CSR once per
half.  With autograd enabled (training) the literal sequence of the reference is used, through the
differentiable ``ops.spmm``.

Out of scope here: the diffusion / denoise networks and their trainer.  The
modality graphs are attributes the caller sets (``image_UI_matrix`` / ``text_UI_matrix``), exactly as
the reference's trainer does (common/trainer.py:564-576); ``set_generated_edges`` builds them on the
device from generated (user, item) edges.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..common.abstract_recommender import GeneralRecommender
from .. import graph as gb
from ..ops import GraphCSR, spmm, spmm_raw
from ._common import BipartiteAdj, as_graph

init = nn.init.xavier_uniform_


class SpAdjDropEdge(nn.Module):
    """Random edge keep with probability keepRate, kept values / keepRate (diffmm.py:287-301).
    Works on COO parts or a torch sparse tensor; returns a GraphCSR."""

    def __init__(self, keepRate):
        super(SpAdjDropEdge, self).__init__()
        self.keepRate = keepRate

    def forward(self, adj, device=None):
        if torch.is_tensor(adj):
            idx, val, shape = adj._indices(), adj._values(), tuple(adj.shape)
        else:
            idx, val, shape = adj
        device = device if device is not None else val.device
        idx, val = gb.drop_edges(idx, val, self.keepRate)
        return GraphCSR.from_coo(idx, val, shape, device)


class DiffMM(GeneralRecommender):
    def __init__(self, config, dataset):
        super(DiffMM, self).__init__(config, dataset)
        self.latdim = config["embedding_size"]
        self.gnn_layer = config["n_layers"]
        self.keepRate = config["keep_rate"]
        self.trans = config["trans_type"] or 0
        self.ris_adj_lambda = config["ris_adj_lambda"]
        self.ris_lambda = config["ris_lambda"]
        self.cl_method = config["cl_method"]
        self.ssl_reg = config["ssl_reg"]
        self.temp = config["temperature"]
        self.reg_weight = config["reg_weight"]
        self.rebuild_k = config["rebuild_k"]
        if self.trans != 0:
            raise NotImplementedError("trans_type %r: the shipped DiffMM.yaml uses 0" % self.trans)

        self.uEmbeds = nn.Parameter(init(torch.empty(self.n_users, self.latdim)))
        self.iEmbeds = nn.Parameter(init(torch.empty(self.n_items, self.latdim)))
        self.edgeDropper = SpAdjDropEdge(self.keepRate)
        self.image_feat_dim = self.v_feat.shape[1] if self.v_feat is not None else 0
        self.text_feat_dim = self.t_feat.shape[1] if self.t_feat is not None else 0
        self.image_trans = nn.Parameter(init(torch.empty(size=(self.image_feat_dim, self.latdim))))
        self.text_trans = nn.Parameter(init(torch.empty(size=(self.text_feat_dim, self.latdim))))
        self.modal_weight = nn.Parameter(torch.Tensor([0.5, 0.5]))
        self.softmax = nn.Softmax(dim=0)
        self.leakyrelu = nn.LeakyReLU(0.2)

        self.image_UI_matrix = None
        self.text_UI_matrix = None
        m = dataset.inter_matrix(form="coo")
        self.norm_adj = self.get_norm_adj_mat(m)

    def __setattr__(self, name, value):
        if name in ("image_UI_matrix", "text_UI_matrix", "norm_adj") and "_graph_version" in self.__dict__:
            self.__dict__["_graph_version"] += 1  # a new graph invalidates the cached propagation
        super().__setattr__(name, value)

    # ---- graphs ----------------------------------------------------------------------------------
    def get_norm_adj_mat(self, interaction_matrix):
        idx, val, _ = gb.norm_adj(interaction_matrix.row, interaction_matrix.col, self.n_users, self.n_items,
                                  device=self.device)
        return BipartiteAdj(idx, val, self.n_users, self.n_items, self.device)

    def build_ui_matrix(self, u_list, i_list):
        """The trainer's buildUIMatrix (common/trainer.py:471-485) on the device; returns COO parts."""
        return gb.ui_matrix(u_list, i_list, self.n_users, self.n_items, device=self.device)

    def set_generated_edges(self, image_edges, text_edges):
        """(u, i) arrays of the edges rebuilt from the denoised interactions -> normalised, edge-dropped
        modality graphs (common/trainer.py:564-576)."""
        self.image_UI_matrix = self.edgeDropper(self.build_ui_matrix(*image_edges), self.device)
        self.text_UI_matrix = self.edgeDropper(self.build_ui_matrix(*text_edges), self.device)

    # ---- feature projections (dense, torch/cuBLAS: SURVEY.md a8) ------------------------------------
    def getItemEmbeds(self):
        return self.iEmbeds

    def getUserEmbeds(self):
        return self.uEmbeds

    def getImageFeats(self):
        return self.leakyrelu(torch.mm(self.v_feat, self.image_trans))

    def getTextFeats(self):
        return self.leakyrelu(torch.mm(self.t_feat, self.text_trans))

    # ---- propagation -----------------------------------------------------------------------------
    def forward_MM(self, adj, image_adj, text_adj):
        if (not torch.is_grad_enabled()) and isinstance(adj, BipartiteAdj) and adj.ui is not None:
            return self._forward_mm_fused(adj, as_graph(image_adj), as_graph(text_adj))
        return self._forward_mm_literal(as_graph(adj), as_graph(image_adj), as_graph(text_adj))

    def _forward_mm_fused(self, adj, image_adj, text_adj):
        nu, d = self.n_users, self.latdim
        u0, i0 = self.uEmbeds.detach(), self.iEmbeds.detach()
        weight = self.softmax(self.modal_weight)
        xi = torch.empty((self.n_items, 3 * d), dtype=torch.float32, device=self.device)
        xi[:, 0:d] = F.normalize(self.getImageFeats())
        xi[:, d:2 * d] = F.normalize(self.getTextFeats())
        xi[:, 2 * d:] = i0
        yu = spmm_raw(adj.ui, xi)                       # [U, 3d] = H_v | H_t | G   (user rows)
        xu = torch.cat([yu[:, :2 * d], u0], dim=1)      # [U, 3d] = H_v | H_t | U0
        yi = spmm_raw(adj.iu, xu)                       # [I, 3d] = G_v | G_t | H   (item rows)
        e0 = torch.cat([u0, i0])
        p_img = spmm_raw(image_adj, e0)
        p_txt = spmm_raw(text_adj, e0)

        def combine(h_u, g_u, g_i, h_i, p):
            e = torch.cat([h_u + g_u, h_i + g_i])
            return e + self.ris_adj_lambda * p

        e_img = combine(yu[:, 0:d], yu[:, 2 * d:], yi[:, 0:d], yi[:, 2 * d:], p_img)
        e_txt = combine(yu[:, d:2 * d], yu[:, 2 * d:], yi[:, d:2 * d], yi[:, 2 * d:], p_txt)
        modal = weight[0] * e_img + weight[1] * e_txt
        embeds = modal
        last = modal
        for _ in range(self.gnn_layer):
            last = spmm_raw(adj.full, last)
            embeds = embeds + last
        embeds = embeds + self.ris_lambda * F.normalize(modal)
        return embeds[:nu], embeds[nu:]

    def _forward_mm_literal(self, adj, image_adj, text_adj):
        nu = self.n_users
        image_feats, text_feats = self.getImageFeats(), self.getTextFeats()
        weight = self.softmax(self.modal_weight)

        def branch(m_adj, feats):
            e_adj = spmm(m_adj, torch.concat([self.uEmbeds, self.iEmbeds]))
            e = spmm(adj, torch.concat([self.uEmbeds, F.normalize(feats)]))
            e_ = spmm(adj, torch.concat([e[:nu], self.iEmbeds]))
            return (e + e_) + self.ris_adj_lambda * e_adj

        e_img = branch(image_adj, image_feats)
        e_txt = branch(text_adj, text_feats)
        modal = weight[0] * e_img + weight[1] * e_txt
        lst = [modal]
        for _ in range(self.gnn_layer):
            lst.append(spmm(adj, lst[-1]))
        embeds = sum(lst) + self.ris_lambda * F.normalize(modal)
        return embeds[:nu], embeds[nu:]

    def forward_cl_MM(self, adj, image_adj, text_adj):
        adj, image_adj, text_adj = as_graph(adj), as_graph(image_adj), as_graph(text_adj)
        nu = self.n_users

        def view(m_adj, feats):
            e = spmm(m_adj, torch.concat([self.uEmbeds, F.normalize(feats)]))
            lst = [e]
            for _ in range(self.gnn_layer):
                lst.append(spmm(adj, lst[-1]))
            return sum(lst)

        e1 = view(image_adj, self.getImageFeats())
        e2 = view(text_adj, self.getTextFeats())
        return e1[:nu], e1[nu:], e2[:nu], e2[nu:]

    def propagate(self):
        if self.image_UI_matrix is None or self.text_UI_matrix is None:
            raise RuntimeError("DiffMM: image_UI_matrix / text_UI_matrix are not set (the reference's trainer "
                               "builds them every epoch, common/trainer.py:564-576)")
        return self.forward_MM(self.norm_adj, self.image_UI_matrix, self.text_UI_matrix)

    # ---- training objective (differentiable through ops.spmm) ----------------------------------------
    def reg_loss(self):
        return self.uEmbeds.norm(2).square() + self.iEmbeds.norm(2).square()

    def contrastLoss(self, embeds1, embeds2, nodes, temp):
        embeds1 = F.normalize(embeds1 + 1e-8, p=2)
        embeds2 = F.normalize(embeds2 + 1e-8, p=2)
        p1, p2 = embeds1[nodes], embeds2[nodes]
        nume = torch.exp(torch.sum(p1 * p2, dim=-1) / temp)
        deno = torch.exp(p1 @ embeds2.T / temp).sum(-1)
        return -torch.log(nume / deno).mean()

    def calculate_loss(self, interaction):
        users, pos_items, neg_items = interaction[0], interaction[1], interaction[2]
        if self.image_UI_matrix is None or self.text_UI_matrix is None:
            return torch.tensor(0.0, requires_grad=True).to(self.device)
        usr, itm = self.forward_MM(self.norm_adj, self.image_UI_matrix, self.text_UI_matrix)
        anc, pos, neg = usr[users], itm[pos_items], itm[neg_items]
        bpr = -torch.log(1e-10 + torch.sigmoid((anc * pos).sum(dim=1) - (anc * neg).sum(dim=1))).mean()
        reg = self.reg_loss() * self.reg_weight
        u1, i1, u2, i2 = self.forward_cl_MM(self.norm_adj, self.image_UI_matrix, self.text_UI_matrix)
        if self.cl_method == 1:
            cl = (self.contrastLoss(usr, u1, users, self.temp) + self.contrastLoss(itm, i1, pos_items, self.temp)
                  + self.contrastLoss(usr, u2, users, self.temp) + self.contrastLoss(itm, i2, pos_items, self.temp)) * self.ssl_reg
        else:
            cl = (self.contrastLoss(u1, u2, users, self.temp) + self.contrastLoss(i1, i2, pos_items, self.temp)) * self.ssl_reg
        return bpr + reg + cl
