"""GenRecV1 -- propagation + full-sort scoring of GenMMRec/src/models/genrecv1.py on the B200 kernels.

``full_sort_predict`` (genrecv1.py:417-427) contracts only the *content* embedding
``softmax(w) . [mean_l A^l E0, mean_l A_gen^l E0]`` (:255-264,330-337); the item-item branch
(:266-353, kNN graphs, ``R`` products, BatchNorm gates) feeds ``side_embedding`` which evaluation
discards.  ``propagate`` therefore runs the 2 * n_layers user-item SpMMs only; ``forward`` still returns
both embeddings for callers that want the reference's full output.

Out of scope: the flip diffusion, the denoise transformer, the
clustering/debiasing pre-steps and their trainer.  The generated graph ``image_UI_matrix`` and the kNN
graphs are attributes the caller sets, as the reference's trainer does (common/trainer.py:676-687,
785-789); ``set_generated_edges`` / ``build_item_item_matrices`` build them on the device.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..common.abstract_recommender import GeneralRecommender
from .. import graph as gb
from ..ops import GraphCSR, bpr_scores, spmm
from ._common import BipartiteAdj, as_graph
from .diffmm import SpAdjDropEdge


class GenRecV1(GeneralRecommender):
    def __init__(self, config, dataset):
        super(GenRecV1, self).__init__(config, dataset)
        self.latdim = config["embedding_size"]
        self.n_layers = config["n_layers"]
        self.keep_rate = config["keep_rate"]
        self.rebuild_k = config["rebuild_k"]
        self.knn_k = config["knn_k"]
        self.reg_weight = config["reg_weight"]
        self.ssl_reg1, self.ssl_reg2 = config["ssl_reg1"] or 0.0, config["ssl_reg2"] or 0.0
        self.temp = config["temperature"] or 1.0
        self.knn_builder = config["knn_builder"] or "fused"
        self.image_embedding = self.v_feat
        self.text_embedding = self.t_feat
        self.sparse = True

        m = dataset.inter_matrix(form="coo")
        idx, val, _ = gb.norm_adj(m.row, m.col, self.n_users, self.n_items, device=self.device)
        self.norm_adj = BipartiteAdj(idx, val, self.n_users, self.n_items, self.device)
        self.R = GraphCSR.from_coo(*gb.binary_r(m.row, m.col, self.n_users, self.n_items, device=self.device), self.device)
        self.edgeDropper = SpAdjDropEdge(self.keep_rate)

        d = self.latdim
        self.origin_weight = nn.Parameter(torch.ones(1))
        self.generation_weight = nn.Parameter(torch.ones(1))
        self.user_embedding = nn.Embedding(self.n_users, d)
        self.item_id_embedding = nn.Embedding(self.n_items, d)
        nn.init.xavier_uniform_(self.user_embedding.weight)
        nn.init.xavier_uniform_(self.item_id_embedding.weight)
        self.res_scale = nn.Parameter(torch.ones(1))

        def proj(in_dim):
            return nn.Sequential(nn.Linear(in_dim, d), nn.BatchNorm1d(d), nn.LeakyReLU(negative_slope=0.2), nn.Dropout(0.1))

        def gate():
            return nn.Sequential(nn.Linear(d, d), nn.BatchNorm1d(d), nn.Sigmoid())

        self.image_residual_project = proj(self.image_embedding.shape[1])
        self.image_modal_project = proj(d)
        self.text_residual_project = proj(self.text_embedding.shape[1])
        self.text_modal_project = proj(d)
        self.caculate_common = nn.Sequential(nn.Linear(d, d), nn.BatchNorm1d(d), nn.Tanh(), nn.Linear(d, 1, bias=False))
        self.gate_image_modal = gate()
        self.gate_text_modal = gate()
        for mod in self.modules():
            if isinstance(mod, nn.Linear):
                nn.init.xavier_uniform_(mod.weight)

        self.image_UI_matrix = None
        self.image_II_matrix = None
        self.text_II_matrix = None

    def __setattr__(self, name, value):
        if name in ("image_UI_matrix", "norm_adj") and "_graph_version" in self.__dict__:
            self.__dict__["_graph_version"] += 1
        super().__setattr__(name, value)

    # ---- graphs the reference's trainer owns ---------------------------------------------------------
    def set_generated_edges(self, image_edges):
        parts = gb.ui_matrix(image_edges[0], image_edges[1], self.n_users, self.n_items, device=self.device)
        self.image_UI_matrix = self.edgeDropper(parts, self.device)

    def build_item_item_matrices(self):
        build = gb.knn_graph_fused if self.knn_builder == "fused" else gb.knn_graph_dense
        self.image_II_matrix = GraphCSR.from_coo(*build(self.image_embedding, self.knn_k, eps_normalize=True), self.device)
        self.text_II_matrix = GraphCSR.from_coo(*build(self.text_embedding, self.knn_k, eps_normalize=True), self.device)

    # ---- propagation -----------------------------------------------------------------------------
    def getItemEmbeds(self):
        return self.item_id_embedding.weight

    def getUserEmbeds(self):
        return self.user_embedding.weight

    def _modal_feats(self, residual, modal, feat):
        x = residual(feat)
        return self.res_scale * x + modal(x)

    def getImageFeats(self):
        return self._modal_feats(self.image_residual_project, self.image_modal_project, self.image_embedding)

    def getTextFeats(self):
        return self._modal_feats(self.text_residual_project, self.text_modal_project, self.text_embedding)

    def user_item_GCN(self, adj):
        adj = as_graph(adj)
        e = torch.cat([self.user_embedding.weight, self.item_id_embedding.weight], dim=0)
        acc = e
        for _ in range(self.n_layers):
            e = spmm(adj, e)
            acc = acc + e
        return acc / float(self.n_layers + 1)

    def content_embedding(self, original_ui_adj, diffusion_ui_image_adj):
        c1 = self.user_item_GCN(original_ui_adj)
        c2 = self.user_item_GCN(diffusion_ui_image_adj)
        w = F.softmax(torch.stack([self.origin_weight, self.generation_weight]), dim=0)
        return w[0] * c1 + w[1] * c2

    def item_item_GCN(self, R, diffusion_ii_image_adj, diffusion_ii_text_adj):
        R, gi, gt = as_graph(R), as_graph(diffusion_ii_image_adj), as_graph(diffusion_ii_text_adj)
        item = self.item_id_embedding.weight

        def branch(g, feats, gate):
            e = torch.multiply(item, gate(feats))
            for _ in range(self.n_layers):
                e = spmm(g, e)
            return torch.cat([spmm(R, e), e], dim=0)

        return branch(gi, self.getImageFeats(), self.gate_image_modal), branch(gt, self.getTextFeats(), self.gate_text_modal)

    def forward(self, R, original_ui_adj, diffusion_ui_image_adj, diffusion_ii_image_adj, diffusion_ii_text_adj):
        content = self.content_embedding(original_ui_adj, diffusion_ui_image_adj)
        img, txt = self.item_item_GCN(R, diffusion_ii_image_adj, diffusion_ii_text_adj)
        att = F.softmax(torch.cat([self.caculate_common(img), self.caculate_common(txt)], dim=-1), dim=-1)
        common = att[:, 0].unsqueeze(dim=1) * img + att[:, 1].unsqueeze(dim=1) * txt
        s_img = torch.multiply(self.gate_image_modal(content), img - common)
        s_txt = torch.multiply(self.gate_text_modal(content), txt - common)
        side = (s_img + s_txt + common) / 4
        return content, side

    def propagate(self):
        if self.image_UI_matrix is None:
            raise RuntimeError("GenRecV1: image_UI_matrix is not set (the reference's trainer builds it every epoch)")
        c = self.content_embedding(self.norm_adj, self.image_UI_matrix)
        return c[:self.n_users], c[self.n_users:]

    @staticmethod
    def infoNCE_loss(view1, view2, temperature):
        """In-batch InfoNCE of genrecv1.py:408-415: -log( exp(<a_i, b_i> / T) / sum_j exp(<a_i, b_j> / T) ), rows normalised."""
        a, b = F.normalize(view1, dim=1), F.normalize(view2, dim=1)
        pos = torch.exp(torch.sum(a * b, dim=-1) / temperature)
        tot = torch.exp(a @ b.t() / temperature).sum(dim=1)
        return -torch.log(pos / tot).mean()

    def calculate_loss(self, interaction):
        """BPR + L2 + the item-item and user-item contrastive terms of genrecv1.py:355-401 (content vs side embedding);
        every graph product goes through the differentiable ``ops.spmm``, the BPR gathers through the fused kernel."""
        users, pos_items, neg_items = interaction[0], interaction[1], interaction[2]
        if self.image_UI_matrix is None:
            return torch.tensor(0.0, requires_grad=True).to(self.device)
        if self.image_II_matrix is None or self.text_II_matrix is None:
            self.build_item_item_matrices()       # the reference's trainer builds them once (common/trainer.py:676-687)
        nu = self.n_users
        content, side = self.forward(self.R, self.norm_adj, self.image_UI_matrix, self.image_II_matrix, self.text_II_matrix)
        ue, ie = content[:nu], content[nu:]
        su, si = side[:nu], side[nu:]
        ps, ns = bpr_scores(ue, ie, users, pos_items, neg_items)   # fused gather + dot (csrc/train_ops.cu)
        bpr = -torch.mean(F.logsigmoid(ps - ns))
        reg = (self.user_embedding.weight.norm(2).square() + self.item_id_embedding.weight.norm(2).square()) * self.reg_weight
        cl1 = self.infoNCE_loss(si[pos_items], ie[pos_items], self.temp) + self.infoNCE_loss(su[users], ue[users], self.temp)
        cl2 = self.infoNCE_loss(ue[users], ie[pos_items], self.temp) + self.infoNCE_loss(ue[users], si[pos_items], self.temp)
        return bpr + reg + cl1 * self.ssl_reg1 + cl2 * self.ssl_reg2
