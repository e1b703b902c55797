"""GUME -- propagation + full-sort scoring of GenMMRec/src/models/gume.py on the B200 kernels.

Graphs (gume.py:52-80,122-201): two modality kNN graphs (k = knn_k, weighted symmetric normalisation),
the enhanced user-item graph ``norm_adj`` (train pairs + items whose image AND text neighbourhoods
agree) and its user-by-item block ``R``.  ``forward`` (gume.py:229-276) is 3 * n_ui_layers +
2 * n_layers + 2 SpMMs; here the three ``conv_ui`` chains that share ``norm_adj`` run as ONE chain
over a 192-wide right-hand side, and the two ``R`` products as one 128-wide pass.  The
``extended_*`` chains only feed training outputs (``train=True``) and are skipped at evaluation.

The reference caches its kNN graphs and the neighbourhood intersection as files inside the dataset
directory (gume.py:52-62,123-149); nothing is written here.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..common.abstract_recommender import GeneralRecommender
from .. import graph as gb
from ..ops import GraphCSR, bpr_scores, linear_proj, spmm, spmm_raw


class GUME(GeneralRecommender):
    def __init__(self, config, dataset):
        super(GUME, self).__init__(config, dataset)
        self.sparse = True
        self.reg_weight_1 = config["reg_weight_1"]
        self.n_ui_layers = config["n_ui_layers"]
        self.embedding_dim = config["embedding_size"]
        self.knn_k = config["knn_k"]
        self.n_layers = config["n_layers"]
        self.knn_builder = config["knn_builder"] or "fused"  # fused (K2, no I x I matrix) | dense
        # weights / temperatures of the auxiliary training terms (gume.py:33-41 of the reference)
        self.reg_weight_2 = config["reg_weight_2"]
        self.bm_loss, self.um_loss, self.vt_loss = config["bm_loss"], config["um_loss"], config["vt_loss"]
        self.bm_temp, self.um_temp = config["bm_temp"], config["um_temp"]

        self.interaction_matrix = dataset.inter_matrix(form="coo")
        d = self.embedding_dim
        self.user_embedding = nn.Embedding(self.n_users, d)
        self.item_id_embedding = nn.Embedding(self.n_items, d)
        self.extended_image_user = nn.Embedding(self.n_users, d)
        self.extended_text_user = nn.Embedding(self.n_users, d)
        for e in (self.user_embedding, self.item_id_embedding, self.extended_image_user, self.extended_text_user):
            nn.init.xavier_uniform_(e.weight)
        self.image_embedding = nn.Embedding.from_pretrained(self.v_feat, freeze=False)
        self.text_embedding = nn.Embedding.from_pretrained(self.t_feat, freeze=False)

        self.image_reduce_dim = nn.Linear(self.v_feat.shape[1], d)
        self.image_trans_dim = nn.Sequential(nn.Linear(d, d), nn.Sigmoid())
        self.image_space_trans = nn.Sequential(self.image_reduce_dim, self.image_trans_dim)
        self.text_reduce_dim = nn.Linear(self.t_feat.shape[1], d)
        self.text_trans_dim = nn.Sequential(nn.Linear(d, d), nn.Sigmoid())
        self.text_space_trans = nn.Sequential(self.text_reduce_dim, self.text_trans_dim)
        self.separate_coarse = nn.Sequential(nn.Linear(d, d), nn.Tanh(), nn.Linear(d, 1, bias=False))
        self.softmax = nn.Softmax(dim=-1)
        self.image_behavior = nn.Sequential(nn.Linear(d, d), nn.Sigmoid())
        self.text_behavior = nn.Sequential(nn.Linear(d, d), nn.Sigmoid())
        self.to(self.device)
        self.build_graphs()

    def build_graphs(self, image_knn=None, text_knn=None):
        """kNN graphs from the CURRENT feature tables (or given as COO parts), then the enhanced
        adjacency and R."""
        build = gb.knn_graph_fused if self.knn_builder == "fused" else gb.knn_graph_dense
        with torch.no_grad():
            img = image_knn if image_knn is not None else build(self.image_embedding.weight.detach(), self.knn_k)
            txt = text_knn if text_knn is not None else build(self.text_embedding.weight.detach(), self.knn_k)
        self.image_original_adj = GraphCSR.from_coo(*img, self.device)
        self.text_original_adj = GraphCSR.from_coo(*txt, self.device)
        k = self.knn_k
        m = self.interaction_matrix
        na, r = gb.gume_adj(m.row, m.col, self.n_users, self.n_items, img[0][1].reshape(-1, k), txt[0][1].reshape(-1, k),
                            device=self.device)
        self.norm_adj = GraphCSR.from_coo(*na, self.device)
        self.R = GraphCSR.from_coo(*r, self.device)
        self.invalidate_cache()

    def conv_ui(self, adj, user_embeds, item_embeds):
        ego = torch.cat([user_embeds, item_embeds], dim=0)
        acc = ego
        for _ in range(self.n_ui_layers):
            ego = spmm(adj, ego)
            acc = acc + ego
        return acc / float(self.n_ui_layers + 1)

    def conv_ii(self, ii_adj, single_modal):
        for _ in range(self.n_layers):
            single_modal = spmm(ii_adj, single_modal)
        return single_modal

    def forward(self, adj, train=False):
        item_embeds = self.item_id_embedding.weight
        user_embeds = self.user_embedding.weight
        # [I, 4096] / [I, 384] feature tables x Linear(., 64): the tcgen05 projection kernel under no_grad
        image_item_embeds = torch.multiply(item_embeds, self.image_trans_dim(
            linear_proj(self.image_embedding.weight, self.image_reduce_dim)))
        text_item_embeds = torch.multiply(item_embeds, self.text_trans_dim(
            linear_proj(self.text_embedding.weight, self.text_reduce_dim)))

        explicit_image_item = self.conv_ii(self.image_original_adj, image_item_embeds)
        explicit_text_item = self.conv_ii(self.text_original_adj, text_item_embeds)
        if torch.is_grad_enabled():
            explicit_image_user = spmm(self.R, explicit_image_item)
            explicit_text_user = spmm(self.R, explicit_text_item)
        else:  # both R products in one 128-wide pass
            both = spmm_raw(self.R, torch.cat([explicit_image_item, explicit_text_item], dim=1))
            d = self.embedding_dim
            explicit_image_user, explicit_text_user = both[:, :d], both[:, d:]
        explicit_image_embeds = torch.cat([explicit_image_user, explicit_image_item], dim=0)
        explicit_text_embeds = torch.cat([explicit_text_user, explicit_text_item], dim=0)

        if train:
            extended_id_embeds = self.conv_ui(adj, user_embeds, item_embeds)
            extended_image_embeds = self.conv_ui(adj, self.extended_image_user.weight, explicit_image_item)
            extended_text_embeds = self.conv_ui(adj, self.extended_text_user.weight, explicit_text_item)
            extended_it_embeds = (extended_image_embeds + extended_text_embeds) / 2
        else:
            extended_id_embeds = self.conv_ui(adj, user_embeds, item_embeds)

        weights = self.softmax(torch.cat([self.separate_coarse(explicit_image_embeds),
                                          self.separate_coarse(explicit_text_embeds)], dim=-1))
        image_weights, text_weights = torch.split(weights, 1, dim=-1)
        coarse = image_weights * explicit_image_embeds + text_weights * explicit_text_embeds
        fine_image = torch.multiply(self.image_behavior(extended_id_embeds), (explicit_image_embeds - coarse))
        fine_text = torch.multiply(self.text_behavior(extended_id_embeds), (explicit_text_embeds - coarse))
        integration = (fine_image + fine_text + coarse) / 3
        all_embeds = extended_id_embeds + integration
        if train:
            return all_embeds, (integration, extended_id_embeds, extended_it_embeds), (explicit_image_embeds, explicit_text_embeds)
        return all_embeds

    def propagate(self):
        e = self.forward(self.norm_adj)
        return e[:self.n_users], e[self.n_users:]

    # ---- training objective (GenMMRec/src/models/gume.py:278-412), differentiable through ops.spmm ------------------
    @staticmethod
    def InfoNCE(view1, view2, temperature, chunk_size=4096):
        """-log( exp(<a_i, b_i> / T) / sum_j exp(<a_i, b_j> / T) + 1e-8 ), mean over i, rows L2-normalised; the
        denominator is accumulated over column chunks so that the [N, N] similarity matrix is never held whole."""
        a, b = F.normalize(view1, dim=1), F.normalize(view2, dim=1)
        pos = torch.exp((a * b).sum(dim=-1) / temperature)
        total = torch.zeros_like(pos)
        for j in range(0, b.shape[0], chunk_size):
            total = total + torch.exp(a @ b[j:j + chunk_size].t() / temperature).sum(dim=1)
        return (-torch.log(pos / total + 1e-8)).mean()

    def cal_noise_loss(self, idx, emb, temp):
        """InfoNCE between two randomly perturbed views of `emb` (sign-preserving noise of norm 0.1), rows `idx`."""
        def perturbed(x):
            return x + torch.sign(x) * F.normalize(torch.rand_like(x), dim=-1) * 0.1

        return self.InfoNCE(perturbed(emb)[idx], perturbed(emb)[idx], temp)

    @staticmethod
    def align_vt(e1, e2):
        """|var - var| + |mean - mean| of the two explicit modality embeddings."""
        return (torch.abs(torch.var(e1) - torch.var(e2)) + torch.abs(torch.mean(e1) - torch.mean(e2))).mean()

    def calculate_loss(self, interaction):
        """BPR + L2 + behaviour/modality alignment (InfoNCE) + visual/text alignment + user-modality enhancement: the
        full objective of gume.py:357-395."""
        users, pos_items, neg_items = interaction[0], interaction[1], interaction[2]
        nu, ni = self.n_users, self.n_items
        e, (integration, ext_id, ext_it), (exp_img, exp_txt) = self.forward(self.norm_adj, train=True)
        ue, ie = e[:nu], e[nu:]
        u, p, n = ue[users], ie[pos_items], ie[neg_items]
        ps, ns = bpr_scores(ue, ie, users, pos_items, neg_items)   # fused gather + dot (csrc/train_ops.cu)
        bpr = -torch.mean(F.logsigmoid(ps - ns))
        reg1 = self.reg_weight_1 * 0.5 * ((u ** 2).sum() + (p ** 2).sum() + (n ** 2).sum()) / self.batch_size
        int_u, int_i = integration[:nu], integration[nu:]
        id_u, id_i = ext_id[:nu], ext_id[nu:]
        it_u, it_i = ext_it[:nu], ext_it[nu:]
        vt = self.vt_loss * self.align_vt(exp_img, exp_txt)
        bm = self.bm_loss * (self.InfoNCE(int_u[users], id_u[users], self.bm_temp)
                             + self.InfoNCE(int_i[pos_items], id_i[pos_items], self.bm_temp))
        um = self.um_loss * (self.InfoNCE(it_u[users], int_u[users], self.um_temp)
                             + self.cal_noise_loss(users, int_u, self.um_temp)
                             + self.cal_noise_loss(users, it_u, self.um_temp))
        reg2 = self.reg_weight_2 * 0.5 * (it_i[pos_items] ** 2).sum() / self.batch_size
        return bpr + (vt + bm) + um + (reg1 + reg2)
