"""Pieces shared by the model plugins: graph containers for the bipartite normalised adjacency and
the BPR / embedding losses (GenMMRec/src/common/loss.py:9-35)."""
import torch
import torch.nn.functional as F

from .. import graph as gb
from ..ops import GraphCSR, spmm


class BipartiteAdj:
    """The normalised user-item adjacency (N x N, N = n_users + n_items) held three ways: the full CSR
    (what ``torch.sparse.mm(norm_adj, X)`` multiplies) and its two off-diagonal blocks
    R_hat [U x I] / R_hat' [I x U], which let a propagation step read only the half of X a row block
    can reach and fuse several right-hand sides into one wide pass.  Row results are bit-identical
    between the two forms (same entries, same order within a row)."""

    def __init__(self, indices, values, n_users, n_items, device):
        n = n_users + n_items
        self.n_users, self.n_items = n_users, n_items
        self.shape = (n, n)
        self.full = GraphCSR.from_coo(indices, values, (n, n), device)
        halves = gb.bipartite_halves(indices, values, n_users, n_items)
        self.ui = self.iu = None
        if halves is not None:
            self.ui = GraphCSR.from_coo(*halves[0], device)
            self.iu = GraphCSR.from_coo(*halves[1], device)

    @property
    def nnz(self):
        return self.full.nnz

    def mm(self, x):
        """A @ x for x [N, D] (autograd-aware)."""
        return spmm(self.full, x)

    def to_torch_coo(self):
        return self.full.to_torch_coo()


def as_graph(g):
    """Accept a GraphCSR, a BipartiteAdj or a torch sparse tensor (the reference's graph objects)."""
    if isinstance(g, BipartiteAdj):
        return g.full
    if isinstance(g, GraphCSR):
        return g
    if torch.is_tensor(g) and g.is_sparse:
        return GraphCSR.from_torch_sparse(g)
    raise TypeError("unsupported graph type %r" % type(g))


def bpr_loss(pos_score, neg_score, gamma=1e-10):
    return -torch.log(gamma + torch.sigmoid(pos_score - neg_score)).mean()


def emb_loss(*embeddings):
    loss = torch.zeros(1, device=embeddings[-1].device)
    for e in embeddings:
        loss = loss + torch.norm(e, p=2)
    return loss / embeddings[-1].shape[0]


def normalize(x):
    return F.normalize(x)
