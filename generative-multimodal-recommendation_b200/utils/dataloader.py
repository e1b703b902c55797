"""Data loaders of the evaluation path, vectorised.

``EvalDataLoader`` reproduces the layout of GenMMRec/src/utils/dataloader.py:330-416 -- eval users in
first-appearance order, the train-history mask as COO ``(position in eval_u, item)``, batches
``[users, mask]`` with mask rows rebased to the batch, ragged ground truth -- but builds it with
sorts and prefix sums (on whatever device the dataset tensors live) instead of per-user
``groupby.get_group`` loops (3.1 s at the Baby shape in the reference, minutes at 1M users), and
additionally exposes the CSR forms the fused kernels consume.

``TrainDataLoader`` provides what model constructors call (``.dataset``, ``inter_matrix``); BPR
negative sampling is outside the hot path and is kept minimal.
"""
import math

import numpy as np
import torch
from scipy.sparse import coo_matrix


class InterCOO(object):
    """The (row, col) view of the train interactions that model constructors read
    (``dataset.inter_matrix(form='coo')``), without the scipy round trip when the data is already on
    the GPU.  ``row`` / ``col`` are int64 tensors; ``astype`` is accepted and ignored (values are 1.0)."""

    def __init__(self, row, col, shape):
        self.row, self.col, self.shape = row, col, shape
        self.nnz = int(row.numel())

    def astype(self, _dtype):
        return self

    def to_scipy(self):
        return coo_matrix((np.ones(self.nnz), (self.row.cpu().numpy(), self.col.cpu().numpy())), shape=self.shape)


class AbstractDataLoader(object):
    def __init__(self, config, dataset, additional_dataset=None, batch_size=1, neg_sampling=False, shuffle=False):
        self.config = config
        self.dataset = dataset
        self.additional_dataset = additional_dataset
        self.batch_size = batch_size
        self.step = batch_size
        self.shuffle = shuffle
        self.neg_sampling = neg_sampling
        self.device = config["device"]
        self.sparsity = 1 - len(dataset) / max(dataset.user_num, 1) / max(dataset.item_num, 1)
        self.pr = 0
        self.inter_pr = 0

    def pretrain_setup(self):
        pass

    def __len__(self):
        return math.ceil(self.pr_end / self.step)

    def __iter__(self):
        if self.shuffle:
            self._shuffle()
        return self

    def __next__(self):
        if self.pr >= self.pr_end:
            self.pr = 0
            self.inter_pr = 0
            raise StopIteration()
        return self._next_batch_data()


class TrainDataLoader(AbstractDataLoader):
    def __init__(self, config, dataset, batch_size=1, shuffle=False):
        super().__init__(config, dataset, additional_dataset=None, batch_size=batch_size, neg_sampling=True,
                         shuffle=shuffle)
        self._perm = None
        seed = config["seed"] if isinstance(config["seed"], int) else 999
        self._gen = torch.Generator(device=dataset.users.device)
        self._gen.manual_seed(seed)
        self._hist_keys = None

    def inter_matrix(self, form="coo", value_field=None):
        """Train interactions, data = 1.0 (dataloader.py:155-210).  'coo' returns a light (row, col)
        view usable on any device; 'csr' / 'scipy' materialise a scipy matrix on the host."""
        ds = self.dataset
        coo = InterCOO(ds.users, ds.items, (ds.user_num, ds.item_num))
        if form == "coo":
            return coo
        if form == "scipy":
            return coo.to_scipy()
        if form == "csr":
            return coo.to_scipy().tocsr()
        raise NotImplementedError("sparse matrix format [{}] has not been implemented.".format(form))

    @property
    def pr_end(self):
        return len(self.dataset)

    def _shuffle(self):
        self._perm = torch.randperm(len(self.dataset), generator=self._gen, device=self.dataset.users.device)

    def _next_batch_data(self):
        """[users; pos items; neg items] int64 [3, B]: one uniform negative per row, rejected against
        the user's train history (vectorised form of dataloader.py:226-275)."""
        ds = self.dataset
        dev = ds.users.device
        idx = (self._perm[self.pr:self.pr + self.step] if self._perm is not None
               else torch.arange(self.pr, min(self.pr + self.step, len(ds)), device=dev))
        self.pr += self.step
        u, i = ds.users[idx], ds.items[idx]
        if self._hist_keys is None:
            self._hist_keys = torch.unique(ds.users * ds.item_num + ds.items)
        hk = self._hist_keys
        neg = torch.randint(0, ds.item_num, (u.numel(),), generator=self._gen, device=dev)
        for _ in range(16):
            key = u * ds.item_num + neg
            pos = torch.searchsorted(hk, key).clamp_(max=hk.numel() - 1)
            bad = hk[pos] == key
            n_bad = int(bad.sum())
            if n_bad == 0:
                break
            neg[bad] = torch.randint(0, ds.item_num, (n_bad,), generator=self._gen, device=dev)
        return torch.stack([u, i, neg]).to(self.device)


def _group(users, items, n_users, order_users):
    """Rows of (users, items) regrouped in `order_users` order, original row order kept inside a user.
    Returns (lengths per listed user, flat items)."""
    order = torch.sort(users, stable=True).indices
    si = items[order]
    counts = torch.bincount(users, minlength=n_users)
    starts = torch.cumsum(counts, 0) - counts
    lens = counts[order_users]
    ends = torch.cumsum(lens, 0)
    total = int(ends[-1]) if lens.numel() else 0
    offs = torch.arange(total, device=users.device) - torch.repeat_interleave(ends - lens, lens)
    flat = si[torch.repeat_interleave(starts[order_users], lens) + offs]
    return lens, flat


class EvalDataLoader(AbstractDataLoader):
    """additional_dataset: the training split (its interactions are masked out at evaluation)."""

    def __init__(self, config, dataset, additional_dataset=None, batch_size=1, shuffle=False):
        super().__init__(config, dataset, additional_dataset=additional_dataset, batch_size=batch_size,
                         neg_sampling=False, shuffle=shuffle)
        if additional_dataset is None:
            raise ValueError("Training datasets is nan")
        n_users, n_items = dataset.user_num, dataset.item_num
        tr = additional_dataset
        work = dataset.users.device
        # eval users in first-appearance order (dataloader.py:345: df[uid].unique())
        uniq, inv = torch.unique(dataset.users, return_inverse=True)
        first = torch.full((uniq.numel(),), dataset.users.numel(), dtype=torch.int64, device=work)
        first.scatter_reduce_(0, inv, torch.arange(dataset.users.numel(), device=work), reduce="amin")
        eval_u = uniq[torch.argsort(first)]
        train_count = torch.bincount(tr.users, minlength=n_users)
        if bool((train_count[eval_u] == 0).any()):
            raise KeyError("an evaluated user has no training interaction (the reference's get_group raises too)")
        train_lens, train_flat = _group(tr.users.to(work), tr.items.to(work), n_users, eval_u)
        eval_lens, eval_flat = _group(dataset.users, dataset.items, n_users, eval_u)

        dev = self.device
        self._train_lens = train_lens
        self._train_pos_len_list = None
        self.eval_len_list = eval_lens.cpu().numpy()
        self._eval_flat = eval_flat
        self._eval_items_per_u = None
        n_eval = eval_u.numel()
        rows = torch.repeat_interleave(torch.arange(n_eval, device=work), train_lens)
        self.pos_items_per_u = torch.stack([rows, train_flat]).to(dev)
        self.eval_u = eval_u.to(dev)
        zero = torch.zeros(1, dtype=torch.int64, device=work)
        train_ptr = torch.cat([zero, torch.cumsum(train_lens, 0)])
        eval_ptr = torch.cat([zero, torch.cumsum(eval_lens, 0)])
        self._train_ptr_host = train_ptr.cpu().numpy()

        # CSR forms for the fused kernels: items ascending inside each row
        self.mask_items = (torch.sort(rows * n_items + train_flat).values % n_items).to(torch.int32).to(dev)
        self.mask_rowptr = train_ptr.to(dev)
        grows = torch.repeat_interleave(torch.arange(n_eval, device=work), eval_lens)
        self.gt_items = (torch.sort(grows * n_items + eval_flat).values % n_items).to(torch.int32).to(dev)
        self.gt_rowptr = eval_ptr.to(dev)

    @property
    def train_pos_len_list(self):
        if self._train_pos_len_list is None:
            self._train_pos_len_list = self._train_lens.cpu().tolist()
        return self._train_pos_len_list

    @property
    def pr_end(self):
        return self.eval_u.shape[0]

    def _shuffle(self):
        pass

    def _next_batch_data(self):
        lo, hi = int(self._train_ptr_host[self.pr]), int(self._train_ptr_host[min(self.pr + self.step, self.pr_end)])
        batch_users = self.eval_u[self.pr:self.pr + self.step]
        batch_mask_matrix = self.pos_items_per_u[:, lo:hi].clone()
        batch_mask_matrix[0] -= self.pr
        self.inter_pr = hi
        self.pr += self.step
        return [batch_users, batch_mask_matrix]

    def batch_mask_csr(self, pr, step):
        """(rowptr int64 [b+1], items int32) of eval-user positions [pr, pr+step)."""
        end = min(pr + step, self.pr_end)
        rp = self.mask_rowptr[pr:end + 1]
        return (rp - rp[0]).contiguous(), self.mask_items[int(rp[0]):int(rp[-1])]

    def get_eval_items(self):
        if self._eval_items_per_u is None:
            ptr = np.concatenate([[0], np.cumsum(self.eval_len_list)])
            self._eval_items_per_u = np.split(self._eval_flat.cpu().numpy(), ptr[1:-1])
        return self._eval_items_per_u

    def get_eval_len_list(self):
        return self.eval_len_list

    def get_eval_users(self):
        return self.eval_u.cpu()
