"""Interaction tables: the reference's ``RecDataset`` contract over flat tensors.

GenMMRec/src/utils/dataset.py:21-141 keeps a pandas DataFrame; the hot path only needs the
(user, item) columns, the split label and the table sizes, so this class stores int64 tensors -- on
the host for file-backed data, on the GPU for the synthetic 1M-user shape, where every loader
structure is then built with device sorts instead of per-user Python loops.  ``RecDataset(config)``
reads the same ``<data_path>/<dataset>/<inter_file_name>`` TSV; ``RecDataset.from_arrays`` skips
the file.  A DataFrame view (``.df``) is built on demand.
"""
import os

import numpy as np
import torch


def _as_i64(a, device=None):
    if a is None:
        return None
    t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
    return t.to(device=device if device is not None else t.device, dtype=torch.int64)


class RecDataset(object):
    def __init__(self, config, df=None, arrays=None):
        self.config = config
        self.dataset_name = config["dataset"]
        self.uid_field = config["USER_ID_FIELD"]
        self.iid_field = config["ITEM_ID_FIELD"]
        self.splitting_label = config["inter_splitting_label"]
        self._df = None
        if arrays is not None:
            self.users, self.items, self.labels = (_as_i64(a) for a in arrays)
        elif df is not None:
            self._set_from_df(df)
        else:
            import pandas as pd

            path = os.path.join(os.path.abspath(config["data_path"] + self.dataset_name), config["inter_file_name"])
            if not os.path.isfile(path):
                raise ValueError("File {} not exist".format(path))
            cols = [self.uid_field, self.iid_field, self.splitting_label]
            df = pd.read_csv(path, usecols=cols, sep=config["field_separator"])
            missing = [c for c in cols if c not in df.columns]
            if missing:
                raise ValueError("File {} lost some required columns: {}.".format(path, ", ".join(missing)))
            self._set_from_df(df)
        self.item_num = int(self.items.max()) + 1 if self.items.numel() else 0
        self.user_num = int(self.users.max()) + 1 if self.users.numel() else 0
        self.inter_num = len(self)

    def _set_from_df(self, df):
        self.users = _as_i64(df[self.uid_field].values)
        self.items = _as_i64(df[self.iid_field].values)
        self.labels = _as_i64(df[self.splitting_label].values) if self.splitting_label in df.columns else None

    @classmethod
    def from_arrays(cls, config, users, items, labels, n_users=None, n_items=None):
        ds = cls(config, arrays=(users, items, labels))
        if n_users is not None:
            ds.user_num = int(n_users)
        if n_items is not None:
            ds.item_num = int(n_items)
        return ds

    @property
    def df(self):
        if self._df is None:
            import pandas as pd

            cols = {self.uid_field: self.users.cpu().numpy(), self.iid_field: self.items.cpu().numpy()}
            if self.labels is not None:
                cols[self.splitting_label] = self.labels.cpu().numpy()
            self._df = pd.DataFrame(cols)
        return self._df

    def split(self):
        """Train / valid / test by ``x_label`` 0/1/2; valid/test rows of users without train rows are
        dropped when ``filter_out_cod_start_users`` (dataset.py:65-82)."""
        parts = []
        for lab in range(3):
            m = self.labels == lab
            parts.append((self.users[m], self.items[m]))
        if self.config["filter_out_cod_start_users"]:
            seen = torch.zeros(self.user_num, dtype=torch.bool, device=self.users.device)
            seen[parts[0][0]] = True
            parts = [parts[0]] + [(u[seen[u]], i[seen[u]]) for u, i in parts[1:]]
        return [self.copy_arrays(u, i) for u, i in parts]

    def copy_arrays(self, users, items):
        nxt = RecDataset(self.config, arrays=(users, items, None))
        nxt.item_num, nxt.user_num = self.item_num, self.user_num
        return nxt

    def copy(self, new_df):
        nxt = RecDataset(self.config, df=new_df)
        nxt.item_num, nxt.user_num = self.item_num, self.user_num
        return nxt

    def get_user_num(self):
        return self.user_num

    def get_item_num(self):
        return self.item_num

    def __len__(self):
        return int(self.users.numel())

    def __str__(self):
        n = len(self)
        nu, ni = int(torch.unique(self.users).numel()), int(torch.unique(self.items).numel())
        info = [self.dataset_name, "The number of users: {}".format(nu), "Average actions of users: {}".format(n / max(nu, 1)),
                "The number of items: {}".format(ni), "Average actions of items: {}".format(n / max(ni, 1)),
                "The number of inters: {}".format(n),
                "The sparsity of the dataset: {}%".format((1 - n / max(nu, 1) / max(ni, 1)) * 100)]
        return "\n".join(info)

    __repr__ = __str__
