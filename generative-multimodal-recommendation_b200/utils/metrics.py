"""Top-K metric vectors from a hit matrix (host-side, numpy) -- same definitions as
GenMMRec/src/utils/metrics.py:12-105.  The product path computes these on the device
(csrc/metrics.cu); this module exists so that callers holding a numpy hit matrix (e.g. results
loaded from a CSV dump) get the same numbers, and to document the formulas."""
import numpy as np


def _cum(pos_index):
    return np.cumsum(pos_index, axis=1, dtype=np.float64)


def recall_(pos_index, pos_len):
    return (_cum(pos_index) / pos_len.reshape(-1, 1)).mean(axis=0)


def recall2_(pos_index, pos_len):
    return _cum(pos_index).sum(axis=0) / pos_len.sum()


def precision_(pos_index, pos_len):
    return (_cum(pos_index) / np.arange(1, pos_index.shape[1] + 1)).mean(axis=0)


def ndcg_(pos_index, pos_len):
    k = pos_index.shape[1]
    disc = 1.0 / np.log2(np.arange(1, k + 1, dtype=np.float64) + 1)
    lim = np.minimum(pos_len, k)
    idcg = np.cumsum(disc)[np.minimum(np.arange(k)[None, :], lim[:, None] - 1)]
    dcg = np.cumsum(np.where(pos_index, disc[None, :], 0.0), axis=1)
    return (dcg / idcg).mean(axis=0)


def map_(pos_index, pos_len):
    k = pos_index.shape[1]
    pre = _cum(pos_index) / np.arange(1, k + 1)
    sum_pre = np.cumsum(pre * pos_index.astype(np.float64), axis=1)
    denom = np.minimum(np.arange(1, k + 1)[None, :], np.minimum(pos_len, k)[:, None])
    return (sum_pre / denom).mean(axis=0)


metrics_dict = {"ndcg": ndcg_, "recall": recall_, "recall2": recall2_, "precision": precision_, "map": map_}
