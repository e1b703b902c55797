"""Registry and small helpers (same entry points as GenMMRec/src/utils/utils.py:28-67,147-197)."""
import importlib
import random

import numpy as np
import torch

from .. import graph as _graph


def get_model(model_name):
    """models/<lower-case name>.py must define a class called `model_name` (utils.py:28-41)."""
    module = importlib.import_module("genmmrec_b200.models." + model_name.lower())
    return getattr(module, model_name)


def get_trainer(model_name=None):
    """Every model on the hot path is evaluated by the same Trainer.evaluate (utils.py:44-58 maps
    DiffMM / GenRecV1 to trainers that only add diffusion training, which is out of scope)."""
    from ..common.trainer import Trainer
    return Trainer


def init_seed(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)


def dict2str(result_dict):
    return "".join(str(m) + ": " + "%.04f" % v + "    " for m, v in result_dict.items())


def build_sim(context):
    """Cosine similarity matrix (utils.py:147-150).  Dense: small item counts only."""
    context_norm = context.div(torch.norm(context, p=2, dim=-1, keepdim=True))
    return torch.mm(context_norm, context_norm.transpose(1, 0))


def build_knn_normalized_graph(adj, topk, is_sparse=True, norm_type="sym"):
    """kNN graph of a dense similarity matrix as a sparse COO tensor (utils.py:184-197), built
    without the reference's Python list comprehension over edges."""
    if not is_sparse or norm_type != "sym":
        raise NotImplementedError("only the sparse 'sym' variant is on the hot path")
    val, ind = torch.topk(adj, topk, dim=-1)
    idx, w, shape = _graph.knn_from_topk(val, ind)
    return torch.sparse_coo_tensor(idx, w, shape)


def set_popularity_groups(config, train_dataset, pop_fraction=0.2, cold_start_threshold=5):
    """Fill ``config['pop_items']`` / ``config['warm_users']`` the way the reference's entry point does before it builds
    the trainer (GenMMRec/src/utils/quick_start.py:46-92): popular items = the first ``int(0.2 * n_unique)`` item ids of
    the train split ordered by interaction count (``value_counts``), warm users = users with more than five train
    interactions.  The evaluator reads both for the ``is_test`` group metrics.  Uses the same pandas call as the
    reference so that ties at the 20 % boundary fall the same way."""
    import pandas as pd

    items = train_dataset.items.cpu().numpy()
    users = train_dataset.users.cpu().numpy()
    unique_items = pd.Series(items).value_counts().index.tolist()
    config["pop_items"] = set(unique_items[:int(len(unique_items) * pop_fraction)])
    user_counts = pd.Series(users).value_counts()
    config["warm_users"] = set(user_counts[user_counts > cold_start_threshold].index.tolist())
    return config
