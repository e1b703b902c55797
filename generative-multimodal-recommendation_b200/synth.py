"""Seeded synthetic datasets in the on-disk format the reference reads.

The reference ships no data (GenMMRec/data/README.md:2 points at a download), so every test and
bench runs on synthetic interactions of the published shapes (GenMMRec/evaluation/README.md:8-9,
BASELINE.json configs).  The generator follows SURVEY.md §8(d):

* unique (user, item) pairs, user degree >= 5 (5-core) and heavy-tailed (log-normal, clipped),
  item choice proportional to rank^-0.8 under a random rank->id permutation;
* per-user split rule of GenMMRec/preprocessing/1splitting.ipynb: n < 10 -> the last two
  interactions are valid/test, otherwise 10 % + 10 %; ``split='loo'`` keeps exactly one valid and
  one test row per user (the Scaled shape: the quoted interaction count is the train set);
* files: ``<name>.inter`` TSV with columns userID itemID rating timestamp x_label
  (GenMMRec/src/configs/dataset/baby.yaml:2-9, overall.yaml:8), ``image_feat.npy`` fp32
  ``relu(N(0,1))`` and ``text_feat.npy`` fp32 row-normalised ``N(0,1)``.

Pure numpy (``numpy.random.default_rng`` is bit-reproducible across machines), no CUDA needed.
"""
import os

import numpy as np

SHAPES = {
    # name: (n_users, n_items, n_interactions, split)
    "toy": (300, 120, 3600, "ratio"),
    "baby": (19445, 7050, 160792, "ratio"),
    "sports": (35598, 18357, 296337, "ratio"),
    "clothing": (39387, 23033, 278677, "ratio"),
    "scaled": (1_000_000, 500_000, 50_000_000, "loo"),
}

FEAT_DIMS = {"image": 4096, "text": 384}


def user_degrees(n_users, total, rng, min_deg=5, max_deg=None, sigma=1.0):
    """Heavy-tailed degrees >= min_deg that sum exactly to ``total``."""
    assert total >= min_deg * n_users, "need at least min_deg interactions per user"
    w = rng.lognormal(mean=0.0, sigma=sigma, size=n_users)
    extra = total - min_deg * n_users
    deg = min_deg + np.floor(w / w.sum() * extra).astype(np.int64)
    if max_deg is not None:
        deg = np.minimum(deg, max_deg)
    short = int(total - deg.sum())
    while short > 0:  # hand the remainder out one by one to random users below the cap
        cand = np.flatnonzero(deg < max_deg) if max_deg is not None else np.arange(n_users)
        take = min(short, cand.size)
        deg[rng.choice(cand, size=take, replace=False)] += 1
        short -= take
    return deg


def item_cdf(n_items, rng, alpha=0.8):
    """Sampling CDF over item ids: popularity ~ rank^-alpha, ranks randomly assigned to ids."""
    p = np.arange(1, n_items + 1, dtype=np.float64) ** (-alpha)
    perm = rng.permutation(n_items)  # perm[rank] = item id
    cdf = np.cumsum(p / p.sum())
    cdf[-1] = 1.0
    return cdf, perm


def make_interactions(n_users, n_items, n_inter, seed=999, split="ratio", alpha=0.8):
    """Return (user, item, x_label) int64 arrays, grouped by user, in per-user time order.

    ``x_label``: 0 train, 1 valid, 2 test.  With ``split='loo'`` ``n_inter`` counts the train rows
    only and one valid + one test row per user are added on top.
    """
    rng = np.random.default_rng(seed)
    extra_per_user = 2 if split == "loo" else 0
    total = n_inter + extra_per_user * n_users
    deg = user_degrees(n_users, total, rng, min_deg=5 + extra_per_user,
                       max_deg=max(8, n_items // 4))
    cdf, perm = item_cdf(n_items, rng, alpha)
    users = np.repeat(np.arange(n_users, dtype=np.int64), deg)
    items = perm[np.searchsorted(cdf, rng.random(users.size), side="right").clip(0, n_items - 1)]
    # de-duplicate (u, i): redraw the repeated positions until none is left
    for it in range(64):
        key = users * n_items + items
        order = np.argsort(key, kind="stable")
        sk = key[order]
        dup = np.zeros(users.size, dtype=bool)
        dup[order[1:]] = sk[1:] == sk[:-1]
        n_dup = int(dup.sum())
        if n_dup == 0:
            break
        if it < 48:
            items[dup] = perm[np.searchsorted(cdf, rng.random(n_dup), side="right").clip(0, n_items - 1)]
        else:  # popular items exhausted for some user: fall back to uniform draws
            items[dup] = rng.integers(0, n_items, size=n_dup)
    else:
        raise RuntimeError("could not de-duplicate synthetic interactions")
    # make sure the largest ids exist (the reference sizes its tables by max id + 1,
    # GenMMRec/src/utils/dataset.py:50-51)
    if not np.any(items == n_items - 1):
        j = int(np.flatnonzero(users == 0)[0])
        items[j] = n_items - 1
    # per-user random time order: shuffle inside each user's segment
    tkey = rng.random(users.size)
    order = np.lexsort((tkey, users))
    users, items = users[order], items[order]
    start = np.concatenate([[0], np.cumsum(deg)[:-1]])
    pos = np.arange(users.size) - np.repeat(start, deg)  # position in the user's sequence
    n_u = np.repeat(deg, deg)
    if split == "loo":
        n_eval = np.ones_like(n_u)
    else:
        n_eval = np.where(n_u < 10, 1, np.maximum(1, (n_u // 10)))
    label = np.zeros(users.size, dtype=np.int64)
    label[pos >= n_u - 2 * n_eval] = 1
    label[pos >= n_u - n_eval] = 2
    return users, items, label


def make_features(n_items, seed=999, image_dim=FEAT_DIMS["image"], text_dim=FEAT_DIMS["text"]):
    rng = np.random.default_rng(seed + 1)
    img = np.maximum(rng.standard_normal((n_items, image_dim), dtype=np.float32), 0.0)
    txt = rng.standard_normal((n_items, text_dim), dtype=np.float32)
    txt /= np.linalg.norm(txt, axis=1, keepdims=True)
    return img, txt.astype(np.float32)


def write_dataset(root, name, n_users, n_items, n_inter, seed=999, split="ratio",
                  image_dim=FEAT_DIMS["image"], text_dim=FEAT_DIMS["text"]):
    """Write ``<root>/<name>/{<name>.inter, image_feat.npy, text_feat.npy}``; return the arrays."""
    import pandas as pd

    d = os.path.join(root, name)
    os.makedirs(d, exist_ok=True)
    users, items, label = make_interactions(n_users, n_items, n_inter, seed, split)
    ts = np.arange(users.size, dtype=np.int64)
    df = pd.DataFrame({"userID": users, "itemID": items, "rating": np.full(users.size, 5.0),
                       "timestamp": ts, "x_label": label})
    df.to_csv(os.path.join(d, name + ".inter"), sep="\t", index=False)
    img, txt = make_features(n_items, seed, image_dim, text_dim)
    np.save(os.path.join(d, "image_feat.npy"), img)
    np.save(os.path.join(d, "text_feat.npy"), txt)
    return users, items, label, img, txt


def make_params(shapes, seed=999):
    """Xavier-uniform-like parameters from per-name numpy streams, so the reference model and this
    package can be loaded with identical values on any machine regardless of which other names
    are requested.  ``shapes``: dict name -> shape tuple."""
    import zlib

    out = {}
    for k in sorted(shapes):
        shp = tuple(int(x) for x in shapes[k])
        rng = np.random.default_rng([seed, zlib.crc32(k.encode())])
        if len(shp) >= 2:
            bound = np.sqrt(6.0 / (shp[0] + shp[1]))
        else:
            bound = 0.5
        out[k] = rng.uniform(-bound, bound, size=shp).astype(np.float32)
    return out


def generated_edges(n_users, n_items, rebuild_k, seed=999):
    """Stand-in for the diffusion-generated edges the reference's trainers extract
    (GenMMRec/src/common/trainer.py:540-562): ``rebuild_k`` distinct uniform items per user."""
    rng = np.random.default_rng(seed + 3)
    u = np.repeat(np.arange(n_users, dtype=np.int64), rebuild_k)
    if rebuild_k == 1:
        i = rng.integers(0, n_items, size=n_users)
    else:
        # distinct per user: random keys, take the k smallest of a candidate window
        i = np.empty((n_users, rebuild_k), dtype=np.int64)
        base = rng.integers(0, n_items, size=n_users)
        step = rng.integers(1, max(2, n_items // rebuild_k), size=n_users)
        for j in range(rebuild_k):
            i[:, j] = (base + j * step) % n_items
        i = i.reshape(-1)
    return u, i.astype(np.int64)
