import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(bytes(z["meta"]).decode()) if "meta" in z.files else {}
    return z, meta


def golden_params(z):
    return {k[len("param/"):]: z[k] for k in z.files if k.startswith("param/")}


def toy_arrays():
    """The toy interactions/features every toy_*.npz fixture was generated from
    (tests/golden/make_golden.py: TOY)."""
    from genmmrec_b200 import synth

    users, items, label = synth.make_interactions(300, 120, 3600)
    img, txt = synth.make_features(120, image_dim=64, text_dim=32)
    return dict(users=users, items=items, label=label, img=img, txt=txt, n_users=300, n_items=120)


@pytest.fixture(scope="session")
def toy_data():
    return toy_arrays()
