import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_sd():
    d = np.load(os.path.join(GOLDEN, "lru_weights_n400.npz"))
    return {k: torch.from_numpy(d[k]) for k in d.files}


CASES = ["left_l20", "left_l50", "left_l200", "holes_l50", "holes_l37"]


def load_case(name):
    d = np.load(os.path.join(GOLDEN, f"lru_case_{name}.npz"))
    return {k: d[k] for k in d.files}


def metrics_vector(m, ks):
    return np.array([m[f"{n}@{k}"] for k in ks for n in ("Recall", "MRR", "NDCG")])
