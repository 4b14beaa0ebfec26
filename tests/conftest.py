"""pytest plumbing: `gpu` marker, repo root on sys.path, golden-fixture loader."""
import json
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


class Golden(dict):
    """arrays as torch tensors; `.meta` dict; `.sd` = state-dict sub-mapping ('sd/' keys)."""

    def __init__(self, name):
        super().__init__()
        with np.load(os.path.join(GOLDEN, name + ".npz")) as f:
            self.meta = json.loads(bytes(f["meta"]).decode())
            self.sd = {}
            for k in f.files:
                if k == "meta":
                    continue
                t = torch.from_numpy(np.array(f[k]))
                if k.startswith("sd/"):
                    self.sd[k[3:]] = t
                else:
                    self[k] = t


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = Golden(name)
        return cache[name]
    return load
