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


@pytest.fixture
def allow_library():
    """Tests on nets narrower than the kernels take (hidden width 8 / 16: attention head dim < 8, Linear K < 32) opt in
    to the library layers explicitly; everything else must run on flowk kernels only (`no_library`)."""
    from flowk import _lib
    old = _lib.ALLOW_LIBRARY
    _lib.ALLOW_LIBRARY = True
    yield _lib
    _lib.ALLOW_LIBRARY = old


@pytest.fixture
def no_library():
    """Asserts that the test body never routed a CUDA tensor to cuDNN / cuBLAS / ATen conditioner layers."""
    from flowk import _lib
    old, before = _lib.ALLOW_LIBRARY, _lib.LIBRARY_FALLBACKS
    _lib.ALLOW_LIBRARY = False
    yield _lib
    _lib.ALLOW_LIBRARY = old
    assert _lib.LIBRARY_FALLBACKS == before


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = Golden(name)
        return cache[name]
    return load
