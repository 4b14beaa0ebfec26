"""CPU-side checks of the drop-in boundary: libflowk.so loads, exports every symbol include/flowk.h
declares (and nothing is bound that the header does not declare), argument validation answers without
touching a GPU, and the product refuses to run anywhere but on CUDA (no fallback)."""
import ctypes

import pytest
import torch

import flowk  # noqa: F401
from flowk import _lib


def test_header_symbols_exported():
    declared = _lib.declared_symbols()
    assert len(declared) >= 17
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), "libflowk.so does not export %s" % name
    assert set(_lib.SIGNATURES) == set(declared)


def test_abi_version_and_strings():
    assert _lib.lib.flowk_abi_version() == 1
    assert _lib.lib.flowk_error_string(0) == b"ok"
    assert _lib.lib.flowk_error_string(1) == b"bad shape"
    assert _lib.lib.flowk_ldj_workspace_bytes(64) == 64 * 65 * 4


def test_argument_validation_needs_no_gpu():
    L = _lib.lib
    one = ctypes.c_void_p(16)       # never dereferenced: validation fails first
    assert L.flowk_squeeze2d(None, one, 1, 1, 2, 2, 2, None) == _lib.FLOWK_ERR_ARG
    assert L.flowk_squeeze2d(one, one, 1, 3, 7, 8, 2, None) == _lib.FLOWK_ERR_SHAPE
    assert L.flowk_unsqueeze2d(one, one, 1, 6, 2, 2, 2, None) == _lib.FLOWK_ERR_SHAPE
    assert L.flowk_mixlogcdf_fwd(one, one, one, one, None, None, None, 1, 12, 16, 16, 0, None) == _lib.FLOWK_ERR_ARG
    assert L.flowk_mixlogcdf_fwd(one, one, one, one, None, None, None, 1, 13, 16, 32, 0, None) == _lib.FLOWK_ERR_SHAPE
    assert L.flowk_affine_coupling_fwd(one, one, one, None, one, None, 1, 12, 16, None) == _lib.FLOWK_ERR_ARG
    assert L.flowk_channel_mix(one, one, None, one, None, None, None, 1, 6, 2, 2, 1, 0, None) == _lib.FLOWK_ERR_SHAPE
    with pytest.raises(AssertionError):
        _lib.check(_lib.FLOWK_ERR_SHAPE, "x")
    with pytest.raises(RuntimeError):
        _lib.check(_lib.FLOWK_ERR_ARG, "x")


def test_no_cpu_fallback():
    from flowk.flow_modules.common_modules import squeeze2d
    with pytest.raises((NotImplementedError, RuntimeError)):
        squeeze2d(torch.zeros(1, 1, 2, 2), 2)


def test_product_does_not_import_oracle():
    import os
    import re
    pkg = os.path.dirname(_lib.LIB_PATH)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
