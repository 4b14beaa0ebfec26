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
    assert _lib.lib.flowk_abi_version() == 3
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


def test_training_entry_points_validate_without_gpu():
    """The training-path entry points (wgrad, weight norm, pointwise, attention, Adamax) reject bad shapes / null
    pointers before any launch, and their planning queries are pure host functions."""
    L = _lib.lib
    one = ctypes.c_void_p(16)
    tr = ctypes.c_int(-1)
    # planning: level-1 3x3 layer splits K over CTAs; gy on the 128-lane side (not transposed); out_conv swaps sides
    assert L.flowk_conv_wgrad_splits(64, 16, 16, 192, 96, 9, ctypes.byref(tr)) == 16 and tr.value == 0
    assert L.flowk_conv_wgrad_splits(64, 16, 16, 96, 588, 9, ctypes.byref(tr)) >= 1 and tr.value == 1
    assert L.flowk_conv_wgrad_splits(64, 4, 4, 192, 96, 9, ctypes.byref(tr)) >= 1          # 4x4 maps: 16-pixel k-blocks
    assert L.flowk_conv_wgrad_splits(64, 2, 4, 192, 96, 9, ctypes.byref(tr)) == 0          # H*W % 16
    assert L.flowk_conv_wgrad_splits(64, 16, 16, 192, 96, 4, None) == 0                    # taps must be 1 or 9
    assert L.flowk_linear_wgrad_splits(16384, 96, 288, ctypes.byref(tr)) >= 1 and tr.value == 1
    assert L.flowk_linear_wgrad_splits(16384, 100, 288, None) == 0                         # K % 32
    assert L.flowk_conv_wgrad(one, None, None, one, one, None, 64, 16, 16, 192, 96, 9, None) == _lib.FLOWK_ERR_ARG
    assert L.flowk_conv_wgrad(one, one, one, one, one, None, 64, 2, 4, 192, 96, 9, None) == _lib.FLOWK_ERR_SHAPE
    assert L.flowk_linear_wgrad(one, None, one, None, 16384, 96, 288, None) == _lib.FLOWK_ERR_ARG
    assert L.flowk_shift_columns(one, one, one, 100, 6, None) == _lib.FLOWK_ERR_SHAPE
    assert L.flowk_weight_norm_operands(one, one, 0, 8, 9, 32, 32, one, None, None, None, None, None, None) == _lib.FLOWK_ERR_SHAPE
    assert L.flowk_weight_norm_operands(one, None, 8, 8, 9, 32, 32, one, None, None, None, None, None, None) == _lib.FLOWK_ERR_ARG
    assert L.flowk_weight_norm_operands_batched(None, 3, 8, None) == _lib.FLOWK_ERR_ARG
    assert L.flowk_weight_norm_operands_batched(None, 0, 8, None) == _lib.FLOWK_OK
    assert L.flowk_weight_norm_bwd_partials(one, one, one, None, one, one, 8, 8, 9, 2, 0, None) == _lib.FLOWK_ERR_ARG
    assert L.flowk_concat_elu_fwd(None, None, None, 0, 4, 1, None) == _lib.FLOWK_OK        # empty batch
    assert L.flowk_concat_elu_fwd(one, one, None, 2, 0, 1, None) == _lib.FLOWK_ERR_SHAPE
    assert L.flowk_glu_bwd(one, None, one, 2, 4, 1, None) == _lib.FLOWK_ERR_ARG
    assert L.flowk_add_layernorm_fwd(one, one, one, one, one, one, one, one, 48, 96, 32, 1, 0, 1e-5, None) == _lib.FLOWK_ERR_SHAPE
    assert L.flowk_add_layernorm_fwd(one, one, None, one, one, one, one, one, 64, 96, 32, 1, 0, 1e-5, None) == _lib.FLOWK_ERR_ARG
    assert L.flowk_add_layernorm_workspace_bytes(16384, 96) == 512 * 2 * 96 * 4
    assert L.flowk_channel_sum(one, one, None, 4, 8, 16, None) == _lib.FLOWK_ERR_ARG
    assert L.flowk_channel_sum_workspace_bytes(96) == 64 * 96 * 4
    assert L.flowk_attention_train_fwd(one, one, one, None, 1, 0.2, 2, 12, 96, 4, None) == _lib.FLOWK_ERR_SHAPE    # HW % 8
    assert L.flowk_attention_train_fwd(one, one, one, None, 1, 1.0, 2, 16, 96, 4, None) == _lib.FLOWK_ERR_SHAPE    # p < 1
    assert L.flowk_attention_train_fwd(one, None, one, None, 1, 0.2, 2, 16, 96, 4, None) == _lib.FLOWK_ERR_ARG
    assert L.flowk_attention_train_bwd(3, one, one, one, one, None, None, 1, 0.2, 2, 16, 96, 4, None) == _lib.FLOWK_ERR_ARG
    assert L.flowk_attention_train_bwd(4, one, one, one, one, one, None, 1, 0.2, 2, 16, 96, 4, None) == _lib.FLOWK_ERR_ARG
    assert L.flowk_attention_dropout_mask(None, 1, 0.2, 0, 16, one, None) == _lib.FLOWK_ERR_SHAPE
    assert L.flowk_adamax_step(None, 0, None, 0.9, 0.999, 1e-8, None) == _lib.FLOWK_OK
    assert L.flowk_adamax_step(None, 4, one, 0.9, 0.999, 1e-8, None) == _lib.FLOWK_ERR_ARG
    # split-K planning of the conv GEMM: the level-3 out_conv input-gradient splits, a level-1 layer does not
    def slices(B, H, W, Cin, N, taps, out_mask):
        args = _lib.ConvGemmArgs(None, None, None, None, None, None, None, None, None, None, None, None, None, None, None,
                                 B, H, W, Cin, N, taps, _lib.PRE_BIAS, out_mask, None, None, None, 0, None)
        return L.flowk_conv_gemm_splitk_slices(ctypes.addressof(args))
    assert slices(64, 4, 4, 2368, 96, 9, _lib.OUT_NCHW) > 1
    assert slices(64, 16, 16, 192, 96, 9, _lib.OUT_NCHW) == 1
    assert slices(64, 4, 4, 2368, 96, 9, _lib.OUT_HILO) == 1                               # only plain fp32 destinations


def test_structs_mirror_the_header():
    """ctypes mirrors of the C structs have the layout the header declares (8-byte pointers, 4-byte ints)."""
    assert ctypes.sizeof(_lib.WnJob) == 8 * 8 + 6 * 4
    assert ctypes.sizeof(_lib.AdamaxChunk) == 4 * 8 + 8
    # N2 is padded to 8 before splitk_ws; then operand_format (int) + acc_scale (float), dilation (int) + reserved (int)
    assert ctypes.sizeof(_lib.ConvGemmArgs) == 15 * 8 + 8 * 4 + 3 * 8 + 8 + 8 + 8 + 8 + 8


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
