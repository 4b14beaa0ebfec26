"""Host side of the tensor-core conditioner layers: operand preparation (3xTF32 hi/lo split, NHWC
weight layout) and the ctypes call into `flowk_conv_gemm` (include/flowk.h)."""
import ctypes

import torch

from . import _lib
from ._lib import (OUT_F32, OUT_HILO, OUT_HILO_CELU, OUT_HILO_POS, OUT_HILO_RELU, OUT_NCHW, PRE_BIAS,  # noqa: F401
                   PRE_GLU_RES_LN, PRE_LSTM)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def split_hilo(t):
    """t = hi + lo with both parts rounded to nearest TF32 (13 low mantissa bits zero), like the device-side split."""
    t = t.contiguous().float()
    if t.is_cuda and t.numel():
        return split_rows(t)                      # one kernel (the training path re-splits weights every step)

    def rna(v):
        return ((v.view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    hi = rna(t)
    return hi, rna(t - hi)


def conv_weight_operand(w, cin_pad=None):
    """[N, Cin, kh, kw] conv weight (or [N, Cin] linear weight) -> K-major [N, taps*Cin_pad] in (tap, c) order, hi/lo."""
    if w.dim() == 2:
        w = w[:, :, None, None]
    n, cin, kh, kw = w.shape
    cin_pad = cin_pad or ((cin + 31) // 32 * 32)
    wk = w.permute(0, 2, 3, 1)                                   # [N, kh, kw, Cin]
    if cin_pad != cin:
        wk = torch.nn.functional.pad(wk, (0, cin_pad - cin))
    return split_hilo(wk.reshape(n, kh * kw * cin_pad))


def conv_weight_operand_f16(w, gain=None, mode=_lib.PACK_PLAIN, factor=1.0, bias=None):
    """[N, Cin, kh, kw] conv weight (or [N, Cin] linear weight) -> fp16 (hi, lo) pair [N, taps * ceil64(Cin)] in (tap, c)
    order for FLOWK_OPERAND_F16, pre-scaled by a power of two s so that max|w s| lies in [2^14, 2^15): the lo part then
    keeps its 11 bits clear of fp16's subnormal range.  `gain` / `mode` fuse a per-output-channel factor (weight norm, or an
    ActNorm's exp(factor * logs) with `bias` scaled alike) - see flowk_pack_weight_f16.  Two launches per weight.
    Returns (hi, lo, 1 / s[, bias * gain]); 1 / s goes to conv_gemm(acc_scale=)."""
    if w.dim() == 2:
        w = w[:, :, None, None]
    n, cin, kh, kw = w.shape
    cin_pad = (cin + 63) // 64 * 64
    taps = kh * kw
    w = w.detach().contiguous().float()
    _lib.check_device(w, "conv_weight_operand_f16")
    hi = torch.empty(n, taps * cin_pad, device=w.device, dtype=torch.float16)
    lo = torch.empty_like(hi)
    ws = torch.empty(n + 2, device=w.device, dtype=torch.float32)
    g = None if gain is None else gain.detach().reshape(-1).contiguous().float()
    b_in = None if bias is None else bias.detach().reshape(-1).contiguous().float()
    b_out = None if bias is None else torch.empty_like(b_in)
    assert g is None or g.numel() == n
    _lib.call("flowk_pack_weight_f16", w.data_ptr(), _p(g), int(mode), float(factor), _p(b_in), _p(b_out), n, cin, taps,
              cin_pad, hi.data_ptr(), lo.data_ptr(), ws.data_ptr(), _stream())
    acc_scale = float(ws[n + 1])                       # (host sync: cached per weight version)
    return (hi, lo, acc_scale) if bias is None else (hi, lo, acc_scale, b_out)


def split_rows_f16(x, scale=1.0):
    hi = torch.empty(x.shape, device=x.device, dtype=torch.float16)
    lo = torch.empty_like(hi)
    _lib.call("flowk_split_hilo_f16", x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), float(scale), _stream())
    return hi, lo


def _p(t):
    return None if t is None else t.data_ptr()


def conv_gemm(a_hi, a_lo, w_hi, w_lo, B, H, W, Cin, N, taps, pre, out_mask, bias=None, res=None, gamma=None,
              beta=None, pos=None, out_f32=None, out_hi=None, out_lo=None, out_nchw=None, status=None, trace=None,
              w2_hi=None, w2_lo=None, out2_f32=None, n2=0, split_k=False, acc_scale=None, dilation=1, acc_scale2=None,
              acc_scale_ptr=None):
    """The operand format follows the tensors: fp32 tensors = TF32 pairs (FLOWK_OPERAND_TF32), fp16 tensors = fp16 pairs
    (FLOWK_OPERAND_F16, inference; `acc_scale` undoes the weights' power-of-two pre-scaling).
    `split_k=True` (training path: one stream, kernel latency matters) lends the kernel a workspace so that layers
    with few 128-row tiles and a long K loop are shared by several CTAs per output tile."""
    _lib.check_device(a_hi, "conv_gemm")
    args = _lib.ConvGemmArgs(_p(a_hi), _p(a_lo), _p(w_hi), _p(w_lo), _p(bias), _p(res), _p(gamma), _p(beta), _p(pos),
                             _p(out_f32), _p(out_hi), _p(out_lo), _p(out_nchw), _p(status), _p(trace),
                             B, H, W, Cin, N, taps, pre, out_mask, _p(w2_hi), _p(w2_lo), _p(out2_f32), n2, None,
                             _lib.OPERAND_F16 if a_hi.dtype == torch.float16 else _lib.OPERAND_TF32,
                             1.0 if acc_scale is None else float(acc_scale), int(dilation),
                             1.0 if acc_scale2 is None else float(acc_scale2), _p(acc_scale_ptr))
    assert a_hi.dtype == a_lo.dtype == w_hi.dtype == w_lo.dtype, "operand pair formats must agree"
    assert out_hi is None or out_hi.dtype == a_hi.dtype, "out_hi/out_lo are written in the input operand format"
    if split_k:
        slices = _lib.lib.flowk_conv_gemm_splitk_slices(ctypes.addressof(args))
        if slices > 1:
            ws = torch.empty(slices * B * H * W * N, device=a_hi.device, dtype=torch.float32)
            args.splitk_ws = ws.data_ptr()
    _lib.call("flowk_conv_gemm", ctypes.addressof(args), _stream(), meta=(B, H, W, Cin, N, taps, pre))


def nchw_to_nhwc_hilo(x, c_pad, f16=False):
    """x: NCHW view whose (C,H,W) block is contiguous (e.g. a channel slice) -> ([B*HW, c_pad] hi, lo)."""
    b, c, h, w = x.shape
    assert x.stride(3) == 1 and x.stride(2) == w and x.stride(1) == h * w, "channel slice of a contiguous NCHW tensor"
    hi = torch.empty(b * h * w, c_pad, device=x.device, dtype=torch.float16 if f16 else torch.float32)
    lo = torch.empty_like(hi)
    _lib.call("flowk_nchw_to_nhwc_hilo_f16" if f16 else "flowk_nchw_to_nhwc_hilo", x.data_ptr(), x.stride(0), b, c, h * w,
              c_pad, hi.data_ptr(), lo.data_ptr(), _stream())
    return hi, lo


def split_rows(x):
    hi = torch.empty_like(x)
    lo = torch.empty_like(x)
    _lib.call("flowk_split_hilo", x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), _stream())
    return hi, lo


ATTENTION_TC = __import__("os").environ.get("FLOWK_ATTENTION_TC", "1") != "0"


def attention_tc_supported(HW, C, heads):
    """Shapes the tcgen05 attention kernel (csrc/attention_tc.cu) takes; everything else runs on the mma.sync kernel."""
    d = C // max(heads, 1)
    if not (ATTENTION_TC and C % heads == 0 and HW in (128, 256) and d % 8 == 0 and d <= 64 and C % 4 == 0):
        return False
    dk = (d + 15) // 16 * 16                     # shared memory: P (over Q, K) + V^T + epilogue staging + static
    smem = max(4 * (HW // 64) * 128 * 64, 4 * HW * 128) + 2 * (HW // 64) * dk * 128 + 2 * 128 * (d // 8) * 16 + 9 * 1024
    return smem <= 227 * 1024


def attention(qkv, B, HW, C, heads, f16=False, status=None, trace=None):
    """qkv [B*HW, 3C] (k | v | q) -> (hi, lo) [B*HW, C] of softmax(q k^T / sqrt(d)) v."""
    hi = torch.empty(B * HW, C, device=qkv.device, dtype=torch.float16 if f16 else torch.float32)
    lo = torch.empty_like(hi)
    if attention_tc_supported(HW, C, heads):
        _lib.call("flowk_attention_tc", qkv.data_ptr(), hi.data_ptr(), lo.data_ptr(), int(f16), B, HW, C, heads,
                  _p(status), _p(trace), _stream())
        return hi, lo
    _lib.call("flowk_attention_f16" if f16 else "flowk_attention", qkv.data_ptr(), hi.data_ptr(), lo.data_ptr(), B, HW, C,
              heads, _stream())
    return hi, lo


def attention_supported(HW, C, heads):
    if attention_tc_supported(HW, C, heads):
        return True
    return C % heads == 0 and (C // heads) in (8, 16, 24, 32, 40, 64) and (HW <= 256 or HW % 256 == 0) and HW % 4 == 0 \
        and 2 * min(HW, 1 << 30) * (C // heads) * 4 * max(1, 256 // max(HW, 1)) <= 220 * 1024


def chain_supported(C, n2, f16=False):
    """gate -> in_proj fusion: both accumulators in TMEM (2C + n2 <= 512 columns) and the operand tiles in smem."""
    if C % (8 if f16 else 32) or n2 % 16 or 2 * C > 256 or 2 * C + n2 > 512:
        return False
    slab = ((4 * 32 * (C + 4) * 4 + 1023) // 1024) * 1024
    kblocks = (C + 63) // 64 if f16 else C // 32
    return slab + kblocks * 32768 + 2 * n2 * 128 + 2048 <= 227 * 1024
