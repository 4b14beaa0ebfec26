"""The mAR channel prior's networks on the flowk tensor-core kernels (inference / sampling, autograd off).

Reference: mar_prior/corr_prior.py:58-139 (likelihood / ancestral sampling over the channel sequence),
mar_prior/lstm.py:7-43 (conv_embed -> Conv2dLSTM -> conv_out1), mar_prior/convolutional_rnn/functional.py:30-52 (the
LSTM cell) and :248-275 ("same" padding of dilated kernels).

Everything is a `flowk_conv_gemm` launch over NHWC rows in TIME-MAJOR order (row = (t, b, pixel)), fp16 (hi, lo) operand
pairs (fp32-accurate two-term split, csrc/tc_gemm.cu):

    conv_embed              k x k conv of all T steps at once                     -> operand pair [T*B*HW, E]
    per LSTM layer          input-to-hidden gates of all T steps in ONE GEMM      -> fp32 [T*B*HW, 4E]
                            T recurrent launches: hidden-to-hidden k x k (dilated) conv with the cell update in the
                            epilogue (FLOWK_PRE_LSTM): gates = W_hh * h_{t-1} + b_hh + gi_t, c_t, h_t; h_t is written
                            as the operand pair the next step AND the next layer read, c_t as fp32 rows
    conv_out1               3 x 3 conv -> (mean, log-std)
    z1_cond_network         5 x 5 conv, ReLU (epilogue), 5 x 5 conv

so one likelihood evaluation of a level is 3 + L (T + 1) launches and nothing but the T-step recurrence is sequential.
"""
import torch

from .. import _lib, tc

ENABLED = True


def usable(module, x):
    """CUDA tensors, autograd off, feature maps the GEMM tiles over (W | 128, H*W | 128 or 128 | H*W)."""
    if not (ENABLED and x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled()):
        return False
    h, w = x.shape[-2], x.shape[-1]
    if w > 128 or 128 % w:
        return False
    hw = h * w
    return (hw % 128 == 0) if hw >= 128 else (128 % hw == 0)


def _pad8(c):
    return (c + 7) // 8 * 8


class _Cache:
    def __init__(self):
        self.key, self.value = None, None

    def get(self, params, build):
        key = _lib.param_key(params)
        if key != self.key:
            with torch.no_grad():
                self.value = build()
            self.key = key
        return self.value


def _prep(weight, bias, pad_out=None):
    """conv weight [N, Cin, k, k] (+ bias [N]) -> ((w_hi, w_lo, acc_scale), bias, taps); N zero-padded to `pad_out`."""
    w, b = weight.detach(), bias.detach()
    if pad_out is not None and pad_out > w.shape[0]:
        w = torch.cat([w, w.new_zeros(pad_out - w.shape[0], *w.shape[1:])], 0)
        b = torch.cat([b, b.new_zeros(pad_out - b.shape[0])], 0)
    return tc.conv_weight_operand_f16(w), b.contiguous(), w.shape[2] * w.shape[3]


def _encoder_operands(enc):
    lstm = enc.lstm
    ops = {"embed": _prep(enc.conv_embed.weight, enc.conv_embed.bias),
           "out": _prep(enc.conv_out1.weight, enc.conv_out1.bias, pad_out=4), "layers": []}
    for layer in range(lstm.num_layers):
        ops["layers"].append((
            _prep(getattr(lstm, "weight_ih_l%d" % layer), getattr(lstm, "bias_ih_l%d" % layer)),
            _prep(getattr(lstm, "weight_hh_l%d" % layer), getattr(lstm, "bias_hh_l%d" % layer))))
    return ops


def _conv(a_hi, a_lo, prepared, images, h, w, cin, n, pre, mask, dilation=1, **kw):
    (w_hi, w_lo, sc), bias, taps = prepared
    tc.conv_gemm(a_hi, a_lo, w_hi, w_lo, images, h, w, cin, n, taps, pre, mask, bias=bias, acc_scale=sc, dilation=dilation,
                 **kw)


def run_sequence(enc, x, state=None):
    """ConvSeqEncoder.forward on the flowk kernels.  x [B, T, Cin, H, W]; `state` = per-layer [(h_hi, h_lo), c] from a
    previous call (ancestral sampling feeds one step at a time) or None.  Returns (out [B, T, 2, H, W], state)."""
    B, T, cin, H, W = x.shape
    HW, E = H * W, enc.embed_ch
    rows, step_rows = T * B * HW, B * HW
    dev = x.device
    lstm = enc.lstm
    cache = enc.__dict__.setdefault("_flowk_cache", _Cache())
    ops = cache.get(list(enc.parameters()), lambda: _encoder_operands(enc))

    def f16(*shape):
        return torch.empty(*shape, device=dev, dtype=torch.float16)

    def f32(*shape):
        return torch.empty(*shape, device=dev, dtype=torch.float32)

    cpad = _pad8(cin)
    x_tb = x.transpose(0, 1).reshape(T * B, cin, H, W).contiguous()                # time-major images
    a_hi, a_lo = tc.nchw_to_nhwc_hilo(x_tb, cpad, True)
    x_hi, x_lo = f16(rows, E), f16(rows, E)
    _conv(a_hi, a_lo, ops["embed"], T * B, H, W, cpad, E, tc.PRE_BIAS, tc.OUT_HILO, out_hi=x_hi, out_lo=x_lo)
    new_state = []
    for layer, (ih, hh) in enumerate(ops["layers"]):
        gi = f32(T, step_rows, 4 * E)                                                # input-to-hidden gates, all steps
        _conv(x_hi, x_lo, ih, T * B, H, W, E, 4 * E, tc.PRE_BIAS, tc.OUT_F32, dilation=lstm.dilation, out_f32=gi)
        h_hi, h_lo = f16(T, step_rows, E), f16(T, step_rows, E)
        if state is None:
            hp_hi = hp_lo = torch.zeros(step_rows, E, device=dev, dtype=torch.float16)
            c_prev = torch.zeros(step_rows, E, device=dev, dtype=torch.float32)
        else:
            (hp_hi, hp_lo), c_prev = state[layer]
        for t in range(T):                                                           # the recurrence: one launch per step
            c_next = f32(step_rows, E)
            _conv(hp_hi, hp_lo, hh, B, H, W, E, 4 * E, tc.PRE_LSTM, 0, dilation=lstm.dilation, res=gi[t], gamma=c_prev,
                  out_f32=c_next, out_hi=h_hi[t], out_lo=h_lo[t])
            hp_hi, hp_lo, c_prev = h_hi[t], h_lo[t], c_next
        new_state.append(((hp_hi, hp_lo), c_prev))
        x_hi, x_lo = h_hi.view(rows, E), h_lo.view(rows, E)
    out = f32(rows, 4)
    _conv(x_hi, x_lo, ops["out"], T * B, H, W, E, 4, tc.PRE_BIAS, tc.OUT_F32, out_f32=out)
    out = out.view(T, B, H, W, 4)[..., :enc.out_ch].permute(1, 0, 4, 2, 3).contiguous()
    return out, new_state


def z1_embedding(net, z1):
    """z1_cond_network (corr_prior.py:26-28): conv5x5 -> ReLU -> conv5x5 on the half that continues through the flow.
    z1 [B, nc, H, W] (a channel slice of a contiguous NCHW tensor) -> [B, 4, H, W]."""
    B, nc, H, W = z1.shape
    dev = z1.device
    cache = net.__dict__.setdefault("_flowk_cache", _Cache())
    ops = cache.get(list(net.parameters()), lambda: (_prep(net[0].weight, net[0].bias), _prep(net[2].weight, net[2].bias)))
    cpad = _pad8(nc)
    if not (z1.stride(3) == 1 and z1.stride(2) == W and z1.stride(1) == H * W):
        z1 = z1.contiguous()
    a_hi, a_lo = tc.nchw_to_nhwc_hilo(z1, cpad, True)
    mid = ops[0][0][0].shape[0]
    h_hi = torch.empty(B * H * W, mid, device=dev, dtype=torch.float16)
    h_lo = torch.empty_like(h_hi)
    _conv(a_hi, a_lo, ops[0], B, H, W, cpad, mid, tc.PRE_BIAS, tc.OUT_HILO_RELU, out_hi=h_hi, out_lo=h_lo)
    n_out = ops[1][0][0].shape[0]
    out = torch.empty(B * H * W, n_out, device=dev, dtype=torch.float32)
    _conv(h_hi, h_lo, ops[1], B, H, W, mid, n_out, tc.PRE_BIAS, tc.OUT_F32, out_f32=out)
    return out.view(B, H, W, n_out).permute(0, 3, 1, 2).contiguous()
