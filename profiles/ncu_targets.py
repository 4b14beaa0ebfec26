"""The kernels whose rooflines DESIGN.md quotes, one after another, for one `ncu --set full` pass:

    python profiles/ncu_targets.py                       # plain run first (exit code 0), then
    ncu --set full --clock-control none --import-source on -o gpurun_out/r2_targets python profiles/ncu_targets.py

Order of launches (3 each unless noted): conv_gemm 3x3 level 1 (fp16 pairs), gate 1x1 + GLU + LayerNorm level 1,
conv_gemm 3x3 level 3, attention_tc level 1, channel_mix C = 12 / 24 / 48 / 96 (working set >> L2, 2 each),
affine coupling forward (2), MixLogCDF forward at B = 64 and B = 1024 (2 each), patch attention (2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowk  # noqa: E402,F401
from flowk import ops, tc  # noqa: E402

dev = torch.device("cuda:0")
B, C = 64, 96


def gemm(H, W, cin, n, taps, pre, reps=3):
    m = B * H * W
    k = 3 if taps == 9 else 1
    a_hi, a_lo = tc.split_rows_f16(torch.randn(m, cin, device=dev))
    w_hi, w_lo, sc = tc.conv_weight_operand_f16(torch.randn(n, cin, k, k, device=dev) / (taps * cin) ** 0.5)
    nout = n // 2 if pre == tc.PRE_GLU_RES_LN else n
    kw = dict(bias=torch.randn(n, device=dev), acc_scale=sc,
              out_hi=torch.empty(m, 2 * nout, device=dev, dtype=torch.float16),
              out_lo=torch.empty(m, 2 * nout, device=dev, dtype=torch.float16))
    mask = tc.OUT_HILO_CELU
    if pre == tc.PRE_GLU_RES_LN:
        kw.update(res=torch.randn(m, nout, device=dev), gamma=torch.ones(nout, device=dev), beta=torch.zeros(nout, device=dev),
                  out_f32=torch.empty(m, nout, device=dev))
        mask |= tc.OUT_F32
    for _ in range(reps):
        tc.conv_gemm(a_hi, a_lo, w_hi, w_lo, B, H, W, cin, n, taps, pre, mask, **kw)
    torch.cuda.synchronize()


gemm(16, 16, 2 * C, C, 9, tc.PRE_BIAS)
gemm(16, 16, 2 * C, 2 * C, 1, tc.PRE_GLU_RES_LN)
gemm(4, 4, 2 * C, C, 9, tc.PRE_BIAS)
qkv = torch.randn(B * 256, 3 * C, device=dev)
for _ in range(3):
    tc.attention(qkv, B, 256, C, 4, True)
torch.cuda.synchronize()
for c, hw in ((12, 16), (24, 8), (48, 4), (96, 4)):
    bl = min(65535, (512 << 20) // (c * hw * hw * 4))
    x = torch.randn(bl, c, hw, hw, device=dev)
    mat = torch.linalg.qr(torch.randn(c, c, device=dev))[0].contiguous()
    for _ in range(2):
        ops.channel_mix(x, mat, torch.randn(c, device=dev), torch.zeros(bl, device=dev), torch.ones(1, device=dev), False, False)
    torch.cuda.synchronize()
    del x
x = torch.randn(16384, 12, 16, 16, device=dev)
h = torch.randn(16384, 12, 16, 16, device=dev)
for _ in range(2):
    ops.affine_coupling(x, h, torch.zeros(16384, device=dev), False)
torch.cuda.synchronize()
del x, h
for bl in (64, 1024):
    x = torch.randn(bl, 12, 16, 16, device=dev)
    raw = torch.randn(bl, 98 * 6, 16, 16, device=dev)
    for _ in range(2):
        ops.mixlogcdf_coupling(x, raw, torch.ones(6, device=dev), torch.zeros(bl, device=dev), False, True, 32)
    torch.cuda.synchronize()
from flowk.flow_modules.transformer import Transformer_attn  # noqa: E402
m = Transformer_attn(12).to(dev).eval()
x = torch.randn(8192, 12, 16, 16, device=dev)
with torch.no_grad():
    for _ in range(2):
        m(x, logdet=torch.zeros(8192, device=dev))
torch.cuda.synchronize()
print("ok")
