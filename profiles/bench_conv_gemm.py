"""Time flowk_conv_gemm on every GEMM shape of the cfg2 conditioner, in isolation (CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowk  # noqa: E402,F401
from flowk import tc  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
FMT = sys.argv[2] if len(sys.argv) > 2 else "f16"           # operand format: f16 | tf32
C = int(sys.argv[3]) if len(sys.argv) > 3 else 96      # conditioner width (cfg2: 96, cfg4: 160)
cin0 = 8 if FMT == "f16" else 32
shapes = []
for (c, H, W) in ((6, 16, 16), (12, 8, 8), (24, 4, 4)):
    shapes += [("in_conv", H, W, cin0, C, 9, tc.PRE_BIAS), ("conv3x3", H, W, 2 * C, C, 9, tc.PRE_BIAS),
               ("gate1x1", H, W, 2 * C, 2 * C, 1, tc.PRE_GLU_RES_LN), ("in_proj", H, W, C, 3 * C, 1, tc.PRE_BIAS),
               ("attn_gate", H, W, C, 2 * C, 1, tc.PRE_GLU_RES_LN), ("out_conv", H, W, C, 98 * c, 9, tc.PRE_BIAS)]
print("%-10s %5s %6s %6s %5s | %9s %9s %8s" % ("layer", "HxW", "Cin", "N", "taps", "us", "TFLOP/s", "MMA-floor us"))
for name, H, W, Cin, N, taps, pre in shapes:
    M = B * H * W
    a = torch.randn(M, Cin, device=dev)
    k = 3 if taps == 9 else 1
    wt = torch.randn(N, Cin, k, k, device=dev) / (taps * Cin) ** 0.5
    odt = torch.float16 if FMT == "f16" else torch.float32
    if FMT == "f16":
        a_hi, a_lo = tc.split_rows_f16(a)
        w_hi, w_lo, sc = tc.conv_weight_operand_f16(wt)
    else:
        a_hi, a_lo = tc.split_hilo(a)
        w_hi, w_lo = tc.conv_weight_operand(wt)
        sc = None
    bias = torch.randn(N, device=dev)
    nout = N // 2 if pre == tc.PRE_GLU_RES_LN else N
    out = torch.empty(M, nout, device=dev)
    oh, ol = torch.empty(M, 2 * nout, device=dev, dtype=odt), torch.empty(M, 2 * nout, device=dev, dtype=odt)
    kw = dict(bias=bias, out_f32=out, out_hi=oh, out_lo=ol, acc_scale=sc)
    mask = {"f32": tc.OUT_F32, "celu": tc.OUT_HILO_CELU, "hilo": tc.OUT_F32 | tc.OUT_HILO}.get(
        os.environ.get("FLOWK_PROBE_MASK", ""), tc.OUT_F32 | tc.OUT_HILO_CELU)
    if pre == tc.PRE_GLU_RES_LN:
        kw.update(res=torch.randn(M, nout, device=dev), gamma=torch.ones(nout, device=dev), beta=torch.zeros(nout, device=dev))
    if name == "out_conv":
        mask = tc.OUT_NCHW
        kw = dict(bias=bias, out_nchw=torch.empty(B, N, H, W, device=dev), acc_scale=sc)
    for _ in range(3):
        tc.conv_gemm(a_hi, a_lo, w_hi, w_lo, B, H, W, Cin, N, taps, pre, mask, **kw)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    graph = torch.cuda.CUDAGraph()                     # back-to-back launches without host launch overhead
    with torch.cuda.graph(graph):
        for _ in range(reps):
            tc.conv_gemm(a_hi, a_lo, w_hi, w_lo, B, H, W, Cin, N, taps, pre, mask, **kw)
    graph.replay()
    torch.cuda.synchronize()
    s.record()
    graph.replay()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / reps
    trace = torch.zeros(16, dtype=torch.int64, device=dev)
    tc.conv_gemm(a_hi, a_lo, w_hi, w_lo, B, H, W, Cin, N, taps, pre, mask, trace=trace, **kw)
    torch.cuda.synchronize()
    t = trace.cpu().tolist()
    tr = "first_full %5d  mma_loop %6d  epi_wait %6d  epilogue %6d  exit %5d" % (
        t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4])
    if pre == tc.PRE_GLU_RES_LN:
        tr += "  | glu %d res %d stats %d emit %d" % (t[6] - t[3], t[7] - t[6], t[8] - t[7], t[4] - t[8])
    flops = 2.0 * M * N * taps * Cin
    ntile = min(N, 256) if pre == tc.PRE_BIAS else N
    n16 = (N + 15) // 16 * 16
    tiles_n = (n16 + 255) // 256 if pre == tc.PRE_BIAS else 1
    per = n16 // tiles_n
    ctas = ((M + 127) // 128) * tiles_n
    waves = (ctas + 147) // 148
    floor_cycles = waves * (taps * Cin // (16 if FMT == 'f16' else 8)) * 3 * (128 * per / 256.0)
    print("%-10s %2dx%-2d %6d %6d %5d | %9.1f %9.1f %8.1f   ctas=%d" % (name, H, W, Cin, N, taps, us, flops / us / 1e6,
                                                              floor_cycles / 1965.0, ctas), " cycles:", tr)
