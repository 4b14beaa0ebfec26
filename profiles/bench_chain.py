"""gate (GLU+res+LN) GEMM with and without the chained in_proj GEMM, plus the stand-alone in_proj launch."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowk  # noqa
from flowk import tc
dev = torch.device("cuda:0")
B, C = 64, 96
for (H, W) in ((16, 16), (8, 8), (4, 4)):
    M = B * H * W
    a_hi, a_lo = tc.split_hilo(torch.randn(M, 2 * C, device=dev))
    w_hi, w_lo = tc.split_hilo(torch.randn(2 * C, 2 * C, device=dev) / 14)
    w2_hi, w2_lo = tc.split_hilo(torch.randn(3 * C, C, device=dev) / 10)
    bias, res = torch.randn(2 * C, device=dev), torch.randn(M, C, device=dev)
    gamma, beta, pos = torch.ones(C, device=dev), torch.zeros(C, device=dev), torch.randn(H * W, C, device=dev)
    x1, qkv = torch.empty(M, C, device=dev), torch.empty(M, 3 * C, device=dev)
    p_hi, p_lo = torch.empty(M, C, device=dev), torch.empty(M, C, device=dev)
    common = dict(bias=bias, res=res, gamma=gamma, beta=beta, pos=pos, out_f32=x1)
    def chained(trace=None):
        tc.conv_gemm(a_hi, a_lo, w_hi, w_lo, B, H, W, 2 * C, 2 * C, 1, tc.PRE_GLU_RES_LN, tc.OUT_F32, trace=trace,
                     w2_hi=w2_hi, w2_lo=w2_lo, out2_f32=qkv, n2=3 * C, **common)
    def separate(trace=None):
        tc.conv_gemm(a_hi, a_lo, w_hi, w_lo, B, H, W, 2 * C, 2 * C, 1, tc.PRE_GLU_RES_LN, tc.OUT_F32 | tc.OUT_HILO_POS,
                     out_hi=p_hi, out_lo=p_lo, trace=trace, **common)
        tc.conv_gemm(p_hi, p_lo, w2_hi, w2_lo, B, H, W, C, 3 * C, 1, tc.PRE_BIAS, tc.OUT_F32, out_f32=qkv)
    for name, fn in (("chained", chained), ("separate", separate)):
        g = torch.cuda.CUDAGraph()
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for _ in range(10): fn()
        g.replay(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); [g.replay() for _ in range(5)]; e.record(); torch.cuda.synchronize()
        us = s.elapsed_time(e) * 1e3 / 50
        tr = torch.zeros(16, dtype=torch.int64, device=dev)
        fn(trace=tr); torch.cuda.synchronize(); t = tr.cpu().tolist()
        extra = ""
        if name == "chained":
            extra = " a3_written %d  chain_mma_wait %d  epi2 %d" % (t[9] - t[3], t[10] - t[9], t[4] - t[10])
        print("%dx%d %-9s %6.1f us (in graph) | first_full %d mma_loop %d epi_wait %d epilogue %d%s" % (
            H, W, name, us, t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], extra))
