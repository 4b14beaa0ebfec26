"""Time the attention kernels in isolation (CUDA graph of back-to-back launches) and print the tcgen05 kernel's phase stamps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowk  # noqa: E402,F401
from flowk import _lib, tc  # noqa: E402

dev = torch.device("cuda:0")
B, heads = 64, 4
for HW, C in ((256, 96), (64, 96), (16, 96), (256, 160)):
    qkv = torch.randn(B * HW, 3 * C, device=dev)
    for name, use_tc in (("tcgen05", True), ("mma.sync", False)):
        tc.ATTENTION_TC = use_tc
        if use_tc and not tc.attention_tc_supported(HW, C, heads):
            continue
        for _ in range(3):
            tc.attention(qkv, B, HW, C, heads, True)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                tc.attention(qkv, B, HW, C, heads, True)
        g.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        e.record()
        torch.cuda.synchronize()
        us = s.elapsed_time(e) * 1e3 / 20
        line = "%-9s seq %4d C %3d: %7.1f us" % (name, HW, C, us)
        if use_tc:
            trace = torch.zeros(8, dtype=torch.int64, device=dev)
            tc.attention(qkv, B, HW, C, heads, True, trace=trace)
            torch.cuda.synchronize()
            t = trace.cpu().tolist()
            line += "   cycles: stage %d  S-mma %d  max-pass %d  exp-pass %d  PV-mma %d  epilogue %d" % (
                t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5])
        print(line)
    tc.ATTENTION_TC = True
