"""Inverse (sampling) pass of the flow stack: decode_latents on latents produced by the forward pass, CUDA-graph replay.
Prints images/s and the round-trip error for the BASELINE configs that name the inverse pass (cfg3, cfg5) and cfg2."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda:0")
for wl in sys.argv[1:] or ["cfg3", "cfg5", "cfg2"]:
    coupling, image, L, K, hidden, batch = bench.WORKLOADS[wl]
    model = bench.build_model(wl, dev)
    x = bench.synthetic_batches(1, batch, image, seed=7)[0].to(dev)
    with torch.no_grad():
        z, outs, ld = model.flow.encode_latents(x, x.new_zeros(batch))
        for _ in range(3):
            xr, ldr = model.flow.decode_latents(z, outs, with_logdet=True)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            xr, ldr = model.flow.decode_latents(z, outs, with_logdet=True)
        g.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 10
        s.record()
        for _ in range(steps):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / steps
    print("%s inverse: %.2f ms/batch, %.0f images/s; round trip max|x - x'| %.2e, max|ld_fwd + ld_inv| %.2e" % (
        wl, ms, batch / ms * 1e3, float((xr - x).abs().max()), float((ld + ldr).abs().max())))
