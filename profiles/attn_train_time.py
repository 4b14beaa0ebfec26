"""Timing of the training attention core (forward, backward) per level."""
import sys, torch
sys.path.insert(0, '/root/repo')
import flowk
from flowk import tc_autograd as ta
dev = torch.device('cuda:0')
torch.manual_seed(0)
for S in (256, 64, 16):
    qkv = torch.randn(64, S, 288, device=dev, requires_grad=True)
    dout = torch.randn(64, S, 96, device=dev)
    def fwd(): return ta.attention_core(qkv, 4, 0.2, 3)
    out = fwd()
    def bwd(): out.backward(dout, retain_graph=True)
    def t(f, n=30):
        for _ in range(5): f()
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): f()
        b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n * 1e3
    print("S=%d  fwd %.1f us   bwd %.1f us" % (S, t(fwd), t(bwd)))
