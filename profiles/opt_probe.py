"""How long does the optimizer part of the graphed training step take?  (fwd+bwd graph vs Adamax graph, replayed alone)"""
import os, sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import flowk
from flowk.marscf import MarScfFlow
from flowk import sharding
dev = torch.device('cuda:0')
torch.manual_seed(0); np.random.seed(0)
model = MarScfFlow(64, (32, 32, 3), 'mixlogcdf', 3, 4, 96).to(dev).train()
x = torch.rand(64, 3, 32, 32, device=dev) - 0.5
with torch.no_grad(): model(x)
tr = sharding.ShardedTrainer(model, lr=1e-4, warm_up=10000, global_batch=64)
for _ in range(5): tr.step(x)
torch.cuda.synchronize()
fb, up, sx, loss = tr._graphs
def t(f, n=10):
    for _ in range(2): f()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
print("fwd+bwd graph %.2f ms   optimizer graph %.2f ms   full step %.2f ms" % (t(fb.replay), t(up.replay), t(lambda: tr.step(x))))
n = sum(p.numel() for p in model.parameters()); print("params %.1f M in %d tensors" % (n / 1e6, len(list(model.parameters()))))
