import os, sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import flowk
from flowk.marscf import MarScfFlow
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda:0')
torch.manual_seed(0); np.random.seed(0)
model = MarScfFlow(64, (32,32,3), 'mixlogcdf', 3, 4, 96).to(dev).train()
x = torch.rand(64,3,32,32, device=dev) - 0.5
with torch.no_grad(): model(x)
opt = torch.optim.Adamax(model.parameters(), lr=1e-4)
def step():
    opt.zero_grad(set_to_none=True)
    _, nll, _ = model(x); nll.mean().backward(); opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
# CPU-side ops sorted by the device time of the kernels they launched
ka = prof.key_averages(group_by_input_shape=True)
rows = [e for e in ka if e.key.startswith("aten::") or e.key.startswith("_") or "Backward" in e.key]
rows = sorted(rows, key=lambda e: -e.device_time_total)[:60]
for e in rows:
    print("%8.2f ms n=%5d  %-40s %s" % (e.device_time_total/1e3, e.count, e.key[:40], str(e.input_shapes)[:110]))
