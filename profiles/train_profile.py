import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ka = prof.key_averages()
tot = sum(e.device_time_total for e in ka)
print("total device time ms", tot/1e3)
for e in sorted(ka, key=lambda e: -e.device_time_total)[:28]:
    print("%8.2f ms %5.1f%% n=%5d  %s" % (e.device_time_total/1e3, 100*e.device_time_total/tot, e.count, e.key[:90]))
