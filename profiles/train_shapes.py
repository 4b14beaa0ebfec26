"""Per-shape device time of every libflowk entry point in ONE eager training step (cfg2, B=64)."""
import sys, torch, numpy as np, collections
sys.path.insert(0, '/root/repo')
import flowk
from flowk import _lib
from flowk.marscf import MarScfFlow
dev = torch.device('cuda:0')
torch.manual_seed(0); np.random.seed(0)
model = MarScfFlow(64, (32, 32, 3), 'mixlogcdf', 3, 4, 96).to(dev).train()
x = torch.rand(64, 3, 32, 32, device=dev) - 0.5
with torch.no_grad(): model(x)
opt = torch.optim.Adamax(model.parameters(), lr=1e-4)
def step():
    opt.zero_grad(set_to_none=True)
    _, nll, _ = model(x); nll.mean().backward(); opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
_lib.TIMING = {}
step(); torch.cuda.synchronize()
t, _lib.TIMING = _lib.TIMING, None
rows = []
for name, evs in t.items():
    by = collections.defaultdict(list)
    for s, e, meta in evs: by[tuple(meta)].append(s.elapsed_time(e) * 1e3)
    for meta, us in by.items(): rows.append((sum(us), name, meta, len(us), sum(us) / len(us)))
tot = sum(r[0] for r in rows)
print("flowk total %.2f ms in %d launches" % (tot / 1e3, sum(r[3] for r in rows)))
for r in sorted(rows, reverse=True)[:40]:
    print("%8.2f ms  n=%4d  %7.1f us  %-32s %s" % (r[0] / 1e3, r[3], r[4], r[1], r[2]))
