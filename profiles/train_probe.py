"""Is the training step launch-bound?  Eager step vs the same step (fwd + bwd + Adamax) replayed as one CUDA graph."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowk  # noqa: E402,F401
from flowk.marscf import MarScfFlow  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = bool(int(os.environ.get("CUDNN_BENCHMARK", "0")))
torch.manual_seed(0)
np.random.seed(0)
model = MarScfFlow(64, (32, 32, 3), "mixlogcdf", 3, 4, 96).to(dev).train()
x = torch.rand(64, 3, 32, 32, device=dev) - 0.5
with torch.no_grad():
    model(x)
opt = torch.optim.Adamax(model.parameters(), lr=torch.tensor(1e-4, device=dev), capturable=True, foreach=True)


def step():
    opt.zero_grad(set_to_none=False)
    _, nll, _ = model(x)
    loss = nll.mean()
    loss.backward()
    opt.step()
    return loss


s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
t0 = time.time()
for _ in range(4):
    step()
torch.cuda.synchronize()
print("eager    %.1f ms/step" % ((time.time() - t0) / 4 * 1e3))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    loss = step()
torch.cuda.synchronize()
for _ in range(2):
    g.replay()
torch.cuda.synchronize()
t0 = time.time()
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
print("graphed  %.1f ms/step   loss %.4f" % ((time.time() - t0) / 5 * 1e3, float(loss)))
