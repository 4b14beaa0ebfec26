"""One launch of the level-1 3x3 weight-gradient kernel (for ncu)."""
import sys, torch
sys.path.insert(0, '/root/repo')
import flowk
from flowk import tc_autograd as ta
dev = torch.device('cuda:0')
torch.manual_seed(0)
x = torch.randn(64, 192, 16, 16, device=dev); gy = torch.randn(64, 96, 16, 16, device=dev)
for _ in range(3):
    part, tr = ta.wgrad_partials(x, gy, 9)
torch.cuda.synchronize()
print("ok", tuple(part.shape), tr)
