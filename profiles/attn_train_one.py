"""One forward + backward of the level-1 training attention core (B=64, 16x16 positions, C=96, 4 heads, p=0.2), for ncu."""
import sys, torch
sys.path.insert(0, '/root/repo')
import flowk
from flowk import tc_autograd as ta
dev = torch.device('cuda:0')
torch.manual_seed(0)
qkv = torch.randn(64, 256, 288, device=dev, requires_grad=True)
dout = torch.randn(64, 256, 96, device=dev)
for _ in range(3):
    out = ta.attention_core(qkv, 4, 0.2, 3)
    out.backward(dout)
torch.cuda.synchronize()
print("ok", tuple(out.shape))
