import sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
import flowk
from flowk.marscf import MarScfFlow
from flowk.flow_modules import mixlogcdf_nn
dev = torch.device('cuda:0')
def run(tag):
    torch.manual_seed(0); np.random.seed(0)
    model = MarScfFlow(64, (32,32,3), 'mixlogcdf', 3, 4, 96).to(dev).train()
    x = torch.rand(64,3,32,32, device=dev) - 0.5
    with torch.no_grad(): model(x)
    opt = torch.optim.Adamax(model.parameters(), lr=1e-4)
    def step():
        opt.zero_grad(set_to_none=True)
        _, nll, _ = model(x); nll.mean().backward(); opt.step()
    for _ in range(2): step()
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(4): step()
    torch.cuda.synchronize(); print(tag, (time.time()-t0)/4*1e3, 'ms/step')
run('baseline NCHW')
orig = mixlogcdf_nn.NN.forward_raw
def fr(self, x, aux=None):
    if self.training or torch.is_grad_enabled():
        x = x.contiguous(memory_format=torch.channels_last)
    return orig(self, x, aux)
mixlogcdf_nn.NN.forward_raw = fr
run('channels_last')
torch.backends.cudnn.benchmark = True
run('channels_last + cudnn.benchmark')
mixlogcdf_nn.NN.forward_raw = orig
run('NCHW + cudnn.benchmark')
