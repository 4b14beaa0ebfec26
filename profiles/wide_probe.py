import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import flowk
from flowk import conditioner_tc
from flowk.marscf import MarScfFlow
dev = torch.device('cuda:0')
for hidden in (160, 256):
    torch.manual_seed(0); np.random.seed(0)
    B = 8
    model = MarScfFlow(B, (32,32,3), 'mixlogcdf', 3, 1, hidden, num_blocks=2).to(dev)
    x = torch.rand(B,3,32,32, device=dev) - 0.5; noise = torch.rand_like(x)
    model.train()
    with torch.no_grad():
        model(x, noise=noise)
        for p in model.parameters(): p.add_(torch.randn_like(p)*0.02)
    model.eval()
    with torch.no_grad():
        z1, n1, _ = model(x, noise=noise)
        conditioner_tc.ENABLED = False
        z2, n2, _ = model(x, noise=noise)
        conditioner_tc.ENABLED = True
    print("hidden", hidden, "z rel err", float((z1-z2).abs().max()/z2.abs().max()), "bits/dim err", float((n1-n2).abs().max()))
