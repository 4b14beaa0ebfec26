"""`flowk.flow_modules.transformer.Transformer_attn` (plug-in) against the fixtures produced by the
reference's own class, and inside FlowStep against the oracle.  Device-agnostic torch ops: runs on CPU here, on the GPU
in the `gpu` variant."""
import pytest
import torch

import flowk  # noqa: F401  (registers the package alias)
from flowk.flow_modules.transformer import Transformer_attn
from oracle import flow_oracle as O


def _check(golden, device):
    for name in ("transformer_attn_c12", "transformer_attn_c24"):
        g = golden(name)
        m = Transformer_attn(g["x"].shape[1])
        assert sorted(m.state_dict().keys()) == sorted(g.sd.keys())
        m.load_state_dict(g.sd, strict=True)
        m = m.to(device).eval()
        with torch.no_grad():
            for permute, sfx in ((False, ""), (True, "_perm")):
                y, ld = m(g["x"].to(device), logdet=g["ld0"].to(device), reverse=False, permute=permute)
                torch.testing.assert_close(y.cpu(), g["y" + sfx], rtol=1e-4, atol=2e-5)
                torch.testing.assert_close(ld.cpu(), g["ld" + sfx], rtol=1e-4, atol=1e-3)
                xr, ldr = m(y, logdet=ld, reverse=True, permute=permute)
                torch.testing.assert_close(xr.cpu(), g["x"], rtol=1e-4, atol=2e-5)
                torch.testing.assert_close(ldr.cpu(), g["ld0"], rtol=1e-4, atol=1e-3)


def test_transformer_attn_matches_reference_fixtures_cpu(golden):
    _check(golden, torch.device("cpu"))


def test_transformer_attn_gradients_flow():
    torch.manual_seed(0)
    m = Transformer_attn(12)
    with torch.no_grad():
        m.scale.fill_(3.0)
    x = torch.randn(2, 12, 8, 8, requires_grad=True)
    y, ld = m(x, logdet=torch.zeros(2))
    (y.square().sum() + ld.sum()).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    assert torch.isfinite(x.grad).all()


@pytest.mark.gpu
def test_transformer_attn_matches_reference_fixtures_gpu(golden):
    _check(golden, torch.device("cuda:0"))


@pytest.mark.gpu
def test_flowstep_with_attn_matches_oracle():
    from flowk.marscf import FlowStep
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    step = FlowStep(8, 8, 12, 12, 12, 16, 1.0, "affine", attn=True).to(dev)
    x = torch.randn(4, 12, 8, 8, device=dev)
    step.train()
    with torch.no_grad():
        step(x, torch.zeros(4, device=dev))            # ActNorm data-dependent init
        for mod in (step.attn1, step.attn2):
            mod.scale.fill_(4.0)
        for prm in step.coupling.parameters():
            prm.add_(0.05 * torch.randn_like(prm))
    step.eval()
    sd = {k: v.detach().cpu() for k, v in step.state_dict().items()}
    with torch.no_grad():
        y, ld = step(x, torch.zeros(4, device=dev))
        xr, ldr = step(y, ld, reverse=True)
    y_o, ld_o = O.flow_step(sd, "", x.cpu(), torch.zeros(4), "affine", False, attn=True)
    torch.testing.assert_close(y.cpu(), y_o, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ld.cpu(), ld_o, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(xr.cpu(), x.cpu(), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ldr.cpu(), torch.zeros(4), rtol=0, atol=2e-3)


@pytest.mark.gpu
def test_full_model_with_attn_matches_oracle():
    """Variant A of the fork: every FlowStep carries the two patch-attention layers."""
    from flowk.marscf import MarScfFlow
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    model = MarScfFlow(4, (16, 16, 3), "affine", 2, 2, 16, attn=True).to(dev)
    x = torch.rand(4, 3, 16, 16, device=dev) - 0.5
    noise = torch.rand(4, 3, 16, 16, device=dev)
    model.train()
    with torch.no_grad():
        model(x, noise=noise)
        for name, prm in model.named_parameters():
            if name.endswith(".scale"):
                prm.fill_(4.0)
            elif "coupling" in name:
                prm.add_(0.05 * torch.randn_like(prm))
    model.eval()
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    with torch.no_grad():
        z, nll, _ = model(x, noise=noise)
    z_o, outs, ldj_o, nll_o = O.normal_flow(sd, x.cpu(), noise.cpu(), 2, 2, "affine", attn=True)
    torch.testing.assert_close(z.cpu(), z_o, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(nll.cpu(), nll_o, rtol=1e-3, atol=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("C,H,B", [(12, 16, 64), (24, 8, 64), (48, 4, 64), (12, 32, 5), (96, 4, 3)])
def test_patch_attention_kernel_matches_torch_ops(C, H, B):
    """csrc/patch_attention.cu (the fused inference / sampling kernel) against the torch-op form of the same layer (taken
    when autograd is on) at every level shape of the BASELINE configs, both mask parities, forward and reverse."""
    from flowk import _lib
    dev = torch.device("cuda:0")
    torch.manual_seed(C + H)
    m = Transformer_attn(C).to(dev).eval()
    with torch.no_grad():
        m.scale.fill_(7.0)                                      # make the attention matrices depend visibly on the scores
    x = torch.randn(B, C, H, H, device=dev)
    ld0 = torch.randn(B, device=dev)
    for permute in (False, True):
        for reverse in (False, True):
            _lib.TIMING = {}
            try:
                with torch.no_grad():
                    y, ld = m(x, logdet=ld0, reverse=reverse, permute=permute)
                assert "flowk_patch_attention" in _lib.TIMING
            finally:
                _lib.TIMING = None
            y_t, ld_t = m(x.clone().requires_grad_(), logdet=ld0, reverse=reverse, permute=permute)     # torch ops
            assert float((y - y_t).abs().max()) <= 1e-5 * max(1.0, float(y_t.abs().max()))
            assert float((ld - ld_t).abs().max()) <= 1e-4 * max(1.0, float(ld_t.abs().max()))
