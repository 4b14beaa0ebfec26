"""The mAR channel-prior port (plain torch, CPU-testable) against fixtures from the reference's own prior."""
import torch

import flowk  # noqa: F401
from flowk.mar_prior import ChannelPriorMultiScale


def test_mar_prior_matches_reference(golden):
    g = golden("mar_prior")
    m = g.meta
    prior = ChannelPriorMultiScale(m["B"], 3, m["image_hwc"][0], m["image_hwc"][1], m["L"], mog=False, dp_rate=0,
                                   num_layers=m["num_layers"], hidden_size=m["hidden"])
    prior.load_state_dict(g.sd, strict=True)            # same module tree / keys as the reference
    prior.eval()
    with torch.no_grad():
        ll1 = prior((g["z1"], g["z2"]), 1, reverse=False)
        ll2 = prior(g["zf"], 2, reverse=False)
        torch.testing.assert_close(ll1, g["ll1"], rtol=1e-5, atol=1e-3)
        torch.testing.assert_close(ll2, g["ll2"], rtol=1e-5, atol=1e-3)
        torch.manual_seed(77)
        s2 = prior(None, 2, reverse=True)
        torch.manual_seed(78)
        s1 = prior(g["z1"], 1, reverse=True)
        torch.testing.assert_close(s2, g["s2"], rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(s1, g["s1"], rtol=1e-4, atol=1e-4)


def test_mar_prior_state_dict_keys_match_reference_layout():
    prior = ChannelPriorMultiScale(1, 3, 32, 32, 3, mog=False, dp_rate=0, num_layers=3, hidden_size=32)
    keys = set(prior.state_dict())
    for k in ("prior_list.0.z1_cond_network.0.weight", "prior_list.0.prior_lstm.lstm.weight_ih_l0",
              "prior_list.0.prior_lstm.lstm.bias_hh_l2", "prior_list.2.prior_lstm.conv_embed.weight",
              "prior_list.2.prior_lstm.conv_out1.bias"):
        assert k in keys, k
    assert prior.prior_list[0].prior_lstm.lstm.weight_ih_l0.shape == (128, 32, 5, 5)
    assert prior.prior_list[2].nc == 48


import pytest  # noqa: E402


@pytest.mark.gpu
def test_mar_prior_cuda_kernels_match_reference(golden):
    """The same fixtures through the flowk tensor-core path (mar_prior/cuda_path.py): 5x5 dilated / 5x5 / 3x3 implicit
    GEMMs, ConvLSTM cell fused into the recurrent GEMM's epilogue, time-major rows - likelihood of both levels and
    ancestral sampling (same host-side noise draws as the reference: corr_prior.py:96-101)."""
    from flowk import _lib
    from flowk.mar_prior import cuda_path
    g = golden("mar_prior")
    m = g.meta
    dev = torch.device("cuda:0")
    prior = ChannelPriorMultiScale(m["B"], 3, m["image_hwc"][0], m["image_hwc"][1], m["L"], mog=False, dp_rate=0,
                                   num_layers=m["num_layers"], hidden_size=m["hidden"])
    prior.load_state_dict(g.sd, strict=True)
    prior.to(dev).eval()
    z1, z2, zf = g["z1"].to(dev), g["z2"].to(dev), g["zf"].to(dev)
    _lib.TIMING = {}
    try:
        with torch.no_grad():
            ll1 = prior((z1, z2), 1, reverse=False)
            ll2 = prior(zf, 2, reverse=False)
            torch.manual_seed(77)
            s2 = prior(None, 2, reverse=True, batch_size=m["B"], device=dev)
            torch.manual_seed(78)
            s1 = prior(z1, 1, reverse=True)
        torch.cuda.synchronize()
        metas = list(_lib.TIMING)
    finally:
        _lib.TIMING = None
    assert "flowk_conv_gemm" in metas                       # the prior ran on the tensor-core kernels
    torch.testing.assert_close(ll1.cpu(), g["ll1"], rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(ll2.cpu(), g["ll2"], rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(s2.cpu(), g["s2"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(s1.cpu(), g["s1"], rtol=1e-4, atol=1e-4)
    # and the torch layers on the same device agree (A/B of the two paths)
    cuda_path.ENABLED = False
    try:
        with torch.no_grad():
            ll1_t = prior((z1, z2), 1, reverse=False)
    finally:
        cuda_path.ENABLED = True
    torch.testing.assert_close(ll1, ll1_t, rtol=1e-5, atol=1e-3)


@pytest.mark.gpu
def test_mar_prior_cuda_full_size_matches_torch_layers():
    """The reference's configuration (marscf_main.py:147-148: hidden 32, 3 layers, 3x32x32, L = 3) at batch 8: flowk
    kernels vs the torch layers (fp32 cuDNN) on the same weights, all three levels."""
    from flowk.mar_prior import cuda_path
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    B = 8
    prior = ChannelPriorMultiScale(B, 3, 32, 32, 3, mog=False, dp_rate=0, num_layers=3, hidden_size=32).to(dev).eval()
    pairs = [(torch.randn(B, 6, 16, 16, device=dev), torch.randn(B, 6, 16, 16, device=dev)),
             (torch.randn(B, 12, 8, 8, device=dev), torch.randn(B, 12, 8, 8, device=dev))]
    zf = torch.randn(B, 48, 4, 4, device=dev)
    with torch.no_grad():
        fast = [prior(pairs[0], 1), prior(pairs[1], 2), prior(zf, 3)]
        cuda_path.ENABLED = False
        try:
            ref = [prior(pairs[0], 1), prior(pairs[1], 2), prior(zf, 3)]
        finally:
            cuda_path.ENABLED = True
    for a, b in zip(fast, ref):
        assert float((a - b).abs().max()) <= 1e-4 * max(1.0, float(b.abs().max()))
