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
