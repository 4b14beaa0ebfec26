"""Pin oracle/flow_oracle.py against fixtures produced by the reference's own modules
(tests/golden/make_golden.py).  CPU only.  fp32 oracle vs fp32 reference: the same ATen
calls in the same order, so the tolerance is a few ulps, not the 1e-4 parity budget."""
import torch

from oracle import flow_oracle as O

TIGHT = dict(rtol=2e-6, atol=2e-6)


def close(a, b, **kw):
    kw = kw or TIGHT
    torch.testing.assert_close(a, b, **kw)


def test_squeeze_index_map_exact(golden):
    g = golden("squeeze")
    assert torch.equal(O.squeeze2d(g["x"]), g["y"])
    assert torch.equal(O.unsqueeze2d(g["y"]), g["back"])
    assert torch.equal(g["back"], g["x"])
    assert torch.equal(O.squeeze2d(g["idx"]), g["idx_squeezed"])


def test_squeeze_rejects_odd():
    import pytest
    with pytest.raises(AssertionError):
        O.squeeze2d(torch.zeros(1, 1, 7, 8))
    with pytest.raises(AssertionError):
        O.unsqueeze2d(torch.zeros(1, 6, 2, 2))


def test_actnorm(golden):
    g = golden("actnorm")
    bias, logs = O.actnorm_init(g["x"], g.meta["scale"])
    close(bias, g["init_bias"])
    close(logs, g["init_logs"])
    y, ldj = O.actnorm(g["x"], bias, logs, g["ldj0"])
    close(y, g["y_init"])
    close(ldj, g["ldj_init"])
    y, ldj = O.actnorm(g["x"], g["bias"], g["logs"], g["ldj0"])
    close(y, g["y"])
    close(ldj, g["ldj"])
    xr, ldjr = O.actnorm(y, g["bias"], g["logs"], ldj, reverse=True)
    close(xr, g["xr"])
    close(ldjr, g["ldjr"])


def test_invconv(golden):
    for name in ("invconv_c12", "invconv_c24"):
        g = golden(name)
        args = [g.sd[k] for k in ("p", "l", "u", "sign_s", "log_s")]
        close(O.invconv_weight(*args, reverse=False), g["w_fwd"])
        close(O.invconv_weight(*args, reverse=True), g["w_rev"], rtol=1e-5, atol=1e-5)
        z, ldj = O.invconv(g["x"], *args, g["ldj0"])
        close(z, g["z"], rtol=1e-5, atol=1e-5)
        close(ldj, g["ldj"])
        xr, ldjr = O.invconv(g["z"], *args, g["ldj"], reverse=True)
        close(xr, g["xr"], rtol=1e-5, atol=1e-5)
        close(ldjr, g["ldjr"])


def test_affine(golden):
    g = golden("affine")
    h = O.affine_conditioner(g.sd, "coupling.NN_net.", g["x"][:, :6])
    close(h, g["h"], rtol=1e-5, atol=1e-5)
    y, ldj = O.affine_coupling(g.sd, "coupling.", g["x"], g["ldj0"])
    close(y, g["y"], rtol=1e-5, atol=1e-5)
    close(ldj, g["ldj"], rtol=1e-5, atol=1e-5)
    xr, ldjr = O.affine_coupling(g.sd, "coupling.", g["y"], g["ldj"], reverse=True)
    close(xr, g["xr"], rtol=1e-5, atol=1e-5)
    close(ldjr, g["ldjr"], rtol=1e-5, atol=1e-5)


def test_mixture_functions(golden):
    g = golden("mixlogcdf_elementwise")
    c = g["x"].shape[1] // 2
    xc = g["x"][:, :c]
    close(O.mix_log_cdf(xc, g["pi"], g["mu"], g["s"]), g["log_cdf"])
    close(O.mix_log_pdf(xc, g["pi"], g["mu"], g["s"]), g["log_pdf"])
    close(O.mix_inv_cdf(g["u"], g["pi"], g["mu"], g["s"]), g["xinv"], rtol=1e-5, atol=1e-5)


def test_mixlogcdf_elementwise(golden):
    g = golden("mixlogcdf_elementwise")
    p = [g[k] for k in ("a", "b", "pi", "mu", "s")]
    y, ldj = O.mixlogcdf_elementwise(g["x"], *p, g["ldj0"])
    close(y, g["y"], rtol=1e-5, atol=1e-5)
    close(ldj, g["ldj"], rtol=1e-5, atol=1e-4)
    xr, ldjr = O.mixlogcdf_elementwise(g["y"], *p, g["ldj"], reverse=True)
    close(xr, g["xr"], rtol=1e-5, atol=1e-5)
    close(ldjr, g["ldjr"], rtol=1e-5, atol=1e-4)


def test_inverse_out_of_range_raises(golden):
    import pytest
    g = golden("mixlogcdf_elementwise")
    bad = g["u"].clone()
    bad.view(-1)[0] = 1.0
    with pytest.raises(RuntimeError, match="outside"):
        O.mix_inv_cdf(bad, g["pi"], g["mu"], g["s"])


def test_mixlogcdf_coupling_with_conditioner(golden):
    g = golden("mixlogcdf_coupling")
    a, b, pi, mu, s = O.mixlogcdf_conditioner(g.sd, "coupling.nn.", g["x"][:, 6:])
    for got, key in ((a, "a"), (b, "b"), (pi, "pi"), (mu, "mu"), (s, "s")):
        close(got, g[key], rtol=2e-5, atol=2e-5)
    y, ldj = O.mixlogcdf_coupling(g.sd, "coupling.", g["x"], g["ldj0"])
    close(y, g["y"], rtol=5e-5, atol=5e-5)
    close(ldj, g["ldj"], rtol=1e-5, atol=1e-4)
    xr, ldjr = O.mixlogcdf_coupling(g.sd, "coupling.", g["y"], g["ldj"], reverse=True)
    close(xr, g["xr"], rtol=5e-5, atol=5e-5)
    close(ldjr, g["ldjr"], rtol=1e-5, atol=1e-4)


def _flownet(golden, name):
    g = golden(name)
    m = g.meta
    z, outs, ldj, nll = O.normal_flow(g.sd, g["x"], g["noise"], m["L"], m["K"], m["coupling"])
    close(z, g["z"], rtol=1e-4, atol=1e-4)
    for i, o in enumerate(outs):
        close(o, g["z2_%d" % i], rtol=1e-4, atol=1e-4)
    close(ldj, g["logdet"], rtol=1e-5, atol=1e-3)
    close(nll, g["nll"], rtol=1e-5, atol=1e-5)
    xr, ldr = O.flownet_decode(g.sd, g["z"], [g["z2_%d" % i] for i in range(len(outs))],
                               m["L"], m["K"], m["coupling"])
    close(xr, g["xr"], rtol=1e-4, atol=1e-4)
    close(ldr, g["ldr"], rtol=1e-5, atol=1e-3)
    # the reference's own round trip: decode(encode(x)) ~ x + noise/256
    close(g["xr"], g["x"] + g["noise"] / 256.0, rtol=1e-3, atol=1e-3)


def test_flownet_affine(golden):
    _flownet(golden, "flownet_affine")


def test_flownet_mixlogcdf(golden):
    _flownet(golden, "flownet_mixlogcdf")


def test_flownet_width32(golden):
    """The width-32 nets (the fixtures the tensor-core conditioner path is pinned to on the GPU)."""
    _flownet(golden, "flownet_affine_h32")
    _flownet(golden, "flownet_mixlogcdf_h32")


def test_float64_oracle_agrees_with_fp32_reference(golden):
    """The same functions in float64 are the high-precision yardstick; fp32 reference outputs
    must sit within the 1e-4 parity budget of them on these well-conditioned fixtures."""
    g = golden("mixlogcdf_elementwise")
    p = [g[k].double() for k in ("a", "b", "pi", "mu", "s")]
    y, ldj = O.mixlogcdf_elementwise(g["x"].double(), *p, g["ldj0"].double())
    close(y.float(), g["y"], rtol=1e-4, atol=1e-4)
    close(ldj.float(), g["ldj"], rtol=1e-4, atol=1e-3)


def test_transformer_attn(golden):
    """The fork's invertible patch attention (flow_modules/transformer.py), SURVEY.md section 8f-2."""
    for name in ("transformer_attn_c12", "transformer_attn_c24"):
        g = golden(name)
        for permute, sfx in ((False, ""), (True, "_perm")):
            y, ld = O.transformer_attn(g.sd, "", g["x"], g["ld0"], False, permute)
            close(y, g["y" + sfx], rtol=1e-5, atol=1e-5)
            close(ld, g["ld" + sfx], rtol=1e-5, atol=1e-4)
            xr, ldr = O.transformer_attn(g.sd, "", g["y" + sfx], g["ld" + sfx], True, permute)
            close(xr, g["xr" + sfx], rtol=1e-5, atol=1e-5)
            close(ldr, g["ldr" + sfx], rtol=1e-5, atol=1e-4)
