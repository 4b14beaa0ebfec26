"""GPU parity tests: the flowk modules (C ABI -> sm_100a kernels) against the golden fixtures the
reference produced and against the CPU oracle on seeded inputs.

Tolerance (BASELINE.json north_star): z and logdet within 1e-4 relative in fp32, bits/dim within
1e-3, squeeze/split index maps exact.  "Relative" is taken against the tensor's largest magnitude
(an element-wise rtol is meaningless at zero crossings): |got - ref| <= 1e-4 * max(1, max|ref|).
"""
import math

import numpy as np

import pytest
import torch

from oracle import flow_oracle as O

pytestmark = pytest.mark.gpu
REL = 1e-4


def dev():
    return torch.device("cuda:0")


def parity(got, ref, rel=REL, what=""):
    ref = ref.to(torch.float32)
    got = got.detach().float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = max(1.0, float(ref.abs().max())) if ref.numel() else 1.0
    err = float((got - ref).abs().max()) if ref.numel() else 0.0
    assert err <= rel * scale, "%s: max abs err %.3e > %.1e * %.3g" % (what, err, rel, scale)


@pytest.fixture(scope="module")
def F():
    import flowk  # noqa: F401
    from flowk import ops
    from flowk.flow_modules import common_modules, affine_coupling, mixlogcdf_coupling, log_dist
    from flowk import marscf

    class NS:
        pass
    ns = NS()
    ns.ops, ns.cm, ns.ac, ns.mc, ns.ld, ns.marscf = ops, common_modules, affine_coupling, mixlogcdf_coupling, log_dist, marscf
    return ns


# ---------------------------------------------------------------------------------------------
# squeeze / unsqueeze: exact
# ---------------------------------------------------------------------------------------------
def test_squeeze_exact_golden(F, golden):
    g = golden("squeeze")
    y = F.cm.squeeze2d(g["x"].to(dev()), 2)
    assert torch.equal(y.cpu(), g["y"])
    assert torch.equal(F.cm.unsqueeze2d(y, 2).cpu(), g["x"])
    assert torch.equal(F.cm.squeeze2d(g["idx"].to(dev()), 2).cpu(), g["idx_squeezed"])


@pytest.mark.parametrize("shape,factor", [((64, 3, 32, 32), 2), ((3, 5, 6, 10), 2), ((2, 2, 9, 6), 3),
                                          ((1, 1, 2, 2), 2), ((2, 4, 8, 8), 1), ((0, 3, 4, 4), 2)])
def test_squeeze_exact_oracle(F, shape, factor):
    x = torch.randn(shape, generator=torch.Generator().manual_seed(1))
    y = F.cm.squeeze2d(x.to(dev()), factor)
    assert torch.equal(y.cpu(), O.squeeze2d(x, factor))
    assert torch.equal(F.cm.unsqueeze2d(y, factor).cpu(), x)


def test_squeeze_shape_errors(F):
    with pytest.raises(AssertionError):
        F.cm.squeeze2d(torch.zeros(1, 1, 7, 8, device=dev()), 2)
    with pytest.raises(AssertionError):
        F.cm.unsqueeze2d(torch.zeros(1, 6, 2, 2, device=dev()), 2)


def test_squeeze_layer_and_split_and_flip(F):
    x = torch.randn(2, 4, 4, 6, device=dev())
    sq = F.cm.SqueezeLayer(2)
    y, ld = sq(x, 0.5)
    assert ld == 0.5 and y.shape == (2, 16, 2, 3)
    back, _ = sq(y, 0.5, reverse=True)
    assert torch.equal(back, x)
    sp = F.cm.Split2dMsC(16, 1)
    (z1, z2), _ = sp(y)
    assert torch.equal(z1, y[:, :8]) and torch.equal(z2, y[:, 8:])
    z, _ = sp((z1, z2), reverse=True)
    assert torch.equal(z, y)
    fl = F.cm.TupleFlip()
    f, _ = fl(y)
    assert torch.equal(f, torch.cat((y[:, 8:], y[:, :8]), 1))
    assert torch.equal(fl(f, reverse=True)[0], y)


# ---------------------------------------------------------------------------------------------
# ActNorm / InvConv
# ---------------------------------------------------------------------------------------------
def test_actnorm_golden(F, golden):
    g = golden("actnorm")
    m = F.cm.Actnormlayer(6, 1.0).to(dev())
    m.train()
    y, ldj = m(g["x"].to(dev()), g["ldj0"].to(dev()))
    parity(m.bias, g["init_bias"], what="init bias")
    parity(m.logs, g["init_logs"], what="init logs")
    parity(y, g["y_init"], what="y after init")
    parity(ldj, g["ldj_init"], what="ldj after init")
    assert float(m.is_initialized) == 1.0
    m.load_state_dict({"bias": g["bias"], "logs": g["logs"], "is_initialized": torch.ones(1)})
    m.eval()
    with torch.no_grad():
        y, ldj = m(g["x"].to(dev()), g["ldj0"].to(dev()))
        parity(y, g["y"], what="y")
        parity(ldj, g["ldj"], what="ldj")
        xr, ldjr = m(y, ldj, reverse=True)
        parity(xr, g["xr"], what="xr")
        parity(ldjr, g["ldjr"], what="ldjr")
        y2, none = m(g["x"].to(dev()), None)
        assert none is None and torch.equal(y2, y)


def test_actnorm_eval_never_initialises(F):
    m = F.cm.Actnormlayer(4).to(dev()).eval()
    m(torch.randn(3, 4, 2, 2, device=dev()) * 5 + 3, None)
    assert float(m.is_initialized) == 0.0 and float(m.bias.detach().abs().sum()) == 0.0


@pytest.mark.parametrize("name", ["invconv_c12", "invconv_c24"])
def test_invconv_golden(F, golden, name):
    g = golden(name)
    m = F.cm.InvertibleConv1x1(g.meta["c"]).to(dev())
    m.load_state_dict(g.sd)
    with torch.no_grad():
        w, _ = m.get_weight(g["x"].to(dev()), False)
        parity(w.view(g.meta["c"], -1), g["w_fwd"], what="W")
        z, ldj = m(g["x"].to(dev()), g["ldj0"].to(dev()))
        parity(z, g["z"], what="z")
        parity(ldj, g["ldj"], what="ldj")
        wr, _ = m.get_weight(z, True)
        parity(wr.view(g.meta["c"], -1), g["w_rev"], what="W^-1")
        xr, ldjr = m(z, ldj, reverse=True)
        parity(xr, g["xr"], what="xr")
        parity(ldjr, g["ldjr"], what="ldjr")


@pytest.mark.parametrize("C,H,W,B", [(12, 16, 16, 64), (24, 8, 8, 64), (48, 4, 4, 64), (96, 4, 4, 8), (4, 14, 14, 5),
                                     (20, 3, 5, 3), (12, 64, 64, 40)])
def test_fused_actnorm_invconv_vs_oracle(F, allow_library, C, H, W, B):
    import numpy as np
    np.random.seed(C)
    gen = torch.Generator().manual_seed(C * 7 + H)
    step = F.marscf.FlowStep(H, W, C, C, C, 8, 1.0, "affine").to(dev()).eval()
    with torch.no_grad():
        step.actnormlayer.bias.copy_(torch.randn(1, C, 1, 1, generator=gen) * 0.3)
        step.actnormlayer.logs.copy_(torch.randn(1, C, 1, 1, generator=gen) * 0.2)
    sd = {k: v.detach().cpu() for k, v in step.state_dict().items()}
    x = torch.randn(B, C, H, W, generator=gen)
    ldj0 = torch.randn(B, generator=gen)
    ref, ref_l = O.actnorm(x, sd["actnormlayer.bias"], sd["actnormlayer.logs"], ldj0)
    ic = [sd["invert_1x1_layer." + n] for n in ("p", "l", "u", "sign_s", "log_s")]
    ref, ref_l = O.invconv(ref, *ic, ref_l)
    with torch.no_grad():
        mat, bias, add = step._folded((H, W), False)
        y, l = F.ops.channel_mix(x.to(dev()), mat, bias, ldj0.to(dev()), add, False, False)
        parity(y, ref, what="fused fwd")
        parity(l, ref_l, what="fused fwd ldj")
        mat, bias, add = step._folded((H, W), True)
        xr, lr = F.ops.channel_mix(y, mat, bias, l, add, False, False)
        parity(xr, x, what="fused round trip")
        parity(lr, ldj0, what="fused round trip ldj")
        if C % 4 == 0:   # squeeze folded into the load / unsqueeze into the store
            xu = O.unsqueeze2d(x)
            mat, bias, add = step._folded((H, W), False)
            y2, _ = F.ops.channel_mix(xu.to(dev()), mat, bias, ldj0.to(dev()), add, True, False)
            assert torch.equal(y2, y)
            mat, bias, add = step._folded((H, W), True)
            xr2, _ = F.ops.channel_mix(y, mat, bias, l, add, False, True)
            assert torch.equal(xr2, F.cm.unsqueeze2d(xr, 2))


# ---------------------------------------------------------------------------------------------
# affine coupling
# ---------------------------------------------------------------------------------------------
def test_affine_golden(F, golden, allow_library):       # 6x6 maps: no 128-row tiling
    g = golden("affine")
    m = F.ac.AffineCoupling(12, 12, 16).to(dev())
    m.load_state_dict({k[len("coupling."):]: v for k, v in g.sd.items()})
    m.eval()
    with torch.no_grad():
        parity(m.NN_net(g["x"][:, :6].to(dev())), g["h"], what="conditioner")
        y, ldj = m(g["x"].to(dev()), g["ldj0"].to(dev()))
        parity(y, g["y"], what="y")
        parity(ldj, g["ldj"], what="ldj")
        xr, ldjr = m(y, ldj, reverse=True)
        parity(xr, g["xr"], what="xr")
        parity(ldjr, g["ldjr"], what="ldjr")


@pytest.mark.parametrize("B,C,H,W", [(128, 12, 16, 16), (128, 48, 4, 4), (7, 6, 3, 5), (2, 12, 64, 64), (0, 4, 2, 2)])
def test_affine_elementwise_vs_oracle(F, B, C, H, W):
    gen = torch.Generator().manual_seed(B + C)
    x = torch.randn(B, C, H, W, generator=gen)
    h = torch.randn(B, C, H, W, generator=gen) * 1.5
    ldj0 = torch.randn(B, generator=gen)
    ref_y, ref_l = O.affine_elementwise(x, h, ldj0)
    y, l = F.ops.affine_coupling(x.to(dev()), h.to(dev()), ldj0.to(dev()), False)
    parity(y, ref_y, what="fwd")
    parity(l, ref_l, what="fwd ldj")
    ref_x, ref_lr = O.affine_elementwise(ref_y, h, ref_l, reverse=True)
    xr, lr = F.ops.affine_coupling(y, h.to(dev()), l, True)
    parity(xr, ref_x, what="inv")
    parity(lr, ref_lr, what="inv ldj")


# ---------------------------------------------------------------------------------------------
# logistic mixture / MixLogCDF coupling
# ---------------------------------------------------------------------------------------------
def rand_mix(B, c, H, W, gen, K=32):
    x = torch.randn(B, 2 * c, H, W, generator=gen)
    raw = torch.randn(B, (2 + 3 * K) * c, H, W, generator=gen)
    r5 = raw.view(B, 2 + 3 * K, c, H, W)
    r5[:, 0] *= 0.5
    r5[:, 1] *= 0.5
    r5[:, 2 + 2 * K:] = r5[:, 2 + 2 * K:] * 0.7 - 0.5
    rescale = torch.rand(c, generator=gen) + 0.5
    return x, raw, rescale


def test_mixture_functions_golden(F, golden):
    g = golden("mixlogcdf_elementwise")
    c = g["x"].shape[1] // 2
    xc = g["x"][:, :c].contiguous().to(dev())
    p = [g[k].to(dev()) for k in ("pi", "mu", "s")]
    parity(F.ld.mixture_log_cdf(xc, *p), g["log_cdf"], what="log cdf")
    parity(F.ld.mixture_log_pdf(xc, *p), g["log_pdf"], what="log pdf")
    xinv = F.ld.mixture_inv_cdf(g["u"].to(dev()), *p)
    parity(xinv, g["xinv"], what="inverse cdf")
    bad = g["u"].clone()
    bad.view(-1)[3] = 0.0
    with pytest.raises(RuntimeError, match="outside"):
        F.ld.mixture_inv_cdf(bad.to(dev()), *p)


def test_mixlogcdf_elementwise_golden(F, golden):
    g = golden("mixlogcdf_elementwise")
    B, C, H, W = g["x"].shape
    c = C // 2
    # rebuild the raw conditioner layout from the fixture's (a, b, pi, mu, s): a = 4 * tanh(a_raw)
    a_raw = torch.atanh(g["a"].double() / 4.0).float()
    raw = torch.cat([a_raw.unsqueeze(1), g["b"].unsqueeze(1), g["pi"], g["mu"], g["s"]], dim=1).reshape(B, -1, H, W)
    ones = torch.full((c,), 4.0)
    for flip in (False, True):
        y, ldj = F.ops.mixlogcdf_coupling(g["x"].to(dev()), raw.to(dev()), ones.to(dev()), g["ldj0"].to(dev()),
                                          False, flip, 32)
        want = O.tuple_flip(g["y"]) if flip else g["y"]
        parity(y, want, what="fwd flip=%s" % flip)
        parity(ldj, g["ldj"], what="fwd ldj")
        xr, ldjr = F.ops.mixlogcdf_coupling(y, raw.to(dev()), ones.to(dev()), ldj, True, flip, 32)
        parity(xr, g["xr"], what="inv flip=%s" % flip)
        parity(ldjr, g["ldjr"], what="inv ldj")


@pytest.mark.parametrize("B,c,H,W", [(64, 6, 16, 16), (64, 12, 8, 8), (64, 24, 4, 4), (3, 5, 3, 7), (1, 1, 1, 1),
                                     (0, 2, 2, 2)])
def test_mixlogcdf_elementwise_vs_oracle(F, B, c, H, W):
    gen = torch.Generator().manual_seed(100 + B + c)
    x, raw, rescale = rand_mix(B, c, H, W, gen)
    ldj0 = torch.randn(B, generator=gen)
    a, b, pi, mu, s = O.mixlogcdf_split_params(raw, rescale.view(-1, 1, 1))
    ref_y, ref_l = O.mixlogcdf_elementwise(x, a, b, pi, mu, s, ldj0)
    y, l = F.ops.mixlogcdf_coupling(x.to(dev()), raw.to(dev()), rescale.to(dev()), ldj0.to(dev()), False, False, 32)
    parity(y, ref_y, what="fwd")
    parity(l, ref_l, what="fwd ldj")
    if B == 0:
        return
    xr, lr = F.ops.mixlogcdf_coupling(y, raw.to(dev()), rescale.to(dev()), l, True, False, 32)
    ref_x, ref_lr = O.mixlogcdf_elementwise(ref_y, a, b, pi, mu, s, ref_l, reverse=True)
    # the bisection result is conditioned by 1/pdf: compare in CDF space and, where the density is
    # not tiny, directly
    cdf_got = O.mix_log_cdf(xr.cpu()[:, :c], pi, mu, s).exp()
    cdf_ref = O.mix_log_cdf(ref_x[:, :c], pi, mu, s).exp()
    assert float((cdf_got - cdf_ref).abs().max()) < 5e-6
    dense = O.mix_log_pdf(ref_x[:, :c], pi, mu, s).exp() > 0.02
    assert float(((xr.cpu()[:, :c] - ref_x[:, :c]).abs() * dense).max()) < 3e-4
    parity(xr[:, c:], ref_x[:, c:], what="pass-through")
    parity(lr, ref_lr, rel=2e-4, what="inv ldj")
    parity(xr, x, rel=1e-3, what="round trip")


def test_mixlogcdf_tail_elements_use_log_domain(F):
    """x far outside every component: the linear-domain sums underflow and the kernel must fall back
    to the reference's log-domain formulation (finite log-pdf, not -inf)."""
    gen = torch.Generator().manual_seed(5)
    B, c, H, W, K = 2, 2, 2, 2, 32
    x, raw, rescale = rand_mix(B, c, H, W, gen)
    r5 = raw.view(B, 2 + 3 * K, c, H, W)
    r5[:, 2 + 2 * K:] = -6.5            # very sharp components
    x[:, :c] = -4.5                     # thousands of widths BELOW every component: u underflows to exactly 0 on
                                        # both sides (above them u rounds to 1 or 1-ulp, where the reference's
                                        # 1e-22 log floor makes the log-det jump by ~34 on a 1-ulp difference)
    ldj0 = torch.zeros(B)
    a, b, pi, mu, s = O.mixlogcdf_split_params(raw, rescale.view(-1, 1, 1))
    ref_lp = O.mix_log_pdf(x[:, :c], pi, mu, s)
    assert torch.isfinite(ref_lp).all() and float(ref_lp.max()) < -80
    got = F.ld.mixture_log_pdf(x[:, :c].contiguous().to(dev()), pi.contiguous().to(dev()), mu.contiguous().to(dev()),
                               s.contiguous().to(dev()))
    parity(got, ref_lp, what="tail log pdf")
    y, l = F.ops.mixlogcdf_coupling(x.to(dev()), raw.to(dev()), rescale.to(dev()), ldj0.to(dev()), False, False, 32)
    ref_y, ref_l = O.mixlogcdf_elementwise(x, a, b, pi, mu, s, ldj0)
    parity(l, ref_l, what="tail ldj")


def test_mixlogcdf_coupling_with_conditioner_golden(F, golden, allow_library):
    g = golden("mixlogcdf_coupling")
    m = F.mc.MixLogCDFCoupling(12, 16, 2, 32, 0.2).to(dev())
    m.load_state_dict({k[len("coupling."):]: v for k, v in g.sd.items()})
    m.eval()
    with torch.no_grad():
        a, b, pi, mu, s = m.nn(g["x"][:, 6:].to(dev()))
        for got, key in ((a, "a"), (b, "b"), (pi, "pi"), (mu, "mu"), (s, "s")):
            parity(got, g[key], what="conditioner " + key)
        y, ldj = m(g["x"].to(dev()), g["ldj0"].to(dev()))
        parity(y, g["y"], what="y")
        parity(ldj, g["ldj"], what="ldj")
        xr, ldjr = m(y, ldj, reverse=True)
        parity(xr, g["xr"], what="xr")
        parity(ldjr, g["ldjr"], what="ldjr")


def test_logdet_bit_reproducible(F):
    gen = torch.Generator().manual_seed(9)
    x, raw, rescale = rand_mix(64, 6, 16, 16, gen)
    args = (x.to(dev()), raw.to(dev()), rescale.to(dev()), torch.zeros(64, device=dev()), False, True, 32)
    y0, l0 = F.ops.mixlogcdf_coupling(*args)
    for _ in range(3):
        y1, l1 = F.ops.mixlogcdf_coupling(*args)
        assert torch.equal(l0, l1) and torch.equal(y0, y1)


# ---------------------------------------------------------------------------------------------
# whole stack
# ---------------------------------------------------------------------------------------------
def build_from_golden(F, g, fuse):
    m = g.meta
    model = F.marscf.MarScfFlow(m["B"], tuple(m["image_hwc"]), m["coupling"], m["L"], m["K"], m["hidden"],
                                num_blocks=max(m["blocks"], 1), fuse_squeeze=fuse)
    model.load_state_dict(g.sd, strict=True)
    return model.to(dev()).eval()


@pytest.mark.parametrize("name", ["flownet_affine", "flownet_mixlogcdf"])
@pytest.mark.parametrize("fuse", [True, False])
def test_flownet_golden(F, golden, allow_library, name, fuse):
    """Width-8 nets: narrower than the kernels' channel blocks, so Linear / attention layers opt in to the library."""
    _flownet_golden(F, golden(name), fuse)


@pytest.mark.parametrize("name", ["flownet_affine_h32", "flownet_mixlogcdf_h32"])
def test_flownet_golden_tensor_core(F, golden, no_library, name):
    """Width-32 nets from the reference's own modules: every conditioner layer runs on the tcgen05 / flowk kernels
    (`no_library` fails the test on any cuDNN / cuBLAS / ATen conditioner layer), pinned to the REFERENCE outputs."""
    from flowk import _lib
    _lib.TIMING = {}
    try:
        _flownet_golden(F, golden(name), True)
        names = set(_lib.TIMING)
    finally:
        _lib.TIMING = None
    assert "flowk_conv_gemm" in names, names
    if "mixlogcdf" in name:
        assert names & {"flowk_attention", "flowk_attention_f16"} and "flowk_mixlogcdf_fwd" in names, names


def _flownet_golden(F, g, fuse):
    model = build_from_golden(F, g, fuse)
    x, noise = g["x"].to(dev()), g["noise"].to(dev())
    with torch.no_grad():
        z, nll, _ = model(x, noise=noise)
        parity(z, g["z"], what="z")
        assert float((nll.cpu() - g["nll"]).abs().max()) < 1e-3          # bits/dim budget
        d = x[0].numel()
        z0 = x + noise / 256.0
        zf, outs, logdet = model.flow.encode_latents(z0, x.new_full((x.shape[0],), -math.log(256.0) * d))
        parity(logdet, g["logdet"], what="logdet")
        for i, o in enumerate(outs):
            parity(o, g["z2_%d" % i], what="z2_%d" % i)
        z2s = [g["z2_%d" % i].to(dev()) for i in range(len(outs))]
        xr, ldr = model.flow.decode_latents(g["z"].to(dev()), z2s, with_logdet=True)
        parity(xr, g["xr"], what="decode")
        parity(ldr, g["ldr"], what="decode logdet")


@pytest.mark.parametrize("coupling,hidden,B,image,L", [("affine", 64, 32, (32, 32, 3), 3),
                                                       ("mixlogcdf", 32, 16, (32, 32, 3), 3),
                                                       ("affine", 32, 4, (64, 64, 3), 4)])
def test_full_size_round_trip_properties(F, coupling, hidden, B, image, L):
    """Size-independent properties at the BASELINE shapes: decode(encode(x)) = x,
    logdet_fwd + logdet_rev = 0, fused-squeeze path == explicit squeeze path, batch independence."""
    import numpy as np
    torch.manual_seed(3)
    np.random.seed(3)
    model = F.marscf.MarScfFlow(B, image, coupling, L, 4, hidden, num_blocks=2).to(dev())
    x = torch.rand(B, image[2], image[0], image[1], device=dev()) - 0.5
    model.train()
    with torch.no_grad():
        model(x)                                     # ActNorm data-dependent init
    with torch.no_grad():                            # move the zero-initialised output convs off zero
        gen = torch.Generator(device="cpu").manual_seed(4)
        for n, p in model.named_parameters():
            p.add_((torch.randn(p.shape, generator=gen) * 0.02).to(p.device))
    model.eval()
    with torch.no_grad():
        z, outs, ld = model.flow.encode_latents(x, x.new_zeros(B))
        xr, ldr = model.flow.decode_latents(z, outs, with_logdet=True)
        parity(xr, x.cpu(), rel=2e-3, what="round trip")
        assert float((ld + ldr).abs().max()) < 2e-3 * max(1.0, float(ld.abs().max()))
        model.flow.fuse_squeeze = False
        z_b, outs_b, ld_b = model.flow.encode_latents(x, x.new_zeros(B))
        assert torch.equal(z, z_b) and torch.equal(ld, ld_b)
        model.flow.fuse_squeeze = True
        # each sample is processed independently of its batch neighbours
        z_half, _, ld_half = model.flow.encode_latents(x[: B // 2].contiguous(), x.new_zeros(B // 2))
        parity(z_half, z[: B // 2].cpu(), rel=1e-5, what="batch independence")
        parity(ld_half, ld[: B // 2].cpu(), rel=1e-5, what="batch independence ldj")


# ---------------------------------------------------------------------------------------------
# gradients of the forward ops (training) against autograd through the oracle
# ---------------------------------------------------------------------------------------------
def test_affine_backward_vs_oracle(F):
    gen = torch.Generator().manual_seed(31)
    B, C, H, W = 5, 8, 4, 6
    x = torch.randn(B, C, H, W, generator=gen, dtype=torch.float64, requires_grad=True)
    h = torch.randn(B, C, H, W, generator=gen, dtype=torch.float64, requires_grad=True)
    l0 = torch.randn(B, generator=gen, dtype=torch.float64, requires_grad=True)
    gy = torch.randn(B, C, H, W, generator=gen, dtype=torch.float64)
    gl = torch.randn(B, generator=gen, dtype=torch.float64)
    y, l = O.affine_elementwise(x, h, l0)
    ((y * gy).sum() + (l * gl).sum()).backward()
    xd, hd, ld_ = (t.detach().float().to(dev()).requires_grad_() for t in (x, h, l0))
    yd, lo = F.ops.affine_coupling(xd, hd, ld_, False)
    ((yd * gy.float().to(dev())).sum() + (lo * gl.float().to(dev())).sum()).backward()
    parity(xd.grad, x.grad, what="dx")
    parity(hd.grad, h.grad, what="dh")
    parity(ld_.grad, l0.grad, what="dldj")


@pytest.mark.parametrize("flip", [False, True])
def test_mixlogcdf_backward_vs_oracle(F, flip):
    gen = torch.Generator().manual_seed(32)
    B, c, H, W = 3, 3, 4, 5
    x, raw, rescale = rand_mix(B, c, H, W, gen)
    x64 = x.double().requires_grad_()
    raw64 = raw.double().requires_grad_()
    res64 = rescale.double().requires_grad_()
    l0 = torch.randn(B, generator=gen, dtype=torch.float64, requires_grad=True)
    gy = torch.randn(B, 2 * c, H, W, generator=gen, dtype=torch.float64)
    gl = torch.randn(B, generator=gen, dtype=torch.float64)
    a, b, pi, mu, s = O.mixlogcdf_split_params(raw64, res64.view(-1, 1, 1))
    y, l = O.mixlogcdf_elementwise(x64, a, b, pi, mu, s, l0)
    if flip:
        y = O.tuple_flip(y)
    ((y * gy).sum() + (l * gl).sum()).backward()
    xd, rd, sd_, ld_ = (t.detach().float().to(dev()).requires_grad_() for t in (x, raw, rescale, l0))
    yd, lo = F.ops.mixlogcdf_coupling(xd, rd, sd_, ld_, False, flip, 32)
    ((yd * gy.float().to(dev())).sum() + (lo * gl.float().to(dev())).sum()).backward()
    parity(xd.grad, x64.grad, rel=2e-4, what="dx")
    parity(rd.grad, raw64.grad, rel=2e-4, what="draw")
    parity(sd_.grad, res64.grad, rel=2e-4, what="drescale")
    parity(ld_.grad, l0.grad, what="dldj")


def test_flowstep_backward_vs_oracle(F, golden, allow_library):
    """Gradients w.r.t. every parameter of a small affine FlowNet, against autograd through the oracle."""
    g = golden("flownet_affine")
    model = build_from_golden(F, g, True)
    m = g.meta
    sd64 = {k: v.double().requires_grad_(v.dtype.is_floating_point and "is_initialized" not in k and
                                         not k.endswith(".p") and not k.endswith("sign_s"))
            for k, v in g.sd.items()}
    x = g["x"].double()
    z, outs, ldj, nll = O.normal_flow(sd64, x, g["noise"].double(), m["L"], m["K"], m["coupling"])
    nll.mean().backward()
    for p in model.parameters():
        p.requires_grad_(True)
    _, nll_d, _ = model(g["x"].to(dev()), noise=g["noise"].to(dev()))
    nll_d.mean().backward()
    checked = 0
    for name, p in model.named_parameters():
        ref = sd64[name].grad
        if ref is None:
            continue
        if name.endswith(".l") or name.endswith(".u"):
            c = ref.shape[0]
            mask = torch.tril(torch.ones(c, c), -1) if name.endswith(".l") else torch.triu(torch.ones(c, c), 1)
            ref = ref * mask
        parity(p.grad, ref, rel=5e-4, what="grad " + name)
        checked += 1
    assert checked > 20


def test_cfg2_architecture_vs_oracle(F, no_library):
    """The BASELINE cfg2 architecture (MixLogCDF, L=3, K=4, C=96, 10 blocks) at a small batch: the whole GPU stack
    (tcgen05 conditioners + fused flow kernels) against the CPU oracle with the same weights."""
    import numpy as np
    torch.manual_seed(0)
    np.random.seed(0)
    B = 4
    model = F.marscf.MarScfFlow(B, (32, 32, 3), "mixlogcdf", 3, 4, 96).to(dev())
    gen = torch.Generator().manual_seed(1)
    x = torch.rand(B, 3, 32, 32, generator=gen) - 0.5
    noise = torch.rand(B, 3, 32, 32, generator=gen)
    model.train()
    with torch.no_grad():
        model(x.to(dev()), noise=noise.to(dev()))
    model.eval()
    with torch.no_grad():
        z, nll, _ = model(x.to(dev()), noise=noise.to(dev()))
        zf, outs, ld = model.flow.encode_latents((x + noise / 256.0).to(dev()), torch.zeros(B, device=dev()))
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    z_ref, outs_ref, ld_ref, nll_ref = O.normal_flow(sd, x, noise, 3, 4, "mixlogcdf")
    parity(z, z_ref, what="z")
    for a, b in zip(outs, outs_ref):
        parity(a, b, what="z2")
    parity(ld, ld_ref - float(-math.log(256.0) * 3072), what="logdet")
    assert float((nll.cpu() - nll_ref).abs().max()) < 1e-3


def test_marscf_with_mar_prior(F):
    """MarScfFlow with the reference's mAR channel prior plugged in: bits/dim = -(flow logdet + prior log-lik) /
    (ln2 D) with the prior evaluated exactly as marscf_main.py:159-164 does, and unconditional sampling runs."""
    import numpy as np
    torch.manual_seed(5)
    np.random.seed(5)
    B = 4
    model = F.marscf.MarScfFlow(B, (32, 32, 3), "affine", 3, 2, 32, prior="mar").to(dev())
    x = torch.rand(B, 3, 32, 32, device=dev()) - 0.5
    noise = torch.rand_like(x)
    model.train()
    with torch.no_grad():
        model(x, noise=noise)
    model.eval()
    with torch.no_grad():
        z, nll, _ = model(x, noise=noise)
        pairs = []
        d = 3 * 32 * 32
        zf, outs, ld = model.flow.encode_latents(x + noise / 256.0, x.new_full((B,), -math.log(256.0) * d), pairs)
        total = ld + model.flow.c_prior(zf, 3) + sum(model.flow.c_prior(p, i + 1) for i, p in enumerate(pairs))
        assert torch.allclose(nll, -total / (math.log(2.0) * d), rtol=1e-5, atol=1e-5)
        sample = model(None, None, reverse=True, eps_std=1.0)
    assert sample.shape == (B, 3, 32, 32) and torch.isfinite(sample).all()
    keys = set(model.state_dict())
    assert "flow.c_prior.prior_list.0.prior_lstm.lstm.weight_ih_l0" in keys


def test_mixlogcdf_flownet_backward_vs_oracle(F, golden, allow_library):
    """Training gradients of a small (width-8) MixLogCDF FlowNet against float64 autograd through the oracle, for every
    parameter: convs and pointwise layers on the flowk forward/backward kernels, the width-8 Linear / attention layers
    on the library (see test_cfg2_training_gradients_vs_oracle for the all-flowk full-width step)."""
    g = golden("flownet_mixlogcdf")
    model = build_from_golden(F, g, True)
    m = g.meta
    sd64 = {k: v.double().requires_grad_(v.dtype.is_floating_point and "is_initialized" not in k and
                                         not k.endswith(".p") and not k.endswith("sign_s"))
            for k, v in g.sd.items()}
    z, outs, ldj, nll = O.normal_flow(sd64, g["x"].double(), g["noise"].double(), m["L"], m["K"], m["coupling"])
    nll.mean().backward()
    for p in model.parameters():
        p.requires_grad_(True)
    _, nll_d, _ = model(g["x"].to(dev()), noise=g["noise"].to(dev()))      # eval mode: dropout off, like the oracle
    nll_d.mean().backward()
    checked = 0
    for name, p in model.named_parameters():
        ref = sd64[name].grad
        if ref is None:
            continue
        if name.endswith(".l") or name.endswith(".u"):
            c = ref.shape[0]
            ref = ref * (torch.tril(torch.ones(c, c), -1) if name.endswith(".l") else torch.triu(torch.ones(c, c), 1))
        parity(p.grad, ref, rel=1e-3, what="grad " + name)
        checked += 1
    assert checked > 40


def test_graphed_training_step_matches_eager(F):
    """ShardedTrainer with the step replayed as CUDA graphs == the eager step (affine model: no dropout; dequantisation
    noise fixed), including the sample-count learning-rate warm-up."""
    import copy
    import numpy as np
    from flowk import sharding

    class FixedNoise(torch.nn.Module):
        def __init__(self, m, noise):
            super().__init__()
            self.m, self.noise = m, noise

        def forward(self, x):
            return self.m(x, noise=self.noise)

    torch.manual_seed(11)
    np.random.seed(11)
    B = 8
    base = F.marscf.MarScfFlow(B, (16, 16, 3), "affine", 2, 2, 32).to(dev())
    x = torch.rand(B, 3, 16, 16, device=dev()) - 0.5
    noise = torch.rand_like(x)
    base.train()
    with torch.no_grad():
        base(x, noise=noise)
        for p in base.parameters():
            p.add_(torch.randn_like(p) * 0.02)
    results = []
    for use_graph in (False, True):
        model = FixedNoise(copy.deepcopy(base), noise)
        trainer = sharding.ShardedTrainer(model, lr=1e-3, warm_up=40, global_batch=B, use_graph=use_graph, graph_after=2)
        losses = [float(trainer.step(x)) for _ in range(5)]
        results.append((losses, [p.detach().clone() for p in model.parameters()]))
    (l0, p0), (l1, p1) = results
    assert l0[-1] < l0[0]                                   # it trains
    for a, b in zip(l0, l1):
        assert abs(a - b) < 1e-4 * max(1.0, abs(a)), (l0, l1)
    for a, b in zip(p0, p1):
        parity(b, a.cpu(), rel=1e-4, what="parameters after 5 steps")


@pytest.mark.parametrize("coupling,hidden,B", [("mixlogcdf", 96, 5), ("mixlogcdf", 32, 1), ("affine", 64, 3),
                                               ("affine", 256, 130)])
def test_odd_batch_sizes_through_tensor_core_path(F, no_library, coupling, hidden, B):
    """Batches that do not fill the 128-row GEMM tiles (TMA zero-fills the missing images, the epilogues mask the rows):
    the tcgen05 conditioner path must agree with the torch/cuDNN path on the same module."""
    import numpy as np
    from flowk import conditioner_tc
    torch.manual_seed(B)
    np.random.seed(B)
    model = F.marscf.MarScfFlow(B, (32, 32, 3), coupling, 3, 2, hidden, num_blocks=2).to(dev())
    x = torch.rand(B, 3, 32, 32, device=dev()) - 0.5
    noise = torch.rand_like(x)
    model.train()
    with torch.no_grad():
        model(x, noise=noise)
        for p in model.parameters():
            p.add_(torch.randn_like(p) * 0.02)
    model.eval()
    with torch.no_grad():
        z_tc, nll_tc, _ = model(x, noise=noise)
        conditioner_tc.ENABLED = False
        try:
            z_ref, nll_ref, _ = model(x, noise=noise)
        finally:
            conditioner_tc.ENABLED = True
    parity(z_tc, z_ref.cpu(), what="z")
    assert float((nll_tc - nll_ref).abs().max()) < 1e-3


@pytest.mark.gpu
def test_fused_adamax_matches_torch_adamax():
    """flowk.optim.FusedAdamax (one launch over all tensors) == torch.optim.Adamax over several steps, incl. a changing
    learning rate, odd tensor sizes (scalar tail / unaligned views) and state_dict interchange."""
    from flowk.optim import FusedAdamax
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(12)
    shapes = [(96, 192, 3, 3), (7,), (3, 5), (40000,), (1, 1, 1), (96,)]
    flat = torch.zeros(sum(int(np.prod(s)) for s in shapes), device=dev)         # gradients as unaligned views of one buffer
    pa = [torch.nn.Parameter(torch.randn(s, generator=g).to(dev)) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    off = 0
    for p in pa:
        p.grad = flat[off:off + p.numel()].view(p.shape)
        off += p.numel()
    oa, ob = FusedAdamax(pa, lr=1e-2), torch.optim.Adamax(pb, lr=1e-2)
    for it in range(6):
        lr = 1e-2 * (it + 1) / 6
        for grp in list(oa.param_groups) + list(ob.param_groups):
            grp["lr"] = lr
        for p, q in zip(pa, pb):
            gr = torch.randn(p.shape, generator=g).to(dev) * (10.0 ** (it - 3))
            p.grad.copy_(gr)
            q.grad = gr.clone()
        oa.step()
        ob.step()
    for p, q in zip(pa, pb):
        assert float((p - q).abs().max()) <= 2e-6 * float(q.abs().max()) + 1e-7
    sa, sb = oa.state_dict()["state"], ob.state_dict()["state"]
    assert sa.keys() == sb.keys()
    for k in sa:
        assert set(sa[k].keys()) == set(sb[k].keys()) == {"step", "exp_avg", "exp_inf"}
        assert float(sa[k]["step"]) == float(sb[k]["step"]) == 6.0
        torch.testing.assert_close(sa[k]["exp_inf"], sb[k]["exp_inf"], rtol=1e-6, atol=1e-12)
    ob2 = torch.optim.Adamax(pb, lr=1e-2)
    ob2.load_state_dict(oa.state_dict())                     # torch's optimizer accepts the fused optimizer's checkpoint


# ---------------------------------------------------------------------------------------------
# every BASELINE.json config at its real width: tcgen05 conditioners + fused flow kernels vs the CPU oracle
# ---------------------------------------------------------------------------------------------
BASELINE_CONFIGS = {
    # name: (coupling, image HWC, L, K, hidden, batch, check inverse against the oracle)
    "cfg1": ("affine", (32, 32, 3), 3, 4, 64, 32, True),
    "cfg2": ("mixlogcdf", (32, 32, 3), 3, 4, 96, 64, False),
    "cfg3": ("affine", (32, 32, 3), 3, 4, 256, 128, True),
    "cfg4": ("mixlogcdf", (32, 32, 3), 3, 4, 160, 8, False),
    "cfg5": ("affine", (64, 64, 3), 4, 4, 256, 8, True),
}


def _initialised_model(F, coupling, image, L, K, hidden, B, seed, **kw):
    import numpy as np
    torch.manual_seed(seed)
    np.random.seed(seed)
    model = F.marscf.MarScfFlow(B, image, coupling, L, K, hidden, **kw).to(dev())
    gen = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(B, image[2], image[0], image[1], generator=gen) - 0.5
    noise = torch.rand(B, image[2], image[0], image[1], generator=gen)
    model.train()
    with torch.no_grad():
        model(x.to(dev()), noise=noise.to(dev()))            # ActNorm data-dependent init (first training batch)
        std = 0.02 if hidden <= 96 else 0.005                # wide nets: keep the (2304-term) output sums O(0.1)
        for p in model.parameters():                         # move the zero-initialised output convs off zero
            p.add_((torch.randn(p.shape, generator=gen) * std).to(p.device))
    return model, x, noise


@pytest.mark.parametrize("cfg", sorted(BASELINE_CONFIGS))
def test_baseline_config_vs_oracle(F, no_library, cfg):
    """Forward (z, every factored-out z2, logdet, bits/dim) of each BASELINE.json config at its REAL width and batch
    through the tensor-core path, against the CPU oracle with the same weights; the affine configs (cfg1, cfg3, cfg5:
    "inverse sampling") also check the inverse pass against the oracle's, the MixLogCDF ones the round trip (their
    bisection is pinned to the oracle in test_mixture_functions_golden / test_mixlogcdf_elementwise_vs_oracle)."""
    from flowk import _lib
    coupling, image, L, K, hidden, B, inverse = BASELINE_CONFIGS[cfg]
    model, x, noise = _initialised_model(F, coupling, image, L, K, hidden, B, seed=100 + sorted(BASELINE_CONFIGS).index(cfg))
    model.eval()
    d = x[0].numel()
    _lib.TIMING = {}
    try:
        with torch.no_grad():
            z, nll, _ = model(x.to(dev()), noise=noise.to(dev()))
            zf, outs, ld = model.flow.encode_latents((x + noise / 256.0).to(dev()), torch.zeros(B, device=dev()))
            xr, ldr = model.flow.decode_latents(zf, outs, with_logdet=True)
        torch.cuda.synchronize()
        names = set(_lib.TIMING)
    finally:
        _lib.TIMING = None
    assert "flowk_conv_gemm" in names, names                   # the tcgen05 conditioner ran
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    z_ref, outs_ref, ld_ref, nll_ref = O.normal_flow(sd, x, noise, L, K, coupling)
    parity(z, z_ref, what=cfg + " z")
    for i, (a, b) in enumerate(zip(outs, outs_ref)):
        parity(a, b, what=cfg + " z2_%d" % i)
    parity(ld, ld_ref - float(-math.log(256.0) * d), what=cfg + " logdet")
    assert float((nll.cpu() - nll_ref).abs().max()) < 1e-3, cfg
    if inverse:
        xr_ref, ldr_ref = O.flownet_decode(sd, z_ref, outs_ref, L, K, coupling)
        parity(xr, xr_ref, what=cfg + " inverse x")
        parity(ldr, ldr_ref, what=cfg + " inverse logdet")
    else:
        parity(xr, x + noise / 256.0, rel=2e-3, what=cfg + " round trip")
        assert float((ld + ldr).abs().max()) < 2e-3 * max(1.0, float(ld.abs().max()))


def test_cfg2_training_gradients_vs_oracle(F, no_library):
    """A cfg2-shaped TRAINING step (MixLogCDF, C=96, all three level shapes 16x16 / 8x8 / 4x4, 2 ConvAttnBlocks,
    train mode with dropout p = 0) against float64 autograd through the oracle, for every parameter.  Exercises the
    batched weight norm, `wn_conv2d` / `wn_linear` forward + dgrad + tcgen05 wgrad, `attention_train_*`,
    `add_layernorm`, concat-ELU / GLU and the flow kernels' backward passes; `no_library` proves none of it fell back."""
    from flowk import _lib
    coupling, image, L, K, hidden, B = "mixlogcdf", (32, 32, 3), 3, 1, 96, 8
    model, x, noise = _initialised_model(F, coupling, image, L, K, hidden, B, seed=77, num_blocks=2, drop_prob=0.0)
    sd64 = {k: v.detach().cpu().double().requires_grad_(v.dtype.is_floating_point and "is_initialized" not in k and
                                                         not k.endswith(".p") and not k.endswith("sign_s"))
            for k, v in model.state_dict().items()}
    sd_blocks = {k: v for k, v in sd64.items()}
    z, outs, ldj, nll = O.normal_flow(sd_blocks, x.double(), noise.double(), L, K, coupling)
    nll.mean().backward()
    model.train()                                              # the real training path (dropout layers present, p = 0)
    _lib.TIMING = {}
    try:
        _, nll_d, _ = model(x.to(dev()), noise=noise.to(dev()))
        nll_d.mean().backward()
        torch.cuda.synchronize()
        names = set(_lib.TIMING)
    finally:
        _lib.TIMING = None
    for need in ("flowk_conv_gemm", "flowk_conv_wgrad", "flowk_linear_wgrad", "flowk_attention_train_fwd",
                 "flowk_attention_train_bwd", "flowk_add_layernorm_fwd", "flowk_add_layernorm_bwd",
                 "flowk_weight_norm_operands_batched", "flowk_weight_norm_bwd_partials", "flowk_mixlogcdf_bwd"):
        assert need in names, (need, sorted(names))
    assert float((nll_d.detach().cpu() - nll.float()).abs().max()) < 1e-3
    checked = 0
    for name, p in model.named_parameters():
        ref = sd64[name].grad
        if ref is None:
            continue
        if name.endswith(".l") or name.endswith(".u"):
            c = ref.shape[0]
            ref = ref * (torch.tril(torch.ones(c, c), -1) if name.endswith(".l") else torch.triu(torch.ones(c, c), 1))
        parity(p.grad, ref, rel=1e-3, what="grad " + name)
        checked += 1
    assert checked > 100


def test_affine_training_gradients_on_tensor_cores(F, no_library):
    """Affine NN_net training (cfg1 width, 16x16 / 8x8 / 4x4): conv forward, input and weight gradients on the flowk
    tcgen05 kernels, ActNorm-in-conv through autograd; every parameter gradient vs float64 oracle autograd."""
    coupling, image, L, K, hidden, B = "affine", (32, 32, 3), 3, 1, 64, 8
    model, x, noise = _initialised_model(F, coupling, image, L, K, hidden, B, seed=78)
    sd64 = {k: v.detach().cpu().double().requires_grad_(v.dtype.is_floating_point and "is_initialized" not in k and
                                                         not k.endswith(".p") and not k.endswith("sign_s"))
            for k, v in model.state_dict().items()}
    z, outs, ldj, nll = O.normal_flow(sd64, x.double(), noise.double(), L, K, coupling)
    nll.mean().backward()
    model.train()
    _, nll_d, _ = model(x.to(dev()), noise=noise.to(dev()))
    nll_d.mean().backward()
    checked = 0
    for name, p in model.named_parameters():
        ref = sd64[name].grad
        if ref is None:
            continue
        if name.endswith(".l") or name.endswith(".u"):
            c = ref.shape[0]
            ref = ref * (torch.tril(torch.ones(c, c), -1) if name.endswith(".l") else torch.triu(torch.ones(c, c), 1))
        parity(p.grad, ref, rel=1e-3, what="grad " + name)
        checked += 1
    assert checked > 20


def test_eval_caches_follow_raw_pointer_updates(F, no_library):
    """FusedAdamax writes parameters through raw device pointers (no torch version bump): the folded / packed /
    weight-normed inference caches must still follow (weights generation, flowk._lib.bump_generation).  eval -> large
    fused steps -> eval must equal a fresh model loaded from the live state dict, and differ from the first eval."""
    from flowk.optim import FusedAdamax
    for coupling in ("mixlogcdf", "affine"):
        model, x, noise = _initialised_model(F, coupling, (16, 16, 3), 2, 2, 32, 4, seed=5, num_blocks=1)
        xd, nd = x.to(dev()), noise.to(dev())
        model.eval()
        with torch.no_grad():
            _, nll0, _ = model(xd, noise=nd)                    # fills every derived-weight cache
        opt = FusedAdamax(model.parameters(), lr=0.004)
        gen = torch.Generator().manual_seed(9)
        for _ in range(3):
            for p in model.parameters():
                p.grad = torch.randn(p.shape, generator=gen).to(dev())
            opt.step()
        with torch.no_grad():
            _, nll1, _ = model(xd, noise=nd)
        clone = F.marscf.MarScfFlow(4, (16, 16, 3), coupling, 2, 2, 32, num_blocks=1).to(dev())
        clone.load_state_dict(model.state_dict())
        clone.eval()
        with torch.no_grad():
            _, nll2, _ = clone(xd, noise=nd)
        assert torch.isfinite(nll1).all() and torch.isfinite(nll2).all()
        assert float((nll1 - nll0).abs().max()) > 1e-3, "the update must change the model output"
        assert float((nll1 - nll2).abs().max()) <= 1e-5 * max(1.0, float(nll2.abs().max())), coupling


@pytest.mark.parametrize("C,H,B", [(48, 4, 64), (96, 4, 40), (48, 8, 16)])
def test_channel_mix_tensor_core_route_matches_cuda_core_kernel(F, C, H, B):
    """The 1x1 invertible conv as a tcgen05 channel GEMM (the route large batches of wide levels take, ops.channel_mix)
    against the CUDA-core kernel and a float64 matmul."""
    from flowk import ops
    gen = torch.Generator().manual_seed(C + H)
    x = torch.randn(B, C, H, H, generator=gen).to(dev())
    mat = torch.linalg.qr(torch.randn(C, C, generator=gen))[0].contiguous().to(dev())
    bias = torch.randn(C, generator=gen).to(dev())
    ldj = torch.randn(B, generator=gen).to(dev())
    add = torch.full((1,), 3.5, device=dev())
    y_ref, l_ref = ops.channel_mix(x, mat, bias, ldj, add, False, False)
    old, old_c = ops.MIX_TC_MIN_ELEMENTS, ops.MIX_TC_MIN_CHANNELS
    ops.MIX_TC_MIN_ELEMENTS, ops.MIX_TC_MIN_CHANNELS = 0, 48
    try:
        from flowk import _lib
        _lib.TIMING = {}
        y_tc, l_tc = ops.channel_mix(x, mat, bias, ldj, add, False, False)
        torch.cuda.synchronize()
        assert "flowk_conv_gemm" in _lib.TIMING
    finally:
        ops.MIX_TC_MIN_ELEMENTS, ops.MIX_TC_MIN_CHANNELS = old, old_c
        _lib.TIMING = None
    ref64 = torch.einsum("oi,bihw->bohw", mat.double(), x.double()) + bias.double().view(1, -1, 1, 1)
    parity(y_tc, ref64.float().cpu(), rel=2e-6, what="tcgen05 channel mix vs fp64")
    parity(y_tc, y_ref.cpu(), rel=2e-6, what="tcgen05 vs CUDA-core channel mix")
    assert torch.allclose(l_tc, l_ref)


@pytest.mark.parametrize("C", [3, 12, 48, 96, 150])
def test_fold_kernel_matches_torch_assembly(F, C):
    """flowk_fold_actnorm_invconv (one launch: masks, triangular solves in fp64, P L U products, ActNorm fold, log-det) vs
    the same fold assembled with torch ops (common_modules.fold_actnorm_invconv with the kernel switched off)."""
    np.random.seed(C)
    gen = torch.Generator().manual_seed(C)
    an = F.cm.Actnormlayer(C).to(dev()).eval()
    ic = F.cm.InvertibleConv1x1(C).to(dev()).eval()
    with torch.no_grad():
        an.bias.copy_(torch.randn(1, C, 1, 1, generator=gen) * 0.3)
        an.logs.copy_(torch.randn(1, C, 1, 1, generator=gen) * 0.2)
        ic.l.add_(torch.randn(C, C, generator=gen).to(dev()) * 0.02)      # also off the triangle: must be masked out
        ic.u.add_(torch.randn(C, C, generator=gen).to(dev()) * 0.02)
        ic.log_s.add_(torch.randn(C, generator=gen).to(dev()) * 0.05)
        for reverse in (False, True):
            got = F.cm.fold_actnorm_invconv(an, ic, (8, 16), reverse)
            F.cm.FOLD_KERNEL = False
            try:
                want = F.cm.fold_actnorm_invconv(an, ic, (8, 16), reverse)
            finally:
                F.cm.FOLD_KERNEL = True
            for g, w, what in zip(got, want, ("matrix", "bias", "ldj")):
                scale = max(1.0, float(w.abs().max()))
                assert g.shape == w.shape, what
                assert float((g - w).abs().max()) <= 2e-6 * scale * (4 if reverse else 1), (what, reverse)


@pytest.mark.parametrize("B,C,H,W", [(64, 12, 16, 16), (5, 6, 3, 5), (1, 48, 4, 4), (3, 7, 1, 1)])
def test_standard_normal_prior_kernel_matches_torch_formula(F, B, C, H, W):
    """flowk_std_normal_logp (the default prior's term added to the objective in one launch) vs GaussianDiag.logp with zero
    mean / log-std (common_modules.py:223-240), on a channel slice and on a whole tensor, with and without a running logdet."""
    gen = torch.Generator().manual_seed(B * 100 + C)
    prior = F.marscf.StandardNormalPrior((32, 32, 3), 3)
    z = torch.randn(B, 2 * C, H, W, generator=gen).to(dev())
    ld = torch.randn(B, generator=gen).to(dev())
    for zz in (z[:, C:], z):
        want = F.cm.GaussianDiag.logp(torch.zeros_like(zz).double(), torch.zeros_like(zz).double(), zz.double())
        got = prior((z[:, :C], zz), 1)
        assert got.shape == (B,)
        assert float((got.double() - want).abs().max()) <= 2e-6 * float(want.abs().max())
        got2 = prior.accumulate(zz, 2, ld)
        assert float((got2.double() - (ld.double() + want)).abs().max()) <= 2e-6 * float(want.abs().max())
        assert torch.equal(prior.accumulate(zz, 2, ld), got2)          # fixed-order reduction
