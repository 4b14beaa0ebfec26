"""tcgen05 implicit-GEMM conv kernel against torch float64 convolutions."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tc():
    import flowk  # noqa: F401
    from flowk import tc as mod
    return mod


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


FORMATS = ["tf32", "f16"]


def run_conv(tc, x, w, bias, taps, out_mask=None, fmt="tf32", **kw):
    """x [B,Cin,H,W] (Cin % 32 == 0 for tf32 pairs, % 8 for fp16 pairs), w [N,Cin,k,k]; returns dict of outputs (the
    hi / lo operand outputs are returned summed in fp32 under "out_hilo", raw under "out_hi" / "out_lo")."""
    B, Cin, H, W = x.shape
    N = w.shape[0]
    rows = nhwc(x).reshape(B * H * W, Cin)
    odt = torch.float32
    if fmt == "f16":
        a_hi, a_lo = tc.split_rows_f16(rows)
        w_hi, w_lo, sc = tc.conv_weight_operand_f16(w)
        kw = dict(kw, acc_scale=sc)
        odt = torch.float16
    else:
        a_hi, a_lo = tc.split_hilo(rows)
        w_hi, w_lo = tc.conv_weight_operand(w)
    status = torch.zeros(1, dtype=torch.int32, device=x.device)
    outs = {"status": status}
    mask = out_mask if out_mask is not None else tc.OUT_F32
    pre = kw.pop("pre", tc.PRE_BIAS)
    nout = N // 2 if pre == tc.PRE_GLU_RES_LN else N
    if mask & tc.OUT_F32:
        outs["out_f32"] = torch.full((B * H * W, nout), float("nan"), device=x.device)
    if mask & (tc.OUT_HILO | tc.OUT_HILO_POS | tc.OUT_HILO_RELU):
        outs["out_hi"] = torch.full((B * H * W, nout), float("nan"), device=x.device, dtype=odt)
        outs["out_lo"] = torch.full((B * H * W, nout), float("nan"), device=x.device, dtype=odt)
    if mask & tc.OUT_HILO_CELU:
        outs["out_hi"] = torch.full((B * H * W, 2 * nout), float("nan"), device=x.device, dtype=odt)
        outs["out_lo"] = torch.full((B * H * W, 2 * nout), float("nan"), device=x.device, dtype=odt)
    if mask & tc.OUT_NCHW:
        outs["out_nchw"] = torch.full((B, N, H, W), float("nan"), device=x.device)
    tc.conv_gemm(a_hi, a_lo, w_hi, w_lo, B, H, W, Cin, N, taps, pre, mask, bias=bias, **kw, **outs)
    torch.cuda.synchronize()
    assert int(status) == 0, "barrier wait timed out inside the kernel"
    if "out_hi" in outs:
        outs["out_hilo"] = outs["out_hi"].float() + outs["out_lo"].float()
    return outs


def rel_err(got, ref):
    return float((got.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("B,Cin,H,W,N,taps", [(2, 32, 8, 16, 16, 1), (4, 64, 16, 16, 96, 1), (64, 96, 16, 16, 288, 1),
                                              (3, 32, 16, 16, 96, 9), (64, 192, 16, 16, 96, 9), (64, 96, 8, 8, 1176, 9),
                                              (19, 96, 4, 4, 2352, 9), (5, 32, 2, 2, 24, 9), (2, 64, 32, 32, 48, 9)])
def test_conv_gemm_matches_fp64(tc, fmt, B, Cin, H, W, N, taps):
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + N)
    k = 3 if taps == 9 else 1
    x = torch.randn(B, Cin, H, W, generator=g).to(dev)
    w = (torch.randn(N, Cin, k, k, generator=g) / (Cin * taps) ** 0.5).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=k // 2)
    outs = run_conv(tc, x, w, bias, taps, out_mask=tc.OUT_F32 | tc.OUT_NCHW | tc.OUT_HILO_CELU, fmt=fmt)
    got = outs["out_f32"].view(B, H, W, N).permute(0, 3, 1, 2)
    torch.backends.cudnn.allow_tf32 = False
    lib32 = rel_err(F.conv2d(x, w, bias, padding=k // 2), ref)        # what the fp32 library conv achieves
    print("K=%d  tcgen05 %s-pair rel err %.2e   cuDNN fp32 rel err %.2e" % (Cin * taps, fmt, rel_err(got, ref), lib32))
    assert rel_err(got, ref) < max(6e-6, 4 * lib32), (rel_err(got, ref), lib32)
    assert rel_err(outs["out_nchw"], ref) < max(6e-6, 4 * lib32)
    celu = F.elu(torch.cat((nhwc(ref), -nhwc(ref)), dim=-1)).reshape(B * H * W, 2 * N)
    assert rel_err(outs["out_hilo"], celu) < max(6e-6, 4 * lib32)
    if fmt == "tf32":
        assert int((outs["out_hi"].view(torch.int32) & 8191).abs().max()) == 0      # hi is exactly TF32-representable


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("B,C,H,W", [(64, 96, 16, 16), (8, 96, 4, 4), (3, 32, 8, 8), (2, 160, 16, 16)])
def test_glu_residual_layernorm_epilogue(tc, fmt, B, C, H, W):
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(B + C)
    x = torch.randn(B, C, H, W, generator=g).to(dev)
    w = (torch.randn(2 * C, C, 1, 1, generator=g) / C ** 0.5).to(dev)
    bias = torch.randn(2 * C, generator=g).to(dev)
    res = torch.randn(B * H * W, C, generator=g).to(dev)
    gamma = (torch.rand(C, generator=g) + 0.5).to(dev)
    beta = torch.randn(C, generator=g).to(dev)
    pos = torch.randn(H * W, C, generator=g).to(dev)
    y = nhwc(F.conv2d(x.double(), w.double(), bias.double())).reshape(B * H * W, 2 * C)
    glu = y[:, :C] * torch.sigmoid(y[:, C:]) + res.double()
    ref = F.layer_norm(glu, (C,), gamma.double(), beta.double())
    outs = run_conv(tc, x, w, bias, 1, out_mask=tc.OUT_F32 | tc.OUT_HILO_POS, pre=tc.PRE_GLU_RES_LN, res=res,
                    gamma=gamma, beta=beta, pos=pos, fmt=fmt)
    assert rel_err(outs["out_f32"], ref) < 1e-5, rel_err(outs["out_f32"], ref)
    ref_pos = ref + pos.double().repeat(B, 1)
    assert rel_err(outs["out_hilo"], ref_pos) < 1e-5


@pytest.mark.parametrize("c,C,H,W,B,blocks", [(6, 96, 16, 16, 64, 2), (12, 96, 8, 8, 64, 2), (24, 96, 4, 4, 64, 2),
                                              (6, 32, 16, 16, 3, 1), (6, 64, 32, 32, 2, 1)])
def test_conditioner_tc_matches_torch_path(tc, c, C, H, W, B, blocks):
    """NN.forward_raw through the tcgen05 chain vs the torch/cuDNN fp32 layers (same module, same weights)."""
    from flowk import conditioner_tc
    from flowk.flow_modules.mixlogcdf_nn import NN
    torch.manual_seed(c + C)
    dev = torch.device("cuda:0")
    net = NN(c, C, blocks, 32, 0.2).to(dev).eval()
    with torch.no_grad():
        for p in net.parameters():
            p.add_(torch.randn_like(p) * 0.02)
        x = torch.randn(B, 2 * c, H, W, device=dev)
        x_id = x[:, c:]
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        got = conditioner_tc.mixlogcdf_nn_raw(net, x_id, status=status)
        conditioner_tc.ENABLED = False
        try:
            ref = net.forward_raw(x_id)
            ref64 = net.double().forward_raw(x_id.double())
        finally:
            conditioner_tc.ENABLED = True
            net.float()
    assert int(status) == 0
    e_tc, e_lib = rel_err(got, ref64), rel_err(ref, ref64)
    print("conditioner c=%d C=%d %dx%d: tcgen05 rel err %.2e, torch fp32 rel err %.2e" % (c, C, H, W, e_tc, e_lib))
    assert e_tc < max(2e-5, 4 * e_lib)


@pytest.mark.parametrize("B,HW,C,heads", [(64, 256, 96, 4), (64, 64, 96, 4), (64, 16, 96, 4), (3, 1024, 64, 4),
                                          (5, 4, 32, 4), (2, 256, 160, 4)])
@pytest.mark.parametrize("fmt", FORMATS)
def test_attention_matches_fp64(tc, fmt, B, HW, C, heads):
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(B + HW)
    qkv = torch.randn(B * HW, 3 * C, generator=g).to(dev)
    hi, lo = tc.attention(qkv, B, HW, C, heads, fmt == "f16")
    if fmt == "f16":
        assert hi.dtype == torch.float16
        hi, lo = hi.float(), lo.float()
    d = C // heads
    t = qkv.double().view(B, HW, 3, heads, d)
    k, v, q = (t[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    ref = (torch.softmax((q * d ** -0.5) @ k.transpose(-1, -2), dim=-1) @ v).permute(0, 2, 1, 3).reshape(B * HW, C)
    assert rel_err(hi + lo, ref) < 2e-5      # fp32-accumulate order over up to 1024 keys; the conditioner budget is 1e-4
    if fmt == "tf32":
        assert int((hi.view(torch.int32) & 8191).abs().max()) == 0


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("B,HW,C,heads", [(64, 256, 96, 4), (3, 128, 64, 4), (2, 256, 160, 4), (5, 256, 32, 4),
                                          (2, 128, 256, 4), (1, 128, 64, 1), (3, 256, 48, 1)])
def test_attention_tcgen05_matches_fp64_and_mma_sync(tc, fmt, B, HW, C, heads):
    """csrc/attention_tc.cu (S = Q K^T and O = P V on tcgen05, softmax from TMEM) against fp64 and against the mma.sync
    kernel it replaces for seq in {128, 256}; head dims 8, 16, 24, 40, 48, 64 (24 and 40 exercise the zero-padded K = 16 step;
    seq 256 with head dim 64 exceeds the shared memory and stays on the mma.sync kernel)."""
    from flowk import _lib
    dev = torch.device("cuda:0")
    assert tc.attention_tc_supported(HW, C, heads)
    g = torch.Generator(device="cpu").manual_seed(B * HW + C)
    qkv = (torch.randn(B * HW, 3 * C, generator=g) * 1.5).to(dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    f16 = fmt == "f16"
    hi, lo = tc.attention(qkv, B, HW, C, heads, f16, status=status)
    torch.cuda.synchronize()
    assert int(status) == 0, "barrier wait timed out inside the kernel"
    got = hi.float() + lo.float()
    d = C // heads
    t = qkv.double().view(B, HW, 3, heads, d)
    k, v, q = (t[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    ref = (torch.softmax((q * d ** -0.5) @ k.transpose(-1, -2), dim=-1) @ v).permute(0, 2, 1, 3).reshape(B * HW, C)
    assert rel_err(got, ref) < 2e-5, rel_err(got, ref)
    if d in (8, 16, 24, 32, 40, 64):
        old_hi = torch.empty_like(hi)
        old_lo = torch.empty_like(lo)
        _lib.call("flowk_attention_f16" if f16 else "flowk_attention", qkv.data_ptr(), old_hi.data_ptr(), old_lo.data_ptr(),
                  B, HW, C, heads, tc._stream())
        assert rel_err(got, (old_hi.float() + old_lo.float()).double()) < 2e-5


@pytest.mark.parametrize("c,hidden,H,W,B", [(6, 64, 16, 16, 32), (12, 256, 8, 8, 128), (24, 256, 4, 4, 128),
                                            (6, 256, 32, 32, 4), (48, 256, 4, 4, 16)])
def test_affine_conditioner_tc_matches_torch_path(tc, c, hidden, H, W, B):
    from flowk import conditioner_tc
    from flowk.flow_modules.affine_coupling import NN_net
    torch.manual_seed(c + hidden)
    dev = torch.device("cuda:0")
    net = NN_net(c, 2 * c, hidden).to(dev).eval()
    with torch.no_grad():
        for p in net.parameters():
            p.add_(torch.randn_like(p) * 0.05)
        x = torch.randn(B, 2 * c, H, W, device=dev)
        z1 = x[:, :c]
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        got = conditioner_tc.affine_nn_net(net, z1, status=status)
        conditioner_tc.ENABLED = False
        try:
            ref = net(z1)
            ref64 = net.double()(z1.double())
        finally:
            conditioner_tc.ENABLED = True
            net.float()
    assert int(status) == 0
    e_tc, e_lib = rel_err(got, ref64), rel_err(ref, ref64)
    print("NN_net c=%d hidden=%d %dx%d: tcgen05 rel err %.2e, torch fp32 rel err %.2e" % (c, hidden, H, W, e_tc, e_lib))
    assert e_tc < max(3e-5, 4 * e_lib)       # K = 9*256: tensor-core fp32 accumulation order


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("B,C,H,W", [(64, 96, 16, 16), (64, 96, 4, 4), (3, 32, 8, 8), (5, 64, 4, 4)])
def test_chained_gate_in_proj(tc, fmt, B, C, H, W):
    """GLU+residual+LayerNorm GEMM with the in_proj GEMM chained inside the same CTA == the two separate launches (both
    operand formats; with fp16 pairs C = 96 / 32 leave the second GEMM's last 64-channel k-block partly zero-padded)."""
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(B * 7 + C)
    M = B * H * W
    f16 = fmt == "f16"
    a = torch.randn(M, 2 * C, generator=g).to(dev)
    w = (torch.randn(2 * C, 2 * C, generator=g) / (2 * C) ** 0.5).to(dev)
    w2 = (torch.randn(3 * C, C, generator=g) / C ** 0.5).to(dev)
    kw = {}
    if f16:
        a_hi, a_lo = tc.split_rows_f16(a)
        w_hi, w_lo, sc = tc.conv_weight_operand_f16(w)
        w2_hi, w2_lo, sc2 = tc.conv_weight_operand_f16(w2)
        kw = dict(acc_scale=sc, acc_scale2=sc2)
    else:
        a_hi, a_lo = tc.split_hilo(a)
        w_hi, w_lo = tc.split_hilo(w)
        w2_hi, w2_lo = tc.split_hilo(w2)
    bias = torch.randn(2 * C, generator=g).to(dev)
    res = torch.randn(M, C, generator=g).to(dev)
    gamma = (torch.rand(C, generator=g) + 0.5).to(dev)
    beta = torch.randn(C, generator=g).to(dev)
    pos = torch.randn(H * W, C, generator=g).to(dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    assert tc.chain_supported(C, 3 * C, f16)
    x1 = torch.empty(M, C, device=dev)
    qkv = torch.full((M, 3 * C), float("nan"), device=dev)
    tc.conv_gemm(a_hi, a_lo, w_hi, w_lo, B, H, W, 2 * C, 2 * C, 1, tc.PRE_GLU_RES_LN, tc.OUT_F32, bias=bias, res=res,
                 gamma=gamma, beta=beta, pos=pos, out_f32=x1, status=status, w2_hi=w2_hi, w2_lo=w2_lo, out2_f32=qkv, n2=3 * C,
                 **kw)
    torch.cuda.synchronize()
    assert int(status) == 0
    y = a.double() @ w.double().t() + bias.double()
    ln = F.layer_norm(y[:, :C] * torch.sigmoid(y[:, C:]) + res.double(), (C,), gamma.double(), beta.double())
    ref = (ln + pos.double().repeat(B, 1)) @ w2.double().t()
    assert rel_err(x1, ln) < 1e-5
    assert rel_err(qkv, ref) < 1e-5, rel_err(qkv, ref)


def test_wide_linear_many_rows(tc):
    """N = 480 (in_proj of a 160-channel conditioner) with more M tiles than SMs: must take the two-N-tile path, the
    single-CTA two-chunk variant would not fit its epilogue slab in shared memory."""
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn(64, 160, 16, 16, generator=g).to(dev)
    w = (torch.randn(480, 160, 1, 1, generator=g) / 160 ** 0.5).to(dev)
    outs = run_conv(tc, x, w, None, 1, out_mask=tc.OUT_F32)
    ref = nhwc(F.conv2d(x.double(), w.double())).reshape(-1, 480)
    assert rel_err(outs["out_f32"], ref) < 1e-5


@pytest.mark.parametrize("fmt", FORMATS)
def test_conv_gemm_random_configuration_sweep(tc, fmt):
    """Host-side tiling decisions (N tiles / chunks, dx-split, stage counts, epilogue staging) over a seeded sweep of
    layer shapes; every configuration must either be refused with a shape error or match the fp64 convolution."""
    import random
    rnd = random.Random(1234)
    dev = torch.device("cuda:0")
    ran = 0
    for trial in range(28):
        H, W = rnd.choice([(2, 2), (4, 4), (8, 8), (16, 16), (32, 32), (8, 16), (4, 32)])
        B = rnd.choice([1, 3, 16, 64, 130]) if H * W <= 256 else rnd.choice([1, 3, 9])
        Cin = rnd.choice([32, 64, 96, 160, 192, 256] + ([8, 24, 40] if fmt == "f16" else []))
        taps = rnd.choice([1, 9])
        pre = rnd.choice([tc.PRE_BIAS, tc.PRE_BIAS, tc.PRE_GLU_RES_LN])
        if pre == tc.PRE_GLU_RES_LN:
            C = rnd.choice([32, 64, 96, 160, 256])
            N = 2 * C
            mask = rnd.choice([tc.OUT_F32, tc.OUT_F32 | tc.OUT_HILO_CELU, tc.OUT_HILO])
        else:
            N = 16 * rnd.randint(1, 40)
            mask = rnd.choice([tc.OUT_F32, tc.OUT_NCHW, tc.OUT_F32 | tc.OUT_HILO_RELU, tc.OUT_HILO_CELU])
        g = torch.Generator(device="cpu").manual_seed(trial)
        k = 3 if taps == 9 else 1
        x = torch.randn(B, Cin, H, W, generator=g).to(dev)
        w = (torch.randn(N, Cin, k, k, generator=g) / (Cin * taps) ** 0.5).to(dev)
        bias = torch.randn(N, generator=g).to(dev)
        kw = {}
        if pre == tc.PRE_GLU_RES_LN:
            kw = dict(pre=pre, res=torch.randn(B * H * W, N // 2, generator=g).to(dev),
                      gamma=(torch.rand(N // 2, generator=g) + 0.5).to(dev), beta=torch.randn(N // 2, generator=g).to(dev))
        try:
            outs = run_conv(tc, x, w, bias, taps, out_mask=mask, fmt=fmt, **kw)
        except AssertionError as e:
            assert "bad shape" in str(e), e          # an honest refusal is fine; anything else is a bug
            continue
        ran += 1
        y = nhwc(F.conv2d(x.double(), w.double(), bias.double(), padding=k // 2)).reshape(B * H * W, N)
        if pre == tc.PRE_GLU_RES_LN:
            C = N // 2
            y = F.layer_norm(y[:, :C] * torch.sigmoid(y[:, C:]) + kw["res"].double(), (C,), kw["gamma"].double(),
                             kw["beta"].double())
        tol = 3e-5
        desc = (trial, B, H, W, Cin, N, taps, pre, mask)
        if mask & tc.OUT_F32:
            assert rel_err(outs["out_f32"], y) < tol, desc
        if mask & tc.OUT_NCHW:
            assert rel_err(outs["out_nchw"], y.view(B, H, W, N).permute(0, 3, 1, 2)) < tol, desc
        if mask & tc.OUT_HILO:
            assert rel_err(outs["out_hilo"], y) < tol, desc
        if mask & tc.OUT_HILO_RELU:
            assert rel_err(outs["out_hilo"], torch.relu(y)) < tol, desc
        if mask & tc.OUT_HILO_CELU:
            assert rel_err(outs["out_hilo"], F.elu(torch.cat((y, -y), dim=-1))) < tol, desc
    assert ran >= 20, ran


def test_training_conv_and_linear_on_tcgen05_match_torch_autograd(tc):
    """tc_autograd.conv2d / linear (forward and input gradient on the tcgen05 kernel) against torch's float64 autograd."""
    from flowk import tc_autograd
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(21)
    for (B, Cin, H, W, N, k) in ((8, 192, 16, 16, 96, 3), (8, 12, 16, 16, 96, 3), (16, 96, 8, 8, 588, 3), (8, 192, 16, 16, 192, 1)):
        x = torch.randn(B, Cin, H, W, generator=g).to(dev).requires_grad_()
        w = (torch.randn(N, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev).requires_grad_()
        b = torch.randn(N, generator=g).to(dev).requires_grad_()
        gy = torch.randn(B, N, H, W, generator=g).to(dev)
        assert tc_autograd.conv_supported(x, w)
        y = tc_autograd.conv2d(x, w, b, k // 2)
        y.backward(gy)
        x64, w64, b64 = (t.detach().double().requires_grad_() for t in (x, w, b))
        y64 = F.conv2d(x64, w64, b64, padding=k // 2)
        y64.backward(gy.double())
        assert rel_err(y, y64.detach()) < 2e-5
        assert rel_err(x.grad, x64.grad) < 6e-5          # dgrad K = 9 * Cout, tensor-core fp32 accumulation order
        assert rel_err(w.grad, w64.grad) < 2e-5
        assert rel_err(b.grad, b64.grad) < 2e-5
    x = torch.randn(4, 16, 16, 96, generator=g).to(dev).requires_grad_()
    w = (torch.randn(288, 96, generator=g) / 10).to(dev).requires_grad_()
    gy = torch.randn(4, 16, 16, 288, generator=g).to(dev)
    assert tc_autograd.linear_supported(x, w)
    y = tc_autograd.linear(x, w, None)
    y.backward(gy)
    x64, w64 = x.detach().double().requires_grad_(), w.detach().double().requires_grad_()
    y64 = F.linear(x64, w64)
    y64.backward(gy.double())
    assert rel_err(y, y64.detach()) < 1e-5 and rel_err(x.grad, x64.grad) < 1e-5 and rel_err(w.grad, w64.grad) < 2e-5


def test_fused_pointwise_layers_match_torch_autograd():
    from flowk import tc_autograd
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(9)
    for shape, dim in (((3, 8, 4, 5), 1), ((2, 4, 4, 12), -1), ((64, 96, 16, 16), 1)):
        x = torch.randn(shape, generator=g).to(dev).requires_grad_()
        mask = None
        if dim == 1:                                    # feature dropout folded in (GatedConv's Dropout2d)
            mask = tc_autograd.feature_dropout_mask(x, 2 * x.shape[1], 0.3)
            assert all(v == 0.0 or abs(v - 1.0 / 0.7) < 1e-6 for v in mask.unique().tolist())
        y = tc_autograd.concat_elu(x, dim, mask)
        gy = torch.randn(y.shape, generator=g).to(dev)
        y.backward(gy)
        x64 = x.detach().double().requires_grad_()
        y64 = F.elu(torch.cat((x64, -x64), dim=dim))
        if mask is not None:
            y64 = y64 * mask.double().view(mask.shape[0], mask.shape[1], 1, 1)
        y64.backward(gy.double())
        assert rel_err(y, y64.detach()) < 1e-6 and rel_err(x.grad, x64.grad) < 1e-6
        x2 = torch.randn(y.shape, generator=g).to(dev).requires_grad_()
        z = tc_autograd.glu(x2, dim)
        gz = torch.randn(z.shape, generator=g).to(dev)
        z.backward(gz)
        x264 = x2.detach().double().requires_grad_()
        a, b = x264.chunk(2, dim=dim)
        z64 = a * torch.sigmoid(b)
        z64.backward(gz.double())
        assert rel_err(z, z64.detach()) < 1e-6 and rel_err(x2.grad, x264.grad) < 1e-6


def test_weight_norm_conv_and_linear_autograd_match_fp64(allow_library):     # Linear with N = 100: no tcgen05 wgrad tile
    from flowk import tc_autograd
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(21)

    def wn64(v, gg):
        nrm = v.reshape(v.shape[0], -1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
        return v * (gg / nrm)

    for (b, cin, n, k, hw) in ((4, 32, 24, 3, 16), (2, 40, 64, 1, 16), (8, 16, 588, 3, 8)):
        x = torch.randn(b, cin, hw, hw, generator=g).to(dev).requires_grad_()
        v = (0.2 * torch.randn(n, cin, k, k, generator=g)).to(dev).requires_grad_()
        gg = (1 + 0.1 * torch.randn(n, 1, 1, 1, generator=g)).to(dev).requires_grad_()
        bias = torch.randn(n, generator=g).to(dev).requires_grad_()
        y = tc_autograd.wn_conv2d(x, v, gg, bias)
        gy = torch.randn(y.shape, generator=g).to(dev)
        y.backward(gy)
        x6, v6, g6, b6 = (t.detach().double().requires_grad_() for t in (x, v, gg, bias))
        y6 = F.conv2d(x6, wn64(v6, g6), b6, padding=k // 2)
        y6.backward(gy.double())
        assert rel_err(y, y6.detach()) < 2e-5
        for a, r in ((x.grad, x6.grad), (v.grad, v6.grad), (gg.grad, g6.grad), (bias.grad, b6.grad)):
            assert rel_err(a, r) < 1e-4, rel_err(a, r)
    for (m, kdim, n) in ((256, 96, 288), (128, 64, 100)):
        x = torch.randn(2, m // 2, kdim, generator=g).to(dev).requires_grad_()
        v = (0.2 * torch.randn(n, kdim, generator=g)).to(dev).requires_grad_()
        gg = (1 + 0.1 * torch.randn(n, 1, generator=g)).to(dev).requires_grad_()
        y = tc_autograd.wn_linear(x, v, gg, None)
        gy = torch.randn(y.shape, generator=g).to(dev)
        y.backward(gy)
        x6, v6, g6 = (t.detach().double().requires_grad_() for t in (x, v, gg))
        y6 = F.linear(x6, wn64(v6, g6))
        y6.backward(gy.double())
        assert rel_err(y, y6.detach()) < 2e-5
        for a, r in ((x.grad, x6.grad), (v.grad, v6.grad), (gg.grad, g6.grad)):
            assert rel_err(a, r) < 1e-4, rel_err(a, r)


@pytest.mark.parametrize("in_nchw,out_nchw", [(True, False), (False, True), (True, True), (False, False)])
def test_add_layernorm_matches_torch_autograd(in_nchw, out_nchw):
    from flowk import tc_autograd
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(33)
    for (b, c, h, w) in ((3, 96, 4, 4), (2, 24, 8, 16), (5, 160, 4, 4)):
        norm = torch.nn.LayerNorm(c).to(dev)
        with torch.no_grad():
            norm.weight.copy_(1 + 0.2 * torch.randn(c, generator=g))
            norm.bias.copy_(0.2 * torch.randn(c, generator=g))
        shape = (b, c, h, w) if in_nchw else (b, h, w, c)
        a = torch.randn(shape, generator=g).to(dev).requires_grad_()
        r = torch.randn(shape, generator=g).to(dev).requires_grad_()
        y = tc_autograd.add_layernorm(a, r, norm, in_nchw, out_nchw)
        gy = torch.randn(y.shape, generator=g).to(dev)
        y.backward(gy)
        got = (y.detach(), a.grad, r.grad, norm.weight.grad.clone(), norm.bias.grad.clone())
        a6, r6 = a.detach().double().requires_grad_(), r.detach().double().requires_grad_()
        n6 = torch.nn.LayerNorm(c).to(dev).double()
        n6.load_state_dict({k: v.double() for k, v in norm.state_dict().items()})
        s6 = a6 + r6
        y6 = n6(s6.permute(0, 2, 3, 1) if in_nchw else s6)
        if out_nchw:
            y6 = y6.permute(0, 3, 1, 2)
        y6.backward(gy.double())
        ref = (y6.detach(), a6.grad, r6.grad, n6.weight.grad, n6.bias.grad)
        for u, v in zip(got, ref):
            assert u.shape == v.shape and rel_err(u, v) < 5e-6, rel_err(u, v)


@pytest.mark.parametrize("b,cin,n,k,h,w", [(2, 32, 32, 1, 8, 16), (4, 24, 40, 3, 16, 16), (8, 192, 96, 3, 8, 8),
                                           (3, 96, 588, 3, 8, 8), (2, 320, 160, 3, 8, 8), (1, 96, 288, 1, 64, 32),
                                           (64, 192, 96, 3, 4, 4), (6, 96, 2352, 3, 4, 4), (8, 192, 192, 1, 4, 4)])
def test_conv_wgrad_tcgen05_matches_fp64(b, cin, n, k, h, w):
    from flowk import tc_autograd
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn(b, cin, h, w, generator=g).to(dev)
    gy = torch.randn(b, n, h, w, generator=g).to(dev)
    out = tc_autograd.wgrad_partials(x, gy, k * k)
    assert out is not None
    partial, transposed = out
    got = (partial.sum(0).permute(2, 1, 0) if transposed else partial.sum(0).permute(1, 2, 0)).reshape(n, cin, k, k)
    ref = torch.nn.grad.conv2d_weight(x.double(), (n, cin, k, k), gy.double(), padding=k // 2)
    assert rel_err(got, ref) < 2e-5
    again, _ = tc_autograd.wgrad_partials(x, gy, k * k)
    assert torch.equal(again, partial)                       # split-K partials are deterministic


@pytest.mark.parametrize("B,Cin,H,W,N,taps,nchw", [(64, 192, 4, 4, 96, 9, True), (16, 2368, 4, 4, 96, 9, True),
                                                   (8, 192, 8, 8, 192, 9, True), (4, 1536, 8, 16, 96, 1, False)])
def test_conv_gemm_split_k_matches_fp64_and_unsplit(B, Cin, H, W, N, taps, nchw):
    """Split-K (training path): few 128-row tiles, long K loop -> several CTAs per output tile + ordered reduce."""
    from flowk import _lib, tc
    import ctypes
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(17)
    k = 3 if taps == 9 else 1
    x = torch.randn(B, Cin, H, W, generator=g).to(dev)
    w = (torch.randn(N, Cin, k, k, generator=g) / (Cin * taps) ** 0.5).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    a_hi, a_lo = tc.nchw_to_nhwc_hilo(x, Cin)
    w_hi, w_lo = tc.conv_weight_operand(w, Cin)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=k // 2)
    outs = []
    for split in (False, True):
        if nchw:
            y = torch.empty(B, N, H, W, device=dev)
            tc.conv_gemm(a_hi, a_lo, w_hi, w_lo, B, H, W, Cin, N, taps, tc.PRE_BIAS, tc.OUT_NCHW, bias=bias, out_nchw=y,
                         split_k=split)
        else:
            rows = torch.empty(B * H * W, N, device=dev)
            tc.conv_gemm(a_hi, a_lo, w_hi, w_lo, B, H, W, Cin, N, taps, tc.PRE_BIAS, tc.OUT_F32, bias=bias, out_f32=rows,
                         split_k=split)
            y = rows.view(B, H, W, N).permute(0, 3, 1, 2)
        outs.append(y)
        # fp32 accumulation in TMEM over K/8 MMA steps: the error grows with K (K = 21 312 for the out_conv dgrad)
        assert rel_err(y, ref) < (2e-5 if Cin * taps < 5000 else 1.5e-4), (split, rel_err(y, ref))
    args = _lib.ConvGemmArgs(None, None, None, None, None, None, None, None, None, None, None, None, None, None, None,
                             B, H, W, Cin, N, taps, tc.PRE_BIAS, tc.OUT_NCHW if nchw else tc.OUT_F32, None, None, None, 0, None)
    assert _lib.lib.flowk_conv_gemm_splitk_slices(ctypes.addressof(args)) > 1        # these shapes do split
    assert rel_err(outs[1], ref) <= 1.5 * rel_err(outs[0], ref) + 1e-6           # shorter chains: split-K is no less accurate


@pytest.mark.parametrize("m,k,n", [(128, 32, 32), (256, 96, 288), (16384, 96, 192), (1024, 96, 288), (512, 160, 480)])
def test_linear_wgrad_tcgen05_mn_major_matches_fp64(m, k, n):
    """Row-major operands as MN-major (transposed) tf32 tiles: 32-byte-atom 128-byte swizzle on both the TMA and UMMA side."""
    from flowk import tc_autograd
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(8)
    x = torch.randn(m, k, generator=g).to(dev)
    gy = torch.randn(m, n, generator=g).to(dev)
    out = tc_autograd.linear_wgrad_partials(x, gy)
    assert out is not None
    partial, transposed = out
    got = partial.sum(0)[0]
    got = got.t() if transposed else got
    ref = gy.double().t() @ x.double()
    assert got.shape == ref.shape and rel_err(got, ref) < 2e-5


@pytest.mark.parametrize("fwd_f16", [False, True])
def test_weight_norm_batch_equals_per_layer_path(fwd_f16, monkeypatch):
    """One training step with the two-launch WeightNormBatch == the per-layer weight-norm path.  With TF32 forward operands
    in both the forward is bit-identical; with the batch's fp16 (hi, lo) forward operands (tc_autograd.FWD_F16, the default)
    the two agree to the operand split's 2^-22."""
    from flowk import tc_autograd
    from flowk.marscf import MarScfFlow
    monkeypatch.setattr(tc_autograd, "FWD_F16", fwd_f16)
    dev = torch.device("cuda:0")
    torch.manual_seed(4)
    model = MarScfFlow(4, (16, 16, 3), "mixlogcdf", 2, 1, 32, num_blocks=1).to(dev)
    x = torch.rand(4, 3, 16, 16, device=dev) - 0.5
    noise = torch.rand(4, 3, 16, 16, device=dev)
    model.train()
    with torch.no_grad():
        model(x, noise=noise)
    for m in model.modules():                           # dropout off so both runs are comparable
        if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout2d)):
            m.p = 0.0
        if hasattr(m, "drop_prob"):
            m.drop_prob = 0.0

    def grads(use_batch):
        model.zero_grad(set_to_none=True)
        if use_batch:
            _, nll, _ = model(x, noise=noise)           # MarScfFlow.normal_flow refreshes the batch
            assert model._wn_batch is not None and len(model._wn_batch.modules) > 4
        else:
            d = x[0].numel()
            z = x + noise / 256.0
            ld = x.new_full((4,), float(-math.log(256.0) * d))
            _, obj = model.flow(z, logdet=ld, reverse=False)   # FlowNet directly: no refresh -> per-layer path
            nll = -obj / (math.log(2.0) * d)
        nll.mean().backward()
        return nll.detach().clone(), [p.grad.clone() for p in model.parameters()]

    nll_a, ga = grads(True)
    nll_b, gb = grads(False)
    if fwd_f16:
        assert float((nll_a - nll_b).abs().max()) <= 2e-6 * float(nll_b.abs().max())
    else:
        assert torch.equal(nll_a, nll_b)                # same operands -> the forward is bit-identical
    tol = 1e-4 if fwd_f16 else 1e-5                     # (a few torch backward ops use atomics: compare to fp32 round-off)
    for a, b in zip(ga, gb):
        assert float((a - b).abs().max()) <= tol * float(b.abs().max()) + 1e-9


@pytest.mark.parametrize("B,S,C,heads,p", [(2, 16, 32, 4, 0.0), (3, 64, 96, 4, 0.2), (2, 256, 96, 4, 0.2), (1, 136, 160, 4, 0.3),
                                           (2, 32, 64, 8, 0.5)])
def test_training_attention_fwd_bwd_match_fp64(B, S, C, heads, p):
    """dropout(softmax(q k^T / sqrt d)) v and its gradients against fp64 autograd using the kernels' own dropout mask."""
    from flowk import tc_autograd
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(23)
    qkv = torch.randn(B, S, 3 * C, generator=g).to(dev).requires_grad_()
    dout = torch.randn(B, S, C, generator=g).to(dev)
    salt, d = 7, C // heads
    out = tc_autograd.attention_core(qkv, heads, p, salt)
    out.backward(dout)
    mask = tc_autograd.attention_dropout_mask(dev, salt, p, B * heads, S).double().view(B, heads, S, S)
    if p > 0:
        keep = float((mask > 0).double().mean())
        assert abs(keep - (1 - p)) < 0.02 and abs(float(mask.max()) - 1 / (1 - p)) < 1e-6
    else:
        assert torch.all(mask == 1)
    x = qkv.detach().double().requires_grad_()
    k, v, q = x[..., :C], x[..., C:2 * C], x[..., 2 * C:]
    hf = lambda m: m.reshape(B, S, heads, d).permute(0, 2, 1, 3)          # noqa: E731
    w = torch.softmax((hf(q) * d ** -0.5) @ hf(k).transpose(-1, -2), dim=-1) * mask
    ref = (w @ hf(v)).permute(0, 2, 1, 3).reshape(B, S, C)
    ref.backward(dout.double())
    assert rel_err(out.detach(), ref.detach()) < 2e-5
    assert rel_err(qkv.grad, x.grad) < 5e-5
    out2 = tc_autograd.attention_core(qkv.detach(), heads, p, salt)       # same seed -> same mask -> bit-identical
    assert torch.equal(out2, out.detach())
    tc_autograd.advance_dropout_seed(dev)
    if p > 0:
        assert not torch.equal(tc_autograd.attention_core(qkv.detach(), heads, p, salt), out.detach())


@pytest.mark.parametrize("shape", [(96, 192, 3), (192, 96, 1), (12, 6, 3), (130, 40, 5), (288, 96, None)])
@pytest.mark.parametrize("mode", ["plain", "weight_norm", "exp_gain"])
def test_pack_weight_f16_matches_torch_formula(tc, shape, mode):
    """flowk_pack_weight_f16 (weight norm / ActNorm gain, permute, pad, power-of-two scaling, hi/lo split in two launches)
    against the same operand written with torch ops: hi + lo reproduces w * gain * 2^e to 2^-22 relative, the padding is
    zero, and acc_scale is the exact inverse power of two with max|w 2^e| in [2^14, 2^15)."""
    from flowk import _lib
    dev = torch.device("cuda:0")
    n, cin, k = shape
    torch.manual_seed(n + cin)
    w = torch.randn((n, cin) if k is None else (n, cin, k, k), device=dev) * 0.07
    taps = 1 if k is None else k * k
    gain = bias = None
    if mode == "plain":
        hi, lo, sc = tc.conv_weight_operand_f16(w)
        w_eff = w
    elif mode == "weight_norm":
        gain = torch.rand(n, 1, device=dev) + 0.5
        hi, lo, sc = tc.conv_weight_operand_f16(w, gain, _lib.PACK_WEIGHT_NORM)
        norm = w.reshape(n, -1).double().norm(dim=1)
        w_eff = (w.double().reshape(n, -1) * (gain.double().reshape(-1) / norm)[:, None]).reshape(w.shape)
    else:
        gain = torch.randn(1, n, 1, 1, device=dev) * 0.3
        bias = torch.randn(n, device=dev)
        hi, lo, sc, b = tc.conv_weight_operand_f16(w, gain, _lib.PACK_EXP_GAIN, 3.0, bias)
        g = torch.exp(gain.double().reshape(-1) * 3.0)
        w_eff = (w.double().reshape(n, -1) * g[:, None]).reshape(w.shape)
        torch.testing.assert_close(b.double(), bias.double() * g, rtol=1e-6, atol=0)
    cin_pad = (cin + 63) // 64 * 64
    assert hi.shape == lo.shape == (n, taps * cin_pad) and hi.dtype == torch.float16
    m, e = math.frexp(1.0 / sc)
    assert m == 0.5                                               # a power of two
    want = w_eff.double().reshape(n, cin, taps).permute(0, 2, 1) / sc          # [n, taps, cin] scaled
    amax = float(want.abs().max())
    assert 2.0 ** 14 * (1 - 1e-6) <= amax < 2.0 ** 15 * (1 + 1e-6)
    got = (hi.double() + lo.double()).reshape(n, taps, cin_pad)
    assert float(got[:, :, cin:].abs().max()) == 0.0 if cin_pad > cin else True
    err = float((got[:, :, :cin] - want).abs().max()) / amax
    assert err < 2.0 ** -21, err


def test_cta_pair_mode_matches_single_cta_kernel(tmp_path):
    """FLOWK_PAIR=1 (csrc/tc_gemm.cu: the dx-split 3x3 main loop on CTA pairs - cta_group::2, M = 256 MMAs, each CTA staging
    half of the weight rows, multicast commits) against fp64 and against the default single-CTA kernel, at the level-1 shape
    of cfg2 and with the one-MMA-per-column-shift variant.  The option is read once per process: subprocesses."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, torch, torch.nn.functional as F
sys.path.insert(0, %r)
import flowk
from flowk import tc
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, Cin, N, H, W = 64, 192, 96, 16, 16
x = torch.randn(B, Cin, H, W, device=dev)
w = torch.randn(N, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5
bias = torch.randn(N, device=dev)
rows = x.permute(0, 2, 3, 1).reshape(B * H * W, Cin).contiguous()
a_hi, a_lo = tc.split_rows_f16(rows)
w_hi, w_lo, sc = tc.conv_weight_operand_f16(w)
out = torch.full((B * H * W, N), float("nan"), device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
tc.conv_gemm(a_hi, a_lo, w_hi, w_lo, B, H, W, Cin, N, 9, tc.PRE_BIAS, tc.OUT_F32, bias=bias, out_f32=out, status=status,
             acc_scale=sc)
torch.cuda.synchronize()
assert int(status) == 0
ref = F.conv2d(x.double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 1).reshape(B * H * W, N)
err = float((out.double() - ref).abs().max() / ref.abs().max())
assert err < 2e-6, err
torch.save(out.cpu(), sys.argv[1])
print("ok", err)
''' % root
    outs = []
    for i, env in enumerate(({"FLOWK_PAIR": "0"}, {"FLOWK_PAIR": "1"}, {"FLOWK_PAIR": "1", "FLOWK_PAIR_NMMA": "3"})):
        path = os.path.join(tmp_path, "out%d.pt" % i)
        r = subprocess.run([sys.executable, "-c", code, path], env=dict(os.environ, **env), capture_output=True, text=True,
                           timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(torch.load(path))
    # same products in the same k order, accumulated by the same tensor cores
    assert float((outs[0] - outs[1]).abs().max()) <= 1e-5 * float(outs[0].abs().max())
    assert float((outs[0] - outs[2]).abs().max()) <= 1e-5 * float(outs[0].abs().max())


def test_pack_weight_f16_degenerate_inputs(tc):
    """All-zero weights keep scale 1 (no division by a zero maximum); a NaN / inf entry does not poison the scale of the
    finite ones (it saturates or stays NaN in its own slot only)."""
    dev = torch.device("cuda:0")
    w = torch.zeros(16, 8, 3, 3, device=dev)
    hi, lo, sc = tc.conv_weight_operand_f16(w)
    assert sc == 1.0 and float(hi.float().abs().max()) == 0.0 and float(lo.float().abs().max()) == 0.0
    w = torch.randn(16, 8, 1, 1, device=dev) * 0.1
    w[3, 2, 0, 0] = float("inf")
    hi, lo, sc = tc.conv_weight_operand_f16(w)
    assert math.isfinite(sc) and sc > 0
    got = (hi.double() + lo.double()).reshape(16, 1, 64)[:, 0, :8] * sc
    mask = torch.ones(16, 8, dtype=torch.bool, device=dev)
    mask[3] = False                                   # (the row with the inf: its own maximum is excluded from the scale)
    ref = w[:, :, 0, 0].double()
    assert float((got[mask] - ref[mask]).abs().max()) <= 2.0 ** -20 * float(ref[mask].abs().max())
