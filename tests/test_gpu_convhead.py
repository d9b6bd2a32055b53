"""GPU parity tests (`-m gpu`) of the conv-fused forward: the tcgen05 GEMM's logits against a plain
PyTorch fp32 reference of the same op on the same bf16-rounded operands, and the head outputs against the streaming
kernel run on those logits and against the fp64 oracle.  Tolerances: logits 1e-5 relative to their scale (exact bf16
products, fp32 accumulation in a different order), peak indices bit-exact, coordinates 1e-5."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import __graft_entry__ as ge
    ge.build()
    return importlib.import_module("x-as-supervision_b200").load_native()


def _case(B, K, D, C, seed, peaked=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, D, D, generator=g)
    w = torch.randn(K * D, C, generator=g) / C ** 0.5
    if peaked:                                   # give the logits structure: a few strong rows / pixels
        w[::7] *= 3.0
        x[:, :, D // 3, D // 2] += 2.0
    bias = torch.randn(K * D, generator=g)
    return x, w, bias


@pytest.mark.parametrize("B,K,D,C,NH,NS", [(2, 2, 64, 64, 3, 15), (2, 3, 64, 256, 3, 15), (3, 17, 64, 256, 3, 15), (2, 4, 32, 128, 2, 5),
                                           (1, 18, 64, 256, 3, 15), (2, 1, 128, 64, 3, 15)])
def test_conv_head_matches_reference(ops, oracle, B, K, D, C, NH, NS):
    dev = torch.device("cuda:0")
    x, w, bias = _case(B, K, D, C, seed=B * 100 + K)
    xb, wb = x.bfloat16().float(), w.bfloat16().float()
    kps, dmap, idx, logits = ops.conv_integral_head(x.to(dev), w.to(dev), bias.to(dev), K, NH, NS, return_logits=True)
    ref = torch.einsum("oc,bchw->bohw", wb.double(), xb.double()) + bias.double().view(1, -1, 1, 1)
    err = float((logits.cpu().double() - ref).abs().max()) / float(ref.abs().max())
    assert err < 1e-5, err
    # head outputs: the streaming kernel on the materialised logits, and the fp64 oracle on the exact logits
    k2, d2, i2 = ops.integral_multi_head(logits, K, NH, NS)
    assert torch.equal(idx, i2)
    assert float((kps - k2).abs().max()) < 1e-5 and float((dmap - d2).abs().max()) < 1e-5
    okps, odmap, oidx = oracle.integral_multi(ref, K, NH, NS)
    assert torch.equal(idx.cpu(), oidx)
    assert float((kps.cpu().double() - okps).abs().max()) < 1e-5
    assert float((dmap.cpu().double() - odmap).abs().max()) < 1e-5 * float(odmap.abs().max())
    # without the validation output, from a channels-last bf16 tensor used in place, and without a bias
    xcl = x.to(dev).to(dtype=torch.bfloat16, memory_format=torch.channels_last)
    k3, d3, i3 = ops.conv_integral_head(xcl, w.to(dev).view(K * D, C, 1, 1), bias.to(dev), K, NH, NS)
    assert torch.equal(k3, kps) and torch.equal(i3, idx) and torch.equal(d3, dmap)
    k4, _, _, l4 = ops.conv_integral_head(xcl, w.to(dev), None, K, NH, NS, return_logits=True)
    ref0 = ref - bias.double().view(1, -1, 1, 1)
    assert float((l4.cpu().double() - ref0).abs().max()) / float(ref0.abs().max()) < 1e-5


def test_pack_kernel_equals_torch_conversion(ops):
    """NCHW fp32 -> channels-last bf16 in one pass: bit-identical to torch's cast + permute."""
    dev = torch.device("cuda:0")
    cabi = importlib.import_module("x-as-supervision_b200._cabi")
    g = torch.Generator().manual_seed(1)
    for B, C, H, W in ((3, 256, 64, 64), (2, 64, 32, 32), (1, 128, 8, 8)):
        x = (torch.randn(B, C, H, W, generator=g) * 3).to(dev)
        out = torch.empty((B, C, H, W), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last)
        cabi.check(cabi.lib.xsup_pack_nhwc_bf16(x.data_ptr(), out.data_ptr(), B, C, H * W, cabi.stream_ptr(dev)), "pack")
        ref = x.to(dtype=torch.bfloat16, memory_format=torch.channels_last)
        assert out.is_contiguous(memory_format=torch.channels_last) and torch.equal(out, ref)
    # and conv_integral_head takes the fp32 NCHW tensor through it
    x, w, bias = _case(2, 3, 64, 256, seed=5)
    a = ops.conv_integral_head(x.to(dev), w.to(dev), bias.to(dev), 3, 3, 15)
    b = ops.conv_integral_head(x.to(dev).to(dtype=torch.bfloat16, memory_format=torch.channels_last), w.to(dev), bias.to(dev), 3, 3, 15)
    assert all(torch.equal(u, v) for u, v in zip(a, b))


def test_conv_head_shape_errors(ops):
    dev = torch.device("cuda:0")
    with pytest.raises(RuntimeError, match="divide 128"):
        ops.conv_integral_head(torch.zeros(1, 64, 48, 48, device=dev), torch.zeros(2 * 48, 64, device=dev), None, 2, 3, 15)
    with pytest.raises(RuntimeError, match="channels"):
        ops.conv_integral_head(torch.zeros(1, 96, 64, 64, device=dev), torch.zeros(2 * 64, 96, device=dev), None, 2, 3, 15)


def test_detector_forward_fused_matches_forward(ops):
    """KPDetector3DMulti.forward_fused on a ResPoseNet-shaped net: same outputs as the unfused forward up to the bf16
    rounding of the conv operands (compared against forward() on a copy of the net whose last conv sees the same
    rounded operands)."""
    from torch import nn
    det_mod = importlib.import_module("x-as-supervision_b200.detector")
    dev = torch.device("cuda:0")
    K, D, C = 17, 64, 256

    class Head(nn.Module):
        def __init__(self):
            super().__init__()
            self.features = nn.ModuleList([nn.Conv2d(8, C, 3, padding=1), nn.ReLU(), nn.Conv2d(C, K * D, 1, bias=True)])

        def forward(self, x):
            for layer in self.features:
                x = layer(x)
            return x

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.backbone = nn.Conv2d(3, 8, 3, padding=1)
            self.head = Head()

        def forward(self, x):
            return self.head(self.backbone(x))

    torch.manual_seed(0)
    net = Net().to(dev)
    det = det_mod.KPDetector3DMulti("resnet_multi", K, D, 3, 15, net=net).eval()
    x = torch.randn(2, 3, D, D, device=dev)
    kps_f, dmap_f = det.forward_fused(x)
    with torch.no_grad():
        y = net.backbone(x)
        for layer in list(net.head.features)[:-1]:
            y = layer(y)
        last = net.head.features[-1]
        logits = torch.nn.functional.conv2d(y.bfloat16().double(), last.weight.bfloat16().double(), last.bias.double())
        kps_r, dmap_r, _ = ops.integral_multi_head(logits.float(), K, 3, 15)
    assert float((kps_f - kps_r).abs().max()) < 1e-5 and float((dmap_f - dmap_r).abs().max()) < 1e-5
    with pytest.raises(RuntimeError, match="1x1"):
        det_mod.KPDetector3DMulti("x", K, D, 3, 15, net=nn.Identity()).forward_fused(x)


def test_conv_head_other_shapes(ops, oracle):
    """C = 192 (three k-blocks), non-square maps (H != W; D == W as the reference requires), one joint per CTA (D = 128)."""
    dev = torch.device("cuda:0")
    for B, K, D, H, C in ((2, 5, 64, 32, 192), (1, 2, 128, 64, 128), (2, 9, 32, 96, 64)):
        g = torch.Generator().manual_seed(B + K + D)
        x = torch.randn(B, C, H, D, generator=g)
        w = torch.randn(K * D, C, generator=g) / C ** 0.5
        bias = torch.randn(K * D, generator=g)
        kps, dmap, idx, logits = ops.conv_integral_head(x.to(dev), w.to(dev), bias.to(dev), K, 3, 15, return_logits=True)
        ref = torch.einsum("oc,bchw->bohw", w.bfloat16().double(), x.bfloat16().double()) + bias.double().view(1, -1, 1, 1)
        assert float((logits.cpu().double() - ref).abs().max()) / float(ref.abs().max()) < 1e-5
        okps, odmap, oidx = oracle.integral_multi(ref, K, 3, 15)
        assert torch.equal(idx.cpu(), oidx)
        assert float((kps.cpu().double() - okps).abs().max()) < 1e-5
        assert float((dmap.cpu().double() - odmap).abs().max()) < 1e-5 * float(odmap.abs().max())


@pytest.mark.parametrize("B,K,D,C", [(2, 3, 64, 128), (3, 17, 64, 256), (2, 4, 32, 64)])
def test_conv_head_backward_matches_autograd(ops, oracle, B, K, D, C):
    """d x, d W, d bias of the differentiable conv-fused head against fp64 autograd of the reference op sequence
    (1x1 conv -> integral head) on the same bf16-rounded operands.  d loss / d logits is written once in bf16, so the
    tolerance is the bf16 one of the north star: 2^-8 relative to the largest gradient (measured ~1e-3); the
    recomputed d loss / d logits itself is checked to that bound element-wise."""
    dev = torch.device("cuda:0")
    NH, NS = 3, 15 if D >= 64 else 5
    x, w, bias = _case(B, K, D, C, seed=B * 31 + K)
    gen = torch.Generator().manual_seed(77)
    gk = torch.randn(B, NH, K, 3, generator=gen)
    xd = x.to(dev).requires_grad_(True)
    wd = w.to(dev).view(K * D, C, 1, 1).requires_grad_(True)
    bd = bias.to(dev).requires_grad_(True)
    kps, dmap, idx = ops.conv_integral_head_train(xd, wd, bd, K, NH, NS)
    kps.backward(gk.to(dev))
    # fp64 reference on the bf16-rounded operands
    x64 = x.bfloat16().double().requires_grad_(True)
    w64 = w.bfloat16().double().requires_grad_(True)
    b64 = bias.double().requires_grad_(True)
    logits = torch.einsum("oc,bchw->bohw", w64, x64) + b64.view(1, -1, 1, 1)
    logits.retain_grad()
    okps, _, oidx = oracle.integral_multi(logits, K, NH, NS)
    assert torch.equal(idx.cpu(), oidx) and float((kps.detach().cpu().double() - okps.detach()).abs().max()) < 1e-5
    okps.backward(gk.double())
    tol = 2.0 ** -8
    for name, ours, ref in (("dx", xd.grad, x64.grad), ("dW", wd.grad.view(K * D, C), w64.grad), ("dbias", bd.grad, b64.grad)):
        err = float((ours.cpu().double() - ref).abs().max()) / float(ref.abs().max())
        assert err < tol, (name, err)
    assert xd.grad.shape == xd.shape and xd.grad.dtype == xd.dtype and wd.grad.shape == wd.shape


@pytest.mark.parametrize("B,K,D,H,C,use_bias", [(2, 5, 64, 32, 192, True), (1, 2, 128, 64, 128, False), (5, 9, 32, 96, 64, True)])
def test_conv_head_backward_bf16_channels_last(ops, oracle, B, K, D, H, C, use_bias):
    """The tensor-core backward (xsup_conv_head_bwd) from bf16 channels-last activations - d x comes back bf16, channels-last,
    written by the TMA from the staged TMEM accumulator - on non-square maps, three k-blocks, one joint per 128 rows and
    without a bias; weight rows that do not fill the last 128-row tile (K*D = 320, 288).  Reference: fp64 autograd of
    conv -> integral head on the same bf16-rounded operands; tolerance 2^-8 of the largest gradient (one bf16 rounding of
    d loss / d logits, one of d x)."""
    dev = torch.device("cuda:0")
    NH, NS = 3, 5
    g = torch.Generator().manual_seed(B * 7 + K)
    x = torch.randn(B, C, H, D, generator=g)
    w = torch.randn(K * D, C, generator=g) / C ** 0.5
    bias = torch.randn(K * D, generator=g) if use_bias else None
    gk = torch.randn(B, NH, K, 3, generator=g)
    xd = x.to(dev).to(dtype=torch.bfloat16, memory_format=torch.channels_last).requires_grad_(True)
    wd = w.to(dev).requires_grad_(True)
    bd = bias.to(dev).requires_grad_(True) if use_bias else None
    kps, dmap, idx = ops.conv_integral_head_train(xd, wd, bd, K, NH, NS)
    kps.backward(gk.to(dev))
    x64 = x.bfloat16().double().requires_grad_(True)
    w64 = w.bfloat16().double().requires_grad_(True)
    b64 = bias.double().requires_grad_(True) if use_bias else None
    logits = torch.einsum("oc,bchw->bohw", w64, x64)
    if use_bias:
        logits = logits + b64.view(1, -1, 1, 1)
    okps, _, oidx = oracle.integral_multi(logits, K, NH, NS)
    assert torch.equal(idx.cpu(), oidx)
    okps.backward(gk.double())
    assert xd.grad.dtype == torch.bfloat16 and xd.grad.is_contiguous(memory_format=torch.channels_last)
    checks = [("dx", xd.grad.float(), x64.grad), ("dW", wd.grad, w64.grad)] + ([("dbias", bd.grad, b64.grad)] if use_bias else [])
    for name, ours, ref in checks:
        err = float((ours.cpu().double() - ref).abs().max()) / float(ref.abs().max())
        assert err < 2.0 ** -8, (name, err)


@pytest.mark.parametrize("reduction,sym", [("batch", True), ("sample", False), ("joint", False)])
def test_conv_fused_reproj_min_loss(ops, oracle, synth, reduction, sym):
    """The whole per-camera op with the head's final conv pulled in (ConvIntegralReprojMinLoss): losses, selected slots,
    coordinates and d x / d W / d bias against the fp64 oracle run on logits formed from the same bf16-rounded operands."""
    dev = torch.device("cuda:0")
    B, K, D, C, NH, NS = 4, 17, 64, 128, 3, 15
    x, w, bias = _case(B, K, D, C, seed=91)
    target = synth.pseudo_joints(B, K, seed=92)
    cams = synth.cameras(B, seed=93)
    weights = dict(w_mse=1.0, w_bone=0.1, w_kp=0.1, w_kp2d=0.0) if sym else dict(w_mse=3.0)
    xd = x.to(dev).requires_grad_(True)
    wd = w.to(dev).requires_grad_(True)
    bd = bias.to(dev).requires_grad_(True)
    dcams = {k: v.to(dev) for k, v in cams.items()}
    n0 = ops.launch_count()
    lp, ls, sel, kps, world, dmap, idx = ops.conv_integral_reproj_min_loss(xd, wd, bd, target.to(dev), dcams, K, NH, NS, reduction=reduction, **weights)
    (lp + ls).backward()
    torch.cuda.synchronize()
    assert ops.launch_count() - n0 <= 7          # pack, conv fwd, loss fwd, loss bwd + coef, row coefficients, d W, d x
    x64 = x.bfloat16().double().requires_grad_(True)
    w64 = w.bfloat16().double().requires_grad_(True)
    b64 = bias.double().requires_grad_(True)
    logits = torch.einsum("oc,bchw->bohw", w64, x64) + b64.view(1, -1, 1, 1)
    c64 = {k: v.double() for k, v in cams.items()}
    olp, ols, osel, okps, oworld, odmap, oidx = oracle.fused_forward(logits, K, NH, NS, target.double(), c64, reduction=reduction, **weights)
    (olp + ols).backward()
    assert torch.equal(idx.cpu(), oidx) and torch.equal(sel.cpu(), osel)
    assert float((kps.detach().cpu().double() - okps.detach()).abs().max()) < 1e-5
    assert abs(float(lp + ls) - float(olp + ols)) <= 1e-5 * abs(float(olp + ols))
    for name, ours, ref in (("dx", xd.grad, x64.grad), ("dW", wd.grad, w64.grad), ("dbias", bd.grad, b64.grad)):
        err = float((ours.cpu().double() - ref).abs().max()) / float(ref.abs().max())
        assert err < 2.0 ** -8, (name, err)


@pytest.mark.parametrize("B,K,D,H,C", [(2, 17, 64, 64, 256), (3, 3, 64, 32, 128), (2, 4, 32, 32, 64), (1, 2, 128, 64, 192)])
def test_conv_head_tf32_variant(ops, oracle, B, K, D, H, C):
    """precision='tf32' (fp32 operands, kind::tf32 MMAs): the logits against the exact fp64 conv of the UNROUNDED operands at the tf32
    bound (inputs keep 10 mantissa bits: relative-to-max error of a C-term dot product well under 2e-3), at least 4x closer than
    the bf16 path on the same inputs, and the head outputs against the streaming kernel run on those very logits."""
    dev = torch.device("cuda:0")
    NH, NS = 3, 5
    g = torch.Generator().manual_seed(B * 3 + K)
    x = torch.randn(B, C, H, D, generator=g)
    w = torch.randn(K * D, C, generator=g) / C ** 0.5
    w[::5] *= 3.0
    bias = torch.randn(K * D, generator=g)
    ref = torch.einsum("oc,bchw->bohw", w.double(), x.double()) + bias.double().view(1, -1, 1, 1)
    kps, dmap, idx, logits = ops.conv_integral_head(x.to(dev), w.to(dev), bias.to(dev), K, NH, NS, return_logits=True, precision="tf32")
    _, _, _, logits16 = ops.conv_integral_head(x.to(dev), w.to(dev), bias.to(dev), K, NH, NS, return_logits=True, precision="bf16")
    scale = float(ref.abs().max())
    e32 = float((logits.cpu().double() - ref).abs().max()) / scale
    e16 = float((logits16.cpu().double() - ref).abs().max()) / scale
    assert e32 < 2e-3 and e32 * 4 < e16, (e32, e16)
    skps, sdmap, sidx = ops.integral_multi_head(logits, K, NH, NS)
    assert torch.equal(idx, sidx)
    assert float((kps - skps).abs().max()) < 1e-5 and float((dmap - sdmap).abs().max()) < 1e-5 * float(sdmap.abs().max())
