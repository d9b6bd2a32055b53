"""Guard-band tests (`-m gpu`): every output / scratch buffer of the ring-, TMEM- and TMA-based kernels is carved out of a larger
allocation with 4 KB of a known byte pattern on both sides, the kernel is called straight through the C ABI, and the bands
must come back untouched.  compute-sanitizer is closed on this GPU pool (profiles/r2_sanitizer.md); this is the stand-in for
its memcheck pass on the kernels that compute their own addresses (bulk copies, tensor maps, TMA stores / reductions), run
at shapes that leave ragged tails: units that do not fill a ring stage, K*D that is not a multiple of 128, C < 256."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
GUARD = 4096
PATTERN = 0xA5


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as ge
    ge.build()
    p = importlib.import_module("x-as-supervision_b200")
    p.load_native()
    return p


class Guarded:
    """A tensor of `numel` elements of `dtype` inside a byte buffer with GUARD pattern bytes before and after it."""

    def __init__(self, numel, dtype, dev, fill=None):
        self.nbytes = numel * torch.empty((), dtype=dtype).element_size()
        pad = (-self.nbytes) % 256                      # keep the rear band 256-byte aligned too
        self.raw = torch.full((GUARD + self.nbytes + pad + GUARD,), PATTERN, dtype=torch.uint8, device=dev)
        self.t = self.raw[GUARD:GUARD + self.nbytes].view(dtype)
        self.rear0 = GUARD + self.nbytes
        if fill is not None:
            self.t.copy_(fill.reshape(-1))

    def ptr(self):
        return self.t.data_ptr()

    def intact(self):
        front = bool((self.raw[:GUARD] == PATTERN).all())
        rear = bool((self.raw[self.rear0:] == PATTERN).all())
        return front and rear


def _assert_intact(bufs):
    torch.cuda.synchronize()
    broken = [name for name, b in bufs.items() if not b.intact()]
    assert not broken, "out-of-bounds write next to: %s" % broken


@pytest.mark.parametrize("B,K,R,H,NH,NS,dtype", [
    (3, 17, 64, 64, 3, 15, torch.float32),      # the headline unit shape
    (5, 5, 32, 32, 3, 5, torch.bfloat16),       # 64 KB units, two tasks per warp and stage
    (2, 3, 32, 48, 2, 5, torch.float32),        # H != W: slices that do not fill the last stage
    (1, 2, 128, 128, 4, 31, torch.float32),     # 8 MB units
    (4, 6, 16, 16, 14, 1, torch.float32),       # tiny units, many hypotheses
])
def test_streaming_head_keeps_inside_its_buffers(pkg, B, K, R, H, NH, NS, dtype):
    cabi = pkg._cabi
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(B * 100 + R)
    logits = torch.randn(B, K * R, H, R, device=dev, generator=g).to(dtype).contiguous()
    shape = cabi.make_shape(B, K, R, H, R, NH, NS, dtype, cabi.HEAD_MULTI)
    out = {
        "kps": Guarded(B * NH * K * 3, torch.float32, dev),
        "dmap": Guarded(K * R, torch.float32, dev),
        "idx": Guarded(B * K * NH, torch.int64, dev),
        "stats": Guarded(int(cabi.lib.xsup_stats_floats(shape)), torch.float32, dev),
        "coef": Guarded(int(cabi.lib.xsup_coef_floats(shape)), torch.float32, dev),
        "g_logits": Guarded(logits.numel(), dtype, dev),
    }
    st = cabi.stream_ptr(dev)
    cabi.check(cabi.lib.xsup_integral_fwd(logits.data_ptr(), out["kps"].ptr(), out["dmap"].ptr(), out["idx"].ptr(), out["stats"].ptr(),
                                          shape, st), "xsup_integral_fwd")
    g_kps = torch.randn(B, NH, K, 3, device=dev, generator=g)
    cabi.check(cabi.lib.xsup_integral_bwd(logits.data_ptr(), out["stats"].ptr(), g_kps.data_ptr(), out["g_logits"].ptr(), out["coef"].ptr(),
                                          shape, st), "xsup_integral_bwd")
    _assert_intact(out)
    assert torch.isfinite(out["kps"].t).all() and torch.isfinite(out["g_logits"].t.float()).all()


@pytest.mark.parametrize("B,K,D,H,C,dx_f32", [
    (2, 17, 64, 64, 256, False),                # K*D = 1088: the last 128-row tile is half padding
    (3, 5, 64, 32, 192, True),                  # three k-blocks, fp32 d x (eight 16 KB drain chunks per item)
    (1, 2, 128, 64, 128, False),
    (2, 9, 32, 96, 64, True),                   # K*D = 288, C = 64: the smallest tensor-map boxes
])
def test_conv_fused_kernels_keep_inside_their_buffers(pkg, B, K, D, H, C, dx_f32):
    cabi = pkg._cabi
    dev = torch.device("cuda:0")
    NH, NS = 3, 5
    g = torch.Generator(device=dev).manual_seed(B * 10 + K)
    x = torch.randn(B, H, D, C, device=dev, generator=g).to(torch.bfloat16)          # channels-last storage [B, H, W, C]
    w = (torch.randn(K * D, C, device=dev, generator=g) / C ** 0.5).to(torch.bfloat16)
    bias = torch.randn(K * D, device=dev, generator=g)
    shape = cabi.make_shape(B, K, D, H, D, NH, NS, torch.bfloat16, cabi.HEAD_MULTI)
    HW = H * D
    out = {
        "kps": Guarded(B * NH * K * 3, torch.float32, dev),
        "dmap": Guarded(K * D, torch.float32, dev),
        "idx": Guarded(B * K * NH, torch.int64, dev),
        "stats": Guarded(int(cabi.lib.xsup_stats_floats(shape)), torch.float32, dev),
        "coef": Guarded(int(cabi.lib.xsup_coef_floats(shape)), torch.float32, dev),
        "rowcoef": Guarded(int(cabi.lib.xsup_conv_bwd_ws_floats(shape)), torch.float32, dev),
        "dx": Guarded(B * HW * C, torch.float32 if dx_f32 else torch.bfloat16, dev),
        "dw": Guarded(K * D * C, torch.float32, dev),
        "dbias": Guarded(K * D, torch.float32, dev),
    }
    st = cabi.stream_ptr(dev)
    cabi.check(cabi.lib.xsup_conv_head_fwd(x.data_ptr(), w.data_ptr(), bias.data_ptr(), out["kps"].ptr(), out["dmap"].ptr(), out["idx"].ptr(),
                                           out["stats"].ptr(), None, shape, C, st), "xsup_conv_head_fwd")
    g_kps = torch.randn(B, NH, K, 3, device=dev, generator=g)
    cabi.check(cabi.lib.xsup_integral_coef(out["stats"].ptr(), g_kps.data_ptr(), out["coef"].ptr(), shape, st), "xsup_integral_coef")
    cabi.check(cabi.lib.xsup_conv_head_bwd(x.data_ptr(), w.data_ptr(), bias.data_ptr(), out["coef"].ptr(), out["rowcoef"].ptr(), out["dx"].ptr(),
                                           1 if dx_f32 else 0, out["dw"].ptr(), out["dbias"].ptr(), shape, C, st), "xsup_conv_head_bwd")
    _assert_intact(out)
    for name in ("kps", "dx", "dw", "dbias"):
        assert torch.isfinite(out[name].t.float()).all(), name
    # every element of the outputs was written (the buffers started as the guard pattern, which is not a plausible gradient)
    pat32 = torch.tensor([PATTERN] * 4, dtype=torch.uint8).view(torch.float32).item()
    assert not bool((out["dw"].t == pat32).any()) and not bool((out["dbias"].t == pat32).any())
    if dx_f32:
        assert not bool((out["dx"].t == pat32).any())


def test_conv_fused_reruns_are_bit_identical_where_the_order_is_fixed(pkg):
    """Stand-in for racecheck on the tensor-core kernels: a race in the ring / TMEM / staging protocols shows up as run-to-run
    differences.  kps (K7) and d x (K8: every element is written once, by the TMA, from a TMEM accumulator whose summation order
    is fixed) must be bit-identical over repeated launches; d W / d bias are accumulated across samples by reduce-add / atomics in
    scheduling order, so they may differ in the last bits - bounded here at 1e-6 of their maximum."""
    ops = pkg.ops
    dev = torch.device("cuda:0")
    B, K, D, C, NH, NS = 20, 17, 64, 256, 3, 15            # 180 + 640 items: several per CTA, so rings and barriers wrap many times
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn(B, C, D, D, device=dev, generator=g).to(dtype=torch.bfloat16, memory_format=torch.channels_last)
    w = torch.randn(K * D, C, device=dev, generator=g) / C ** 0.5
    bias = torch.randn(K * D, device=dev, generator=g)
    gk = torch.randn(B, NH, K, 3, device=dev, generator=g)
    ref = None
    for _ in range(6):
        xd = x.clone().requires_grad_(True)
        wd = w.clone().requires_grad_(True)
        bd = bias.clone().requires_grad_(True)
        kps, _, idx = ops.conv_integral_head_train(xd, wd, bd, K, NH, NS)
        kps.backward(gk)
        cur = (kps.detach().clone(), idx.clone(), xd.grad.clone(), wd.grad.clone(), bd.grad.clone())
        if ref is None:
            ref = cur
            continue
        assert torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1]), "forward differs between launches"
        assert torch.equal(cur[2], ref[2]), "d x differs between launches"
        for a, b in ((cur[3], ref[3]), (cur[4], ref[4])):
            assert float((a - b).abs().max()) <= 1e-6 * float(b.abs().max())
