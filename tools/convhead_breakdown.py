#!/usr/bin/env python
"""Where the fused conv-head training step spends its time (B=256, C=256, K=17, 64x64): each piece timed alone."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("x-as-supervision_b200")
ops = pkg.load_native()
cabi = pkg._cabi
dev = torch.device("cuda:0")
B, K, D, C, NH, NS = 256, 17, 64, 256, 3, 15
KD, HW = K * D, D * D
g_ = torch.Generator(device=dev).manual_seed(3)
x = torch.randn(B, C, D, D, device=dev, generator=g_)
xcl = x.to(dtype=torch.bfloat16, memory_format=torch.channels_last)
wb = (torch.randn(KD, C, device=dev, generator=g_) / 16).bfloat16()
bias = torch.randn(KD, device=dev, generator=g_)
shape = cabi.make_shape(B, K, D, D, D, NH, NS, torch.bfloat16, cabi.HEAD_MULTI)
kps = torch.empty(B, NH, K, 3, device=dev); dmap = torch.empty(K, D, device=dev); idx = torch.empty(B, K, NH, dtype=torch.int64, device=dev)
stats = torch.empty(cabi.lib.xsup_stats_floats(shape), device=dev)
coef = torch.empty(cabi.lib.xsup_coef_floats(shape), device=dev)
gk = torch.randn(B, NH, K, 3, device=dev, generator=g_)
g = torch.empty(B, KD, HW, dtype=torch.bfloat16, device=dev)
gb = torch.empty(B, 4, KD, device=dev)
st = cabi.stream_ptr(dev)
x_flat = xcl.permute(0, 2, 3, 1).reshape(B, HW, C)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0.record()
    for _ in range(n):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return round(t0.elapsed_time(t1) / n, 4)


def fwd():
    cabi.check(cabi.lib.xsup_conv_head_fwd(xcl.data_ptr(), wb.data_ptr(), bias.data_ptr(), kps.data_ptr(), dmap.data_ptr(), idx.data_ptr(),
                                           stats.data_ptr(), None, shape, C, st), "fwd")


def coef_():
    cabi.check(cabi.lib.xsup_integral_coef(stats.data_ptr(), gk.data_ptr(), coef.data_ptr(), shape, st), "coef")


def bwd_g():
    cabi.check(cabi.lib.xsup_conv_head_bwd_g(xcl.data_ptr(), wb.data_ptr(), bias.data_ptr(), coef.data_ptr(), g.data_ptr(), gb.data_ptr(),
                                             shape, C, st), "bwd_g")


rowcoef = torch.empty(cabi.lib.xsup_conv_bwd_ws_floats(shape), device=dev)
dx = torch.empty(B, HW, C, dtype=torch.bfloat16, device=dev)
dw = torch.empty(KD, C, device=dev)
db = torch.empty(KD, device=dev)


def bwd(dx_, dw_, db_):
    cabi.check(cabi.lib.xsup_conv_head_bwd(xcl.data_ptr(), wb.data_ptr(), bias.data_ptr(), coef.data_ptr(), rowcoef.data_ptr(),
                                           dx_.data_ptr() if dx_ is not None else None, 0, dw_.data_ptr() if dw_ is not None else None,
                                           db_.data_ptr() if db_ is not None else None, shape, C, st), "bwd")


out = {"fwd_ms": timeit(fwd), "coef_ms": timeit(coef_), "bwd_g_ms": timeit(bwd_g)}
out["bwd_tc_dw_ms"] = timeit(lambda: bwd(None, dw, db))      # rowcoef + memsets + weight-stationary launch
out["bwd_tc_dx_ms"] = timeit(lambda: bwd(dx, None, None))    # rowcoef + activation-stationary launch
out["bwd_tc_both_ms"] = timeit(lambda: bwd(dx, dw, db))
xg = xcl.detach().clone().requires_grad_(True)
wg = wb.float().view(KD, C, 1, 1).requires_grad_(True)
bg = bias.clone().requires_grad_(True)


def train_step():
    xg.grad = wg.grad = bg.grad = None
    k, _, _ = ops.conv_integral_head_train(xg, wg, bg, K, NH, NS)
    k.backward(gk)


out["train_step_autograd_ms"] = timeit(train_step)
out["dx_bmm_bf16_ms"] = timeit(lambda: torch.matmul(g.transpose(1, 2), wb))
out["dx_bmm_f32out_ms"] = timeit(lambda: torch.bmm(g.transpose(1, 2), wb.unsqueeze(0).expand(B, KD, C), out_dtype=torch.float32))
out["dw_bmm_f32_partials_ms"] = timeit(lambda: torch.bmm(g, x_flat, out_dtype=torch.float32))
part = torch.bmm(g, x_flat, out_dtype=torch.float32)
out["dw_sum_ms"] = timeit(lambda: part.sum(0))
out["dbias_sum_ms"] = timeit(lambda: gb.sum(dim=(0, 1)))
out["pack_ms"] = timeit(lambda: cabi.check(cabi.lib.xsup_pack_nhwc_bf16(x.data_ptr(), xcl.data_ptr(), B, C, HW, st), "pack"))
flops = 2.0 * KD * C * B * HW
out["tflops"] = {k: round(flops / (out[k] * 1e-3) / 1e12, 1) for k in ("fwd_ms", "bwd_g_ms", "dx_bmm_bf16_ms", "dw_bmm_f32_partials_ms")}
out["tflops"].update({k: round(2 * flops / (out[k] * 1e-3) / 1e12, 1) for k in ("bwd_tc_dw_ms", "bwd_tc_dx_ms")})   # two GEMMs per launch
print(json.dumps(out))
