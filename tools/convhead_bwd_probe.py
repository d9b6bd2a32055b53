#!/usr/bin/env python
"""Runs only the tensor-core backward of the conv-fused head (xsup_conv_head_bwd) a few times: the target of ncu captures.
    python tools/convhead_bwd_probe.py [--batch 64] [--iters 3] [--what both|dx|dw]
TRACE=1 additionally prints the clock64 timeline of CTA 0 (library built with XSUP_NVCC_EXTRA="-DXSUP_TRACE", see conv_head_bwd.cu)."""
import argparse, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--what", default="both")
a = ap.parse_args()
pkg = importlib.import_module("x-as-supervision_b200")
ops = pkg.load_native()
cabi = pkg._cabi
dev = torch.device("cuda:0")
B, K, D, C, NH, NS = a.batch, 17, 64, 256, 3, 15
KD, HW = K * D, D * D
g_ = torch.Generator(device=dev).manual_seed(3)
xcl = torch.randn(B, C, D, D, device=dev, generator=g_).to(dtype=torch.bfloat16, memory_format=torch.channels_last)
wb = (torch.randn(KD, C, device=dev, generator=g_) / 16).bfloat16()
bias = torch.randn(KD, device=dev, generator=g_)
shape = cabi.make_shape(B, K, D, D, D, NH, NS, torch.bfloat16, cabi.HEAD_MULTI)
kps = torch.empty(B, NH, K, 3, device=dev); dmap = torch.empty(K, D, device=dev); idx = torch.empty(B, K, NH, dtype=torch.int64, device=dev)
stats = torch.empty(cabi.lib.xsup_stats_floats(shape), device=dev)
coef = torch.empty(cabi.lib.xsup_coef_floats(shape), device=dev)
gk = torch.randn(B, NH, K, 3, device=dev, generator=g_)
rowcoef = torch.empty(cabi.lib.xsup_conv_bwd_ws_floats(shape), device=dev)
dx = torch.empty(B, HW, C, dtype=torch.bfloat16, device=dev)
dw = torch.empty(KD, C, device=dev); db = torch.empty(KD, device=dev)
st = cabi.stream_ptr(dev)
trace = torch.zeros(16 * 16, dtype=torch.int64, device=dev)
if os.environ.get("TRACE"):      # clock64 timeline of CTA 0, tiles 8..23 (see conv_head_bwd.cu TRACE)
    os.environ["XSUP_CONVBWD_TRACE"] = str(trace.data_ptr())
cabi.check(cabi.lib.xsup_conv_head_fwd(xcl.data_ptr(), wb.data_ptr(), bias.data_ptr(), kps.data_ptr(), dmap.data_ptr(), idx.data_ptr(),
                                       stats.data_ptr(), None, shape, C, st), "fwd")
cabi.check(cabi.lib.xsup_integral_coef(stats.data_ptr(), gk.data_ptr(), coef.data_ptr(), shape, st), "coef")
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
t0.record()
for _ in range(a.iters):
    cabi.check(cabi.lib.xsup_conv_head_bwd(xcl.data_ptr(), wb.data_ptr(), bias.data_ptr(), coef.data_ptr(), rowcoef.data_ptr(),
                                           dx.data_ptr() if a.what != "dw" else None, 0, dw.data_ptr() if a.what != "dx" else None,
                                           db.data_ptr() if a.what != "dx" else None, shape, C, st), "bwd")
t1.record()
torch.cuda.synchronize()
print("B=%d %s: %.4f ms per call" % (B, a.what, t0.elapsed_time(t1) / a.iters))
if os.environ.get("TRACE"):
    tr = trace.cpu().view(16, 16)
    base = int(tr[0, 0])
    names = ["mma:iter", "aempty", "xfull", "S-issued", "gfull", "MMA2-issued", "prod:xempty", "-", "epi:afull", "ld-done", "math-done", "gbuf-free",
             "g-published", "item:last tile done", "item:dfull", "item:drained"]
    for gi in range(16):
        print("tile %2d: " % (gi + 8) + "  ".join("%s=%d" % (names[j], int(tr[gi, j]) - base) for j in range(16) if names[j] != "-"))
