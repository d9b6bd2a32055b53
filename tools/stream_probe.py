#!/usr/bin/env python
"""Times the streaming forward (xsup_integral_fwd) and backward (xsup_integral_bwd) alone at one shape; with TRACE=1 and a
library built with XSUP_NVCC_EXTRA="-DXSUP_TRACE" prints the clock64 timeline of CTA 0 of the forward (integral_fwd.cu K1TRACE).
    python tools/stream_probe.py [--res 32] [--batch 4096] [--dtype bf16] [--iters 5]"""
import argparse, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=32)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
pkg = importlib.import_module("x-as-supervision_b200")
ops = pkg.load_native()
dev = torch.device("cuda:0")
K, R, NH, NS = 17, a.res, 3, 15
tdt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
g = torch.Generator(device=dev).manual_seed(1)
logits = torch.empty(a.batch, K * R, R, R, device=dev, dtype=tdt)
step = max(1, (1 << 28) // (K * R ** 3))
for i in range(0, a.batch, step):
    logits[i:i + step] = torch.randn(min(step, a.batch - i), K * R, R, R, device=dev, generator=g).to(tdt)
trace = torch.zeros(16 * 16, dtype=torch.int64, device=dev)
if os.environ.get("TRACE"):
    os.environ["XSUP_K1_TRACE"] = str(trace.data_ptr())
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    ops.integral_multi_head(logits, K, NH, NS)
torch.cuda.synchronize()
t0.record()
for _ in range(a.iters):
    ops.integral_multi_head(logits, K, NH, NS)
t1.record()
torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / a.iters
print("fwd res %d B=%d %s: %.4f ms per call, %.1f GB/s" % (R, a.batch, a.dtype, ms, logits.numel() * logits.element_size() / ms / 1e6))
_, shape, kps, _, _, stats = ops._head_forward(logits, K, NH, NS, ops.cabi.HEAD_MULTI)
g_kps = torch.randn(kps.shape, device=dev, generator=g)
g_logits = torch.empty_like(logits)
for _ in range(2):
    ops._head_backward(logits, stats, shape, g_kps, inplace=False)
torch.cuda.synchronize()
t0.record()
for _ in range(a.iters):
    ops._head_backward(logits, stats, shape, g_kps, inplace=False)
t1.record()
torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / a.iters
print("bwd res %d B=%d %s: %.4f ms per call (coef + stream), %.1f GB/s" % (R, a.batch, a.dtype, ms, 2 * logits.numel() * logits.element_size() / ms / 1e6))
if os.environ.get("TRACE"):
    tr = trace.cpu().view(16, 16)
    base = int(tr[0, 0])
    names = {0: "claim", 1: "stage0", 2: "stageN", 4: "c0:first", 5: "c0:flush", 8: "fin:start", 9: "fin:merged", 10: "fin:done"}
    for u in range(16):
        print("unit %2d: " % (u + 4) + "  ".join("%s=%d" % (n, int(tr[u, j]) - base) for j, n in names.items()))
