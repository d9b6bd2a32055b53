#!/usr/bin/env python
"""Host cost of one eager fused step (autograd.Function forward + backward: Python, ctypes, torch allocations, launches): wall time
per step with the GPU kept far from the bottleneck (tiny volumes), and a cProfile of the same loop.
    python tools/host_profile.py [--batch 32] [--res 16] [--steps 2000]"""
import argparse, cProfile, importlib, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--res", type=int, default=16)
ap.add_argument("--steps", type=int, default=2000)
ap.add_argument("--top", type=int, default=28)
a = ap.parse_args()
pkg = importlib.import_module("x-as-supervision_b200")
ops, synth = pkg.load_native(), pkg.synth
dev = torch.device("cuda:0")
K, R, NH, NS, B = 17, a.res, 3, 15, a.batch
logits = torch.randn(B, K * R, R, R, device=dev, requires_grad=True)
target = synth.pseudo_joints(B, K, seed=14).to(dev)
cams = {k: v.to(dev) for k, v in synth.cameras(B, seed=15).items()}


def step():
    logits.grad = None
    lp, ls, *_ = ops.integral_reproj_min_loss(logits, target, cams, K, NH, NS, w_mse=3.0)
    (lp + ls).backward()


for _ in range(50):
    step()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(a.steps):
    step()
host = time.perf_counter() - t
torch.cuda.synchronize()
total = time.perf_counter() - t
print("B=%d res=%d: host %.1f us per step to enqueue, %.1f us per step until the GPU is done" % (B, R, 1e6 * host / a.steps, 1e6 * total / a.steps))
pr = cProfile.Profile()
pr.enable()
for _ in range(a.steps):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(a.top)
print("\n".join(l for l in s.getvalue().splitlines() if l.strip())[:6000])
