#!/usr/bin/env python
"""Duration of the two geometry + loss launches (K2 forward: world lift, loss terms, batch sums, selection; K2 backward: loss
VJP + coefficient blocks) inside the fused step, against the per-GPU batch.  CUDA-event pairs around each launch (ops.set_event_sink);
the streaming kernels before them are long enough for the host to run ahead, so the pairs time the kernels, not the launch gap.
    python tools/k2_probe.py [--res 32] [--dtype f32] [--batches 64 256 1024 4096] [--hypos 3 16]"""
import argparse, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=32)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--batches", type=int, nargs="*", default=[64, 256, 1024, 4096])
ap.add_argument("--hypos", type=int, nargs="*", default=[3, 16])
ap.add_argument("--sym", action="store_true", help="with the symmetry terms (bone / midpoint)")
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
pkg = importlib.import_module("x-as-supervision_b200")
ops, synth = pkg.load_native(), pkg.synth
dev = torch.device("cuda:0")
K, R, NS = 17, a.res, 15
tdt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
for NH in a.hypos:
    for B in a.batches:
        g = torch.Generator(device=dev).manual_seed(1)
        logits = torch.empty(B, K * R, R, R, device=dev, dtype=tdt)
        step = max(1, (1 << 28) // (K * R ** 3))
        for i in range(0, B, step):
            logits[i:i + step] = torch.randn(min(step, B - i), K * R, R, R, device=dev, generator=g).to(tdt)
        logits.requires_grad_(True)
        target = synth.pseudo_joints(B, K, seed=14).to(dev)
        cams = {k: v.to(dev) for k, v in synth.cameras(B, seed=15).items()}
        kw = dict(w_mse=3.0)
        if a.sym:
            kw.update(w_bone=0.1, w_kp=0.1, w_kp2d=0.1)

        def one():
            lp, ls, *_ = ops.integral_reproj_min_loss(logits, target, cams, K, NH, NS, **kw)
            (lp + ls).backward()
            logits.grad = None
        for _ in range(3):
            one()
        torch.cuda.synchronize()
        ev = {}
        ops.set_event_sink(ev)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(a.iters):
            one()
        t1.record()
        torch.cuda.synchronize()
        ops.set_event_sink(None)
        ms = {k: sum(s.elapsed_time(e) for s, e in v) / len(v) for k, v in ev.items()}
        print("res %d %s NH=%d B=%d: step %.4f ms | %s" % (R, a.dtype, NH, B, t0.elapsed_time(t1) / a.iters,
              "  ".join("%s %.1f us" % (k.replace("xsup.", ""), 1e3 * v) for k, v in ms.items())), flush=True)
        del logits
