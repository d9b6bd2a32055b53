#!/usr/bin/env python
"""Device timing of the skeleton rasteriser + mask-loss path on one B200.

    python tools/bench_skeleton.py [--batch 256] [--size 256] [--steps 50] [--warmup 5]

One step = fused rasterise + max + weighted/clipped mask loss forward, then backward to the 2-D
keypoints (the SurS1 configuration: geodesic weight map + clip, model.py:185-188).  Prints one JSON
line with samples/s and the HBM roofline of the two kernels.  Algorithmic bytes per pixel:
forward  = write recon (4) + write winner byte (1) + read gt (4) + read weight (4) = 13
backward = read recon (4) + read winner byte (1) + read gt (4) + read weight (4)  = 13
(the keypoints, per-CTA partial sums and the [B,K,2] gradient are < 0.1 % and excluded).
Also times the un-maxed `draw_lines` [B,L,S,S] forward (write-bound: 4*L bytes per pixel).
"""
import argparse
import importlib
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--heat-batch", type=int, default=32, help="batch of the un-maxed draw_lines timing")
    ap.add_argument("--raw-only", action="store_true", help="only the bare C-ABI timing (profiling runs)")
    args = ap.parse_args()
    import torch
    pkg = importlib.import_module("x-as-supervision_b200")
    pkg.load_native()
    sk, synth, cabi = pkg.skeleton, pkg.synth, pkg._cabi
    dev = torch.device("cuda:0")
    B, S, K = args.batch, args.size, 18
    parent, child = sk.cal_links(synth.H36M_PARENTS, synth.LINE_SELECT)
    pose = synth.skeleton_pose2d(B, K, seed=60).to(dev)
    nb = min(B, 32)
    gt = synth.silhouette_mask(synth.skeleton_pose2d(nb, K, seed=61), S).repeat((B + nb - 1) // nb, 1, 1, 1)[:B].contiguous().to(dev)
    wmap = synth.geodesic_weight(gt.cpu()[:nb], seed=62).repeat((B + nb - 1) // nb, 1, 1, 1)[:B].contiguous().to(dev)
    # > 126 MB of L2 is touched per step at the default size (recon+gt+weight = 201 MB), so no explicit flush
    ev = {"fwd": [], "bwd": []}
    F = sk.SkeletonMaskLoss
    orig_f, orig_b = F.forward, F.backward
    rec = {"on": False}

    def wrap(kind, fn):
        def inner(*a, **k):
            if not rec["on"]:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            ev[kind].append((e0, e1))
            return out
        return staticmethod(inner)
    F.forward, F.backward = wrap("fwd", orig_f), wrap("bwd", orig_b)

    def step():
        kp = pose.clone().requires_grad_(True)
        recon, loss = sk.skeleton_mask_loss(kp, gt, wmap, S, parent, child, synth.BODY_WIDTH, use_clip=True)
        loss.backward()
        return loss

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    rec["on"] = True
    n0 = cabi.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step()
    t1.record()
    torch.cuda.synchronize()
    launches = cabi.launch_count() - n0
    rec["on"] = False
    ms = t0.elapsed_time(t1) / args.steps
    k_f = statistics.median(a.elapsed_time(b) for a, b in ev["fwd"])
    k_b = statistics.median(a.elapsed_time(b) for a, b in ev["bwd"])
    px = B * S * S
    peak = 6542.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    # ---- the bare C-ABI calls on preallocated buffers, back to back (what a CUDA-graph replay of the step costs)
    kp = pose.contiguous()
    skel = cabi.Skel(B, K, S, len(parent), K * 2, 2, synth.BODY_WIDTH)
    for i, (a, b_) in enumerate(zip(parent, child)):
        skel.parent[i], skel.child[i] = a, b_
    cfg = cabi.MaskLoss(B * S * S, cabi.MASK_WEIGHTED, 1)
    recon = torch.empty(B, 1, S, S, device=dev)
    lidx = torch.empty(B, S, S, dtype=torch.uint8, device=dev)
    sums = torch.zeros(4, device=dev)
    ws = torch.empty(cabi.lib.xsup_skel_ws_floats(skel), device=dev)
    g_loss = torch.ones(1, device=dev)
    g_kp = torch.empty(B, K, 2, device=dev)
    st = cabi.stream_ptr(dev)

    def raw_fwd():
        cabi.check(cabi.lib.xsup_skeleton_mask_fwd(kp.data_ptr(), skel, recon.data_ptr(), lidx.data_ptr(), gt.data_ptr(), wmap.data_ptr(),
                                                   cfg, sums.data_ptr(), ws.data_ptr(), st), "fwd")

    def raw_bwd():
        cabi.check(cabi.lib.xsup_skeleton_mask_bwd(kp.data_ptr(), skel, recon.data_ptr(), lidx.data_ptr(), None, gt.data_ptr(), wmap.data_ptr(),
                                                   cfg, sums.data_ptr(), g_loss.data_ptr(), g_kp.data_ptr(), ws.data_ptr(), st), "bwd")

    def raw_time(fn, n=args.steps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0.record()
        for _ in range(n):
            fn()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / n
    r_f, r_b = raw_time(raw_fwd), raw_time(raw_bwd)
    if args.raw_only:
        print(json.dumps({"raw_fwd_ms": r_f, "raw_bwd_ms": r_b}))
        return
    # un-maxed heat-maps (API parity with util.draw_lines)
    hb = min(args.heat_batch, B)
    hp = pose[:hb].contiguous()
    for _ in range(3):
        sk.draw_lines(hp, S, parent, child, synth.BODY_WIDTH)
    torch.cuda.synchronize()
    t0.record()
    for _ in range(10):
        sk.draw_lines(hp, S, parent, child, synth.BODY_WIDTH)
    t1.record()
    torch.cuda.synchronize()
    ms_heat = t0.elapsed_time(t1) / 10
    line = {"metric": "skeleton rasterise+max+mask-loss fwd+bwd samples/sec", "value": round(B / (ms * 1e-3), 1), "unit": "samples/s",
            "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "dtype": "f32",
            "config": {"workload": "draw_lines+max+compute_mask_reconstruction_loss(weight, use_clip), batch %d, %dx%d, 25 lines" % (B, S, S),
                       "l2": "recon+gt+weight = %.0f MB per step" % (px * 12 / 1e6)},
            "roofline": {"bound": "hbm", "kernel": "skeleton_mask_fwd_kernel + mask_loss_finalize_kernel (13 B/pixel), bare C-ABI calls back to back",
                         "achieved": round(13 * px / (r_f * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(13 * px / (r_f * 1e-3) / 1e9 / peak, 4), "ms_per_launch": round(r_f, 4),
                         "bwd_kernel": {"kernel": "skeleton_mask_bwd_kernel + skeleton_scatter_kernel (13 B/pixel)",
                                        "achieved": round(13 * px / (r_b * 1e-3) / 1e9, 1), "frac": round(13 * px / (r_b * 1e-3) / 1e9 / peak, 4),
                                        "ms_per_launch": round(r_b, 4)},
                         "autograd_function_body_ms": {"fwd": round(k_f, 4), "bwd": round(k_b, 4),
                                                       "note": "event pairs around the Python autograd.Function bodies (allocations + launches): host-launch-bound at this size"}},
            "draw_lines_fwd": {"batch": hb, "ms": round(ms_heat, 4), "achieved_gbs": round(hb * len(parent) * S * S * 4 / (ms_heat * 1e-3) / 1e9, 1)},
            "gpu_launches": int(launches)}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
