#!/usr/bin/env python
"""Device timing of the conv-fused forward against the unfused path on one B200.

    python tools/bench_convhead.py [--batch 256] [--steps 20]

fused     : xsup_conv_head_fwd on channels-last bf16 activations resident in HBM (logits never materialised)
fused+cvt : the same, from the NCHW fp32 tensor a plain backbone produces (one conversion pass added)
unfused   : torch 1x1 conv (cuDNN/cuBLAS; fp32 with TF32 allowed as PyTorch's default, and bf16) writing the
            [B, K*D, H, W] logits, then the streaming kernel xsup_integral_fwd reading them
Roofline of the fused kernel: tensor - algorithmic flops = 2 * K*D * C per pixel over the launch time, against
the measured dense bf16 peak of MEASURED_PEAKS.json (burst figure: the kernel is timed alone).
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--fused-only", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.nn.functional as F
    pkg = importlib.import_module("x-as-supervision_b200")
    ops = pkg.load_native()
    dev = torch.device("cuda:0")
    B, K, D, C, NH, NS = args.batch, 17, 64, 256, 3, 15
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn(B, C, D, D, device=dev, generator=g)
    w = torch.randn(K * D, C, device=dev, generator=g) / 16
    bias = torch.randn(K * D, device=dev, generator=g)
    xcl = x.to(dtype=torch.bfloat16, memory_format=torch.channels_last)
    wb = w.bfloat16()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timeit(fn, n=args.steps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0.record()
        for _ in range(n):
            fn()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / n

    n0 = ops.launch_count()
    ms_fused = timeit(lambda: ops.conv_integral_head(xcl, wb, bias, K, NH, NS))
    launches = ops.launch_count() - n0
    flops = 2.0 * K * D * C * B * D * D
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        peak = 1590.0
    out = {"metric": "conv1x1 + integral head forward samples/sec", "config": {"workload": "B=%d, C=%d -> K*D=%d, %dx%d, NH=%d" % (B, C, K * D, D, D, NH)},
           "fused": {"ms": round(ms_fused, 4), "samples_per_s": round(B / ms_fused * 1e3, 1)},
           "roofline": {"bound": "tensor", "kernel": "conv_head_fwd_kernel", "achieved": round(flops / (ms_fused * 1e-3) / 1e12, 1), "peak": peak,
                        "unit": "TFLOP/s", "frac": round(flops / (ms_fused * 1e-3) / 1e12 / peak, 4),
                        "hbm_traffic_algorithmic_gb": round(xcl.numel() * 2 / 1e9, 3)},
           "gpu_launches": int(launches)}
    if not args.fused_only:
        ms_cvt = timeit(lambda: ops.conv_integral_head(x, wb, bias, K, NH, NS))
        w4, wb4 = w.view(K * D, C, 1, 1), wb.view(K * D, C, 1, 1)

        def unfused32():
            logits = F.conv2d(x, w4, bias)
            return ops.integral_multi_head(logits, K, NH, NS)

        def unfused16():
            logits = F.conv2d(xcl, wb4, bias.bfloat16())
            return ops.integral_multi_head(logits.contiguous(), K, NH, NS)
        ms_u32 = timeit(unfused32)
        ms_conv32 = timeit(lambda: F.conv2d(x, w4, bias))
        ms_u16 = timeit(unfused16)
        out["fused_from_nchw_fp32"] = {"ms": round(ms_cvt, 4), "samples_per_s": round(B / ms_cvt * 1e3, 1)}
        out["unfused_fp32_tf32conv"] = {"ms": round(ms_u32, 4), "conv_ms": round(ms_conv32, 4), "samples_per_s": round(B / ms_u32 * 1e3, 1)}
        out["unfused_bf16_conv_bf16_logits"] = {"ms": round(ms_u16, 4), "samples_per_s": round(B / ms_u16 * 1e3, 1)}
        out["speedup_vs_unfused_fp32"] = round(ms_u32 / ms_fused, 2)
        # the strongest unfused pipeline we know of: a bf16 batched GEMM straight into NCHW bf16 logits (cuBLAS), then the
        # streaming kernel on those bf16 logits
        x_flat_t = xcl.permute(0, 2, 3, 1).reshape(B, D * D, C).transpose(1, 2)           # [B, C, HW] view of the channels-last storage
        wexp = wb.unsqueeze(0).expand(B, K * D, C)

        def unfused_best():
            logits = torch.baddbmm(bias.bfloat16().view(1, -1, 1), wexp, x_flat_t)           # [B, K*D, HW] bf16
            return ops.integral_multi_head(logits.view(B, K * D, D, D), K, NH, NS)
        ms_ub = timeit(unfused_best)
        out["unfused_bf16_bmm_bf16_logits"] = {"ms": round(ms_ub, 4), "samples_per_s": round(B / ms_ub * 1e3, 1),
                                                "note": "cuBLAS batched GEMM writing 2.28 GB of bf16 logits + integral_fwd_kernel<bf16> reading them"}
        out["speedup_vs_best_unfused"] = round(ms_ub / ms_fused, 2)
        # ---- training step: forward + backward to d x, d W, d bias (gradient of a random cotangent on kps)
        gk = torch.randn(B, NH, K, 3, device=dev, generator=g)
        xg = x.clone().requires_grad_(True)
        xclg = xcl.clone().requires_grad_(True)
        wg = w4.clone().requires_grad_(True)
        bg = bias.clone().requires_grad_(True)

        def fused_train(xin):
            def run():
                xin.grad = wg.grad = bg.grad = None
                kps, _, _ = ops.conv_integral_head_train(xin, wg, bg, K, NH, NS)
                kps.backward(gk)
            return run

        def unfused_train():
            xg.grad = wg.grad = bg.grad = None
            logits = F.conv2d(xg, wg, bg)
            kps, _, _ = ops.integral_multi_head(logits, K, NH, NS)
            kps.backward(gk)
        n1 = ops.launch_count()
        ms_ft = timeit(fused_train(xg), n=10)
        per_step = (ops.launch_count() - n1) // 13
        ms_ft_cl = timeit(fused_train(xclg), n=10)
        ms_ut = timeit(unfused_train, n=5)
        wbg = wb.clone().requires_grad_(True)
        bbg = bias.bfloat16().clone().requires_grad_(True)

        def unfused_best_train():
            xclg.grad = wbg.grad = bbg.grad = None
            xf = xclg.permute(0, 2, 3, 1).reshape(B, D * D, C).transpose(1, 2)
            logits = torch.baddbmm(bbg.view(1, -1, 1), wbg.unsqueeze(0).expand(B, K * D, C), xf)
            kps, _, _ = ops.integral_multi_head(logits.view(B, K * D, D, D), K, NH, NS)
            kps.backward(gk)
        ms_ubt = timeit(unfused_best_train, n=5)
        out["train_fwd_bwd"] = {"fused_from_nchw_fp32_ms": round(ms_ft, 3), "fused_from_nhwc_bf16_ms": round(ms_ft_cl, 3),
                                "unfused_cudnn_conv_plus_streaming_head_ms": round(ms_ut, 3), "speedup": round(ms_ut / ms_ft, 2),
                                "unfused_bf16_bmm_autograd_plus_streaming_head_bf16_ms": round(ms_ubt, 3),
                                "speedup_vs_best_unfused_bf16": round(ms_ubt / ms_ft_cl, 2),
                                "xsup_launches_per_step": int(per_step),
                                "note": "fused: pack, conv_head_fwd, coef, conv_rowcoef, conv_head_bwd dW + dX launches (all ours, tcgen05); "
                                        "unfused: cuDNN conv fwd/bwd (TF32) + integral_fwd/bwd on 4.56 GB of fp32 logits"}
        # ---- operand precision: the reference's conv is fp32 (TF32 on this GPU unless disabled); ours rounds x and W to bf16.
        # Coordinates (normalised to [-1, 1]; one heat-map bin = 2/64 = 0.031) against a true-fp32 conv, for two logit scales.
        prec = {}
        Bp = min(B, 16)
        for name, scale in (("logit_std_1", 1.0), ("logit_std_4", 4.0)):
            ws = (w * scale).contiguous()
            old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
            ref, _, ridx = ops.integral_multi_head(F.conv2d(x[:Bp], ws.view(K * D, C, 1, 1), bias), K, NH, NS)
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
            tf, _, tidx = ops.integral_multi_head(F.conv2d(x[:Bp], ws.view(K * D, C, 1, 1), bias), K, NH, NS)
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
            ours, _, oidx = ops.conv_integral_head(x[:Bp], ws, bias, K, NH, NS)
            ours32, _, o32idx = ops.conv_integral_head(x[:Bp], ws, bias, K, NH, NS, precision="tf32")
            same_t, same_o = (tidx == ridx).all(dim=-1), (oidx == ridx).all(dim=-1)       # [B,K]: all NH peak bins agree
            same_o32 = (o32idx == ridx).all(dim=-1)

            def err(a, mask):
                d = (a - ref).abs()[..., :2]                                            # x, y: defined whatever the peaks do
                dz = (a - ref).abs()[..., 2][mask.unsqueeze(1).expand(-1, NH, -1)]        # z where the same peaks were picked
                return {"xy_max": float(d.max()), "xy_mean": float(d.mean()), "z_max_same_peaks": float(dz.max()) if dz.numel() else None,
                        "units_with_same_peaks": float(mask.float().mean())}
            prec[name] = {"tf32_conv_vs_fp32": err(tf, same_t), "bf16_fused_vs_fp32": err(ours, same_o), "tf32_fused_vs_fp32": err(ours32, same_o32)}
        out["operand_precision"] = prec
        xcl32 = x.contiguous(memory_format=torch.channels_last)
        ms_tf32 = timeit(lambda: ops.conv_integral_head(xcl32, w, bias, K, NH, NS, precision="tf32"))
        out["fused_tf32"] = {"ms": round(ms_tf32, 4), "samples_per_s": round(B / ms_tf32 * 1e3, 1),
                             "note": "xsup_conv_head_fwd_tf32 from channels-last fp32 activations, including the round-to-nearest-tf32 pass over x and W"}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
