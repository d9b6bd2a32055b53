#!/usr/bin/env python
"""Throughput sweep of the integral + multi-hypothesis reprojection-loss path on one B200
(BASELINE.json configs[4]: heat-map resolution 32/64/128, hypotheses 1-16, batch 64-4096).

    python tools/sweep.py [--out gpurun_out/sweep.csv] [--dtype f32|bf16] [--quick]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/sweep.py --quick --scope global            # the same grid with `batch` samples on EACH of the 8 GPUs

Under torchrun every rank runs the grid on its own shard (weak scaling, the way bench.py scales), the step is timed with
a barrier on both sides, the time is the max over ranks and samples/s is the whole job's.  `--scope global` adds the
path's one exchange step (the NVLink mailbox all-reduce of the [4,NH] partial sums, `dist.PeerExchange`).

One row per (resolution, NH, batch): fwd+bwd samples/s, ms/step and the fraction of the HBM roofline
(algorithmic bytes = 3 * K * R^3 * sizeof per sample over the measured copy bandwidth of MEASURED_PEAKS.json
and over 8 TB/s).  Batches whose logits + gradient exceed --mem-gb are skipped and say so.  Same timing rules
as bench.py: >= 3 warm-up steps, CUDA events on the launching stream, inputs larger than L2 except the rows
flagged `l2_resident` (logits < 126 MB), which are reported but are not HBM numbers.
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
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.csv"))
    ap.add_argument("--dtype", choices=["f32", "bf16"], default="f32")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--mem-gb", type=float, default=150.0)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay each step as one CUDA graph (ops.GraphedReprojStep)")
    ap.add_argument("--res", type=int, nargs="*", default=None)
    ap.add_argument("--batches", type=int, nargs="*", default=None)
    ap.add_argument("--hypos", type=int, nargs="*", default=None)
    ap.add_argument("--scope", choices=["local", "global"], default="local")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("x-as-supervision_b200")
    ops, synth = pkg.load_native(), pkg.synth
    world, rank, lrank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", lrank)
    torch.cuda.set_device(dev)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        if args.scope == "global":
            group = pkg.dist.PeerExchange(dist.group.WORLD, dev)
    K, NS = 17, 15
    tdt = torch.float32 if args.dtype == "f32" else torch.bfloat16
    es = 4 if args.dtype == "f32" else 2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    res_list = args.res or [32, 64, 128]
    nh_list = args.hypos or ([1, 3, 16] if args.quick else [1, 2, 3, 4, 8, 16])
    b_list = args.batches or ([64, 1024] if args.quick else [64, 256, 1024, 4096])
    rows = ["res,num_hypo,batch_per_gpu,dtype,ms_per_step,samples_per_s,gbs_per_gpu,frac_of_measured_%.0f,frac_of_8TBs,note" % peak]
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for R in res_list:
        for B in b_list:
            vol = B * K * R ** 3 * es
            if 2 * vol > args.mem_gb * 1e9:
                rows.append("%d,*,%d,%s,,,,,,skipped: logits + gradient = %.0f GB" % (R, B, args.dtype, 2 * vol / 1e9))
                continue
            logits = torch.empty(B, K * R, R, R, device=dev, dtype=tdt)
            step_b = max(1, (1 << 28) // (K * R ** 3))
            for i in range(0, B, step_b):
                logits[i:i + step_b] = torch.randn(min(step_b, B - i), K * R, R, R, device=dev, generator=gen).to(tdt)
            logits.requires_grad_(True)
            target = synth.pseudo_joints(B, K, seed=2 + 10 * rank).to(dev)
            cams = {k: v.to(dev) for k, v in synth.cameras(B, seed=3 + 10 * rank).items()}
            for NH in nh_list:
                if NH > R - 2:
                    continue

                def step():
                    logits.grad = None
                    lp, ls, *_ = ops.integral_reproj_min_loss(logits, target, cams, K, NH, NS, w_mse=1.0, w_bone=0.1, w_kp=0.1,
                                                              w_kp2d=0.0, reduction="batch", group=group)
                    (lp + ls).backward()
                if args.graph:
                    g = ops.GraphedReprojStep(logits, target, cams, K, NH, NS, w_mse=1.0, w_bone=0.1, w_kp=0.1, w_kp2d=0.0,
                                              reduction="batch", group=group)
                    step = g.__call__
                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                    torch.cuda.synchronize()
                t0.record()
                for _ in range(args.steps):
                    step()
                t1.record()
                torch.cuda.synchronize()
                ms = t0.elapsed_time(t1) / args.steps
                if world > 1:
                    tm = torch.tensor([ms], device=dev)
                    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                    ms = float(tm.item())
                gbs = 3 * vol / (ms * 1e-3) / 1e9
                note = ("l2_resident" if vol < 126e6 else "") + (" graph" if args.graph else "") + \
                    ((" n_gpus=%d scope=%s" % (world, args.scope)) if world > 1 else "")
                if args.graph:
                    del g
                rows.append("%d,%d,%d,%s,%.4f,%.1f,%.1f,%.4f,%.4f,%s" % (R, NH, B, args.dtype, ms, world * B / ms * 1e3, gbs, gbs / peak, gbs / 8000.0, note))
                if rank == 0:
                    print(rows[-1], flush=True)
            del logits
            torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            f.write("\n".join(rows) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
