#!/usr/bin/env python
"""bench.py — integral + multi-hypothesis reprojection-loss fwd+bwd, samples/s and % of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path (head forward, geometry + loss + slot selection, backward to the
heat-map gradient) over one batch of synthetic input.  Workload at N=1: BASELINE.json configs[1]
(HM36_Multi_SurS1, batch 256, K=17, 64^3, fp32).  With N ranks the batch shards by sample (256 per GPU,
weak scaling); the only exchange is one all-reduce of the [4,NH] partial loss sums ('global' scope).

Prints ONE JSON line (rank 0).  `value` is timed with inputs resident in HBM; `e2e` goes through the
public API with pinned HOST buffers, H2D/D2H inside the timed region; `roofline` is the dominant
kernel (the streaming backward) timed live with CUDA events on its stream; `cpu_baseline` is the
oracle port timed on the host cores (rank 0, N=1).  `--impl reference` times that CPU port as the
reference arm.
"""
import argparse
import importlib
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "integral+reproj-loss fwd+bwd samples/sec"
UNIT = "samples/s"

CONFIGS = {
    # name: (per-GPU batch, K, R, NH, NS, dtype, loss weights, mpi cameras, description)
    "c2": dict(B=256, K=17, R=64, NH=3, NS=15, dtype="f32", w=(3.0, None, None, None), mpi=False,
               workload="HM36_Multi_SurS1 integral head + multi-hyp reprojection loss, batch 256/GPU, 17 joints, 64^3 fp32"),
    "c3": dict(B=256, K=17, R=64, NH=3, NS=15, dtype="bf16", w=(1.0, 0.1, 0.1, 0.0), mpi=False,
               workload="HM36_Multi_SynthS2 finetune-stage loss path, batch 256/GPU, 17 joints, 64^3 bf16 heatmaps"),
    "c4": dict(B=64, K=18, R=64, NH=3, NS=15, dtype="f32", w=(1.0, None, None, None), mpi=True,
               workload="MPI_Multi_SurS1 integral+reproj, batch 64/GPU, 18 joints, 64^3 fp32"),
}
CPU_SAMPLE_B = 32      # BASELINE.json configs[0]: the reference's own CPU-runnable case


def bytes_per_sample(c):
    return 3 * c["K"] * c["R"] ** 3 * (4 if c["dtype"] == "f32" else 2)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(kernel)
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------- CPU port (reference arm)
def cpu_step_fn(c, B, threads):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    oracle = importlib.import_module("xsup_oracle")
    synth = importlib.import_module("x-as-supervision_b200.synth")
    torch.set_num_threads(threads)
    K, R, NH, NS = c["K"], c["R"], c["NH"], c["NS"]
    logits = synth.iid_logits(B, K, R, R, R, seed=0)
    if c["dtype"] == "bf16":
        logits = logits.bfloat16().float()      # the reference has no bf16 path: fp32 math on bf16-rounded logits
    target = synth.pseudo_joints(B, K, seed=2)
    cams = synth.cameras(B, seed=3, mpi=c["mpi"])
    w = c["w"]

    def step():
        x = logits.clone().requires_grad_(True)
        lp, ls, *_ = oracle.fused_forward(x, K, NH, NS, target, cams, w_mse=w[0], w_bone=w[1], w_kp=w[2], w_kp2d=w[3],
                                          reduction="batch")
        (lp + ls).backward()
        return float((lp + ls).detach())
    return step


def time_cpu(c, steps, warmup, B=CPU_SAMPLE_B, min_seconds=0.0, max_steps=400):
    """Median step time of the CPU port over `steps` steps, continued until `min_seconds` of timed work (bounded by
    `max_steps`).  Returns (samples/s, threads, median seconds per step, steps timed)."""
    threads = os.cpu_count() or 1
    step = cpu_step_fn(c, B, threads)
    for _ in range(warmup):
        step()
    ts = []
    while len(ts) < steps or (sum(ts) < min_seconds and len(ts) < max_steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return B / statistics.median(ts), threads, statistics.median(ts), len(ts)


def run_reference(args, c):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = max(1, min(args.steps, 10))
    warmup = max(1, min(args.warmup, 2))
    v, threads, t, _ = time_cpu(c, steps, warmup)
    sample = "B=%d of the workload per step (%s, K=%d, %d^3, NH=%d), torch CPU fp32 port of the reference ops, %d steps" % (
        CPU_SAMPLE_B, c["dtype"], c["K"], c["R"], c["NH"], steps)
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 2), "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": round(t * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": c["dtype"], "data": "synthetic",
            "config": config_block(c, args.gpus, cpu=True),
            "cpu_baseline": {"value": round(v, 2), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": round(v, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def config_block(c, n, cpu=False, scope="global", exchange="none", launch="eager (one C-ABI call per kernel)"):
    return {"workload": c["workload"], "batch_per_gpu": c["B"] if not cpu else CPU_SAMPLE_B, "global_batch": c["B"] * n if not cpu else CPU_SAMPLE_B,
            "num_kp": c["K"], "heatmap": [c["R"]] * 3, "num_hypo": c["NH"], "neighbor_size": c["NS"],
            "loss_weights": {"mse": c["w"][0], "bone": c["w"][1], "kp": c["w"][2], "kp_2d": c["w"][3]},
            "reduction": "batch", "scope": scope if n > 1 else "local", "exchange": exchange,
            "parallelism": "sample-sharded x%d" % n, "launch": launch,
            "l2": "inputs (%.2f GB of logits per GPU) exceed the 126 MB L2; no explicit flush" % (
                (CPU_SAMPLE_B if cpu else c["B"]) * bytes_per_sample(c) / 3 / 1e9)}


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._halt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "clock sampling unavailable"}
        return {"sm_mhz": int(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------- GPU arm
def run_ours(args, c):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: xsup_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner there at VERSION/INFO level
        os.environ["NCCL_DEBUG"] = os.environ.get("XSUP_NCCL_DEBUG", "NONE")
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)                                   # anything NCCL prints while connecting goes to stderr
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
        group = None

    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        if rank == 0:
            ge.build()
        if world > 1:
            dist.barrier()
    pkg = importlib.import_module("x-as-supervision_b200")
    ops, synth = pkg.load_native(), pkg.synth
    exchange = "none"
    if world > 1 and args.scope == "global":
        if args.exchange == "nvlink":
            try:
                group = pkg.dist.PeerExchange(dist.group.WORLD, dev)
                exchange = "nvlink-p2p kernel (xsup_partial_allreduce)"
            except Exception as e:                      # no peer mapping on this box: say so and use NCCL
                sys.stderr.write("[bench] PeerExchange unavailable (%s); using NCCL all_reduce\n" % (e,))
        if group is None:
            group = dist.group.WORLD
            exchange = "nccl all_reduce"
    args.exchange_used = exchange

    B, K, R, NH, NS = c["B"], c["K"], c["R"], c["NH"], c["NS"]
    tdt = torch.float32 if c["dtype"] == "f32" else torch.bfloat16
    w = c["w"]
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    logits = torch.empty(B, K * R, R, R, device=dev, dtype=tdt)
    for i in range(0, B, 32):
        logits[i:i + 32] = torch.randn(min(32, B - i), K * R, R, R, device=dev, generator=gen).to(tdt)
    logits.requires_grad_(True)
    target = synth.pseudo_joints(B, K, seed=2 + rank).to(dev)
    cams = {k: v.to(dev) for k, v in synth.cameras(B, seed=3 + rank, mpi=c["mpi"]).items()}

    # per-kernel CUDA events on the launching stream (torch's current stream) around the two volume kernels
    ev = {"fwd": [], "bwd": [], "xchg": []}
    record = {"on": False}
    orig_fwd, orig_bwd = ops._head_forward, ops._head_backward

    def timed(kind, fn):
        def wrap(*a, **k):
            if not record["on"]:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            ev[kind].append((e0, e1))
            return out
        return wrap
    ops._head_forward, ops._head_backward = timed("fwd", orig_fwd), timed("bwd", orig_bwd)
    if group is not None:
        ops.xdist.reduce_partials = timed("xchg", ops.xdist.reduce_partials)

    def step():
        logits.grad = None
        lp, ls, sel, kps, world_, dmap, idx = ops.integral_reproj_min_loss(
            logits, target, cams, K, NH, NS, w_mse=w[0], w_bone=w[1], w_kp=w[2], w_kp2d=w[3], reduction="batch", group=group)
        (lp + ls).backward()
        return lp, ls, sel, kps

    graphed = None
    if args.graph:
        if group is not None and not isinstance(group, pkg.dist.PeerExchange):
            raise SystemExit("--graph with --scope global needs the NVLink exchange (--exchange nvlink), not NCCL")
        for _ in range(3):
            step()                                          # per-kernel events of the eager path (roofline block)
        torch.cuda.synchronize()
        record["on"] = True
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        record["on"] = False
        graphed = ops.GraphedReprojStep(logits, target, cams, K, NH, NS, w_mse=w[0], w_bone=w[1], w_kp=w[2], w_kp2d=w[3],
                                        reduction="batch", group=group)
        eager_step, launches_per_step = step, None
        n_a = ops.launch_count()
        eager_step()
        launches_per_step = ops.launch_count() - n_a
        step = graphed.__call__

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    # rank-0-only set-up goes BEFORE the barrier: every rank must enter the timed region together, otherwise the
    # first exchange of the other ranks waits for rank 0 and that wait is charged to the max-over-ranks time
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    sync_all()
    record["on"] = graphed is None
    n0 = ops.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step()
    t1.record()
    sync_all()
    launches = ops.launch_count() - n0
    if graphed is not None:
        launches = launches_per_step * args.steps            # replayed through cudaGraphLaunch: the same kernels, counted per eager step
    record["on"] = False
    clocks = sampler.finish() if sampler else None
    ms = t0.elapsed_time(t1) / args.steps
    k_fwd = statistics.mean(a.elapsed_time(b) for a, b in ev["fwd"])
    k_bwd = statistics.mean(a.elapsed_time(b) for a, b in ev["bwd"])
    k_x = [a.elapsed_time(b) for a, b in ev["xchg"]]
    if k_x:
        sys.stderr.write("[bench] rank %d: exchange (incl. waiting for peers) mean %.1f us, median %.1f us, max %.1f us; fwd %.4f ms bwd %.4f ms step %.4f ms\n"
                         % (rank, 1e3 * statistics.mean(k_x), 1e3 * statistics.median(k_x), 1e3 * max(k_x), k_fwd, k_bwd, ms))
    if world > 1:
        t = torch.tensor([ms, k_fwd, k_bwd], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, k_fwd, k_bwd = t.tolist()
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    value = B * world / (ms * 1e-3)

    # ---- the same step replayed as ONE CUDA graph (extra information; `value` above stays the eager number so that the
    # per-kernel events of the roofline block sit inside its timed region).  Single process only: a failure here must not
    # leave other ranks waiting in the exchange.
    args.graph_info = None
    if world == 1 and graphed is None:
        try:
            gstep = ops.GraphedReprojStep(logits, target, cams, K, NH, NS, w_mse=w[0], w_bone=w[1], w_kp=w[2], w_kp2d=w[3],
                                          reduction="batch")
            for _ in range(3):
                gstep()
            torch.cuda.synchronize()
            t0.record()
            for _ in range(args.steps):
                gstep()
            t1.record()
            torch.cuda.synchronize()
            gms = t0.elapsed_time(t1) / args.steps
            args.graph_info = {"value": round(B / (gms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(gms, 4),
                               "note": "ops.GraphedReprojStep: the step's launches captured once, one cudaGraphLaunch per step"}
            del gstep
        except Exception as e:                                   # never fail the bench line over the extra measurement
            args.graph_info = {"error": str(e)[:200]}

    # ---- end to end through the public API with pinned host buffers (H2D of every input, D2H of the results)
    if args.no_e2e:
        return finish(args, c, rank, world, dev, B, K, R, NH, ms, k_fwd, k_bwd, value, launches, clocks, None)
    e2e_steps = max(2, min(args.steps, 5))
    h_logits = torch.empty(logits.shape, dtype=tdt, pin_memory=True)
    h_logits.copy_(logits.detach())
    h_target = target.cpu().pin_memory()
    h_cams = {k: v.cpu().pin_memory() for k, v in cams.items()}
    h_out = {"loss": torch.empty(2, pin_memory=True), "sel": torch.empty(2, dtype=torch.int64, pin_memory=True),
             "kps": torch.empty(B, NH, K, 3, pin_memory=True)}
    d_logits = torch.empty_like(logits).requires_grad_(True)
    d_target = torch.empty_like(target)
    d_cams = {k: torch.empty_like(v) for k, v in cams.items()}

    def e2e_step():
        with torch.no_grad():
            d_logits.copy_(h_logits, non_blocking=True)
            d_target.copy_(h_target, non_blocking=True)
            for k in d_cams:
                d_cams[k].copy_(h_cams[k], non_blocking=True)
        d_logits.grad = None
        lp, ls, sel, kps, *_ = ops.integral_reproj_min_loss(d_logits, d_target, d_cams, K, NH, NS, w_mse=w[0], w_bone=w[1],
                                                            w_kp=w[2], w_kp2d=w[3], reduction="batch", group=group)
        (lp + ls).backward()
        h_out["loss"].copy_(torch.stack((lp.detach(), ls.detach())), non_blocking=True)
        h_out["sel"].copy_(sel, non_blocking=True)
        h_out["kps"].copy_(kps.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the caller reads the loss every step

    e2e_step()
    sync_all()
    t0.record()
    for _ in range(e2e_steps):
        e2e_step()
    t1.record()
    sync_all()
    e2e_ms = t0.elapsed_time(t1) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    h2d = h_logits.numel() * h_logits.element_size() + h_target.numel() * 4 + sum(v.numel() * 4 for v in h_cams.values())
    d2h = 2 * 4 + 2 * 8 + h_out["kps"].numel() * 4

    e2e = {"value": round(B * world / (e2e_ms * 1e-3), 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d) * world,
           "d2h_bytes_per_step": int(d2h) * world, "bytes_are": "whole job (all %d ranks)" % world, "ms_per_step": round(e2e_ms, 3),
           "steps": e2e_steps,
           "note": "pinned host -> device copy of logits/target/cameras, fused op fwd+bwd, loss/sel/kps read back; PCIe-bound"}
    return finish(args, c, rank, world, dev, B, K, R, NH, ms, k_fwd, k_bwd, value, launches, clocks, e2e)


def finish(args, c, rank, world, dev, B, K, R, NH, ms, k_fwd, k_bwd, value, launches, clocks, e2e):
    import torch.distributed as dist
    if rank == 0:
        peak, peak_src = measured_peak()
        unit_bytes = K * R ** 3 * (4 if c["dtype"] == "f32" else 2) * B          # one pass over this rank's volume
        bwd_gbs = 2 * unit_bytes / (k_bwd * 1e-3) / 1e9
        fwd_gbs = unit_bytes / (k_fwd * 1e-3) / 1e9
        step_gbs = 3 * unit_bytes / (ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": c["dtype"], "data": "synthetic", "config": config_block(c, world, scope=args.scope, exchange=getattr(args, "exchange_used", "none"),
                                                                                   launch="cuda graph replay" if args.graph else "eager (one C-ABI call per kernel)"),
                "roofline": {"bound": "hbm", "kernel": "integral_bwd_kernel (read logits + write grad, 2 passes)",
                             "achieved": round(bwd_gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(bwd_gbs / peak, 4),
                             "traffic": ncu_traffic("integral_bwd_kernel"), "peak_source": peak_src,
                             "ms_per_launch": round(k_bwd, 4), "share_of_step": round(k_bwd / ms, 4),
                             "fwd_kernel": {"achieved": round(fwd_gbs, 1), "frac": round(fwd_gbs / peak, 4),
                                            "ms_per_launch": round(k_fwd, 4), "share_of_step": round(k_fwd / ms, 4),
                                            "traffic": ncu_traffic("integral_fwd_kernel")},
                             "whole_step": {"achieved": round(step_gbs, 1), "frac": round(step_gbs / peak, 4),
                                            "frac_of_8TBs": round(step_gbs / 8000.0, 4)}},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        if getattr(args, "graph_info", None):
            line["cuda_graph_replay"] = args.graph_info
        if world == 1 and not args.no_cpu:
            v, threads, t, n_cpu = time_cpu(c, 3, 1, min_seconds=10.0)
            line["cpu_baseline"] = {"value": round(v, 2), "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "B=%d (BASELINE configs[0]) of the same workload, torch CPU fp32 port of the reference ops, "
                                              "median of %d steps (10 s of CPU work) after 1 warm-up, %.2f s per step" % (CPU_SAMPLE_B, n_cpu, t)}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", choices=sorted(CONFIGS), default="c2")
    ap.add_argument("--batch", type=int, default=None, help="override the per-GPU batch (sweeps)")
    ap.add_argument("--scope", choices=["global", "local"], default="global",
                    help="N>1: 'global' all-reduces the [4,NH] partial sums (single-process semantics on the global batch); "
                         "'local' selects per rank like the reference under DDP (no collective)")
    ap.add_argument("--exchange", choices=["nvlink", "nccl"], default="nvlink",
                    help="transport of the global-scope all-reduce: in-kernel NVLink peer-memory exchange, or torch NCCL")
    ap.add_argument("--graph", action="store_true",
                    help="replay the step as one CUDA graph (rank-local selection only); per-kernel events are then unavailable, "
                         "the roofline block reports the whole step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (profiling runs)")
    args = ap.parse_args()
    c = dict(CONFIGS[args.config])
    if args.batch:
        c["B"] = args.batch
    if args.impl == "reference":
        return run_reference(args, c)
    return run_ours(args, c)


if __name__ == "__main__":
    sys.exit(main())
