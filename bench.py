#!/usr/bin/env python
"""bench.py — integral + multi-hypothesis reprojection-loss fwd+bwd, samples/s and % of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path (head forward, geometry + loss + slot selection, backward to the
heat-map gradient) over one batch of synthetic input.  Workload at N=1: BASELINE.json configs[1]
(HM36_Multi_SurS1, batch 256, K=17, 64^3, fp32).  With N ranks the batch shards by sample (256 per GPU,
weak scaling); the only exchange is one all-reduce of the [4,NH] partial loss sums ('global' scope), done
inside the loss kernel over NVLink peer memory.

Order of the run (rank 0 prints ONE JSON line at the end):
  1. parity gate — before any timing: at N=1 the CUDA path against the CPU port on the cpu_baseline's own sample
     (slots exact, loss / coordinates / gradient <= 1e-5), plus eager == CUDA-graph replay; at N>1 the sharded
     global-scope step against a single-GPU run on the gathered batch (slots exact, loss and gradient <= 1e-6).
     A failed gate prints the line with "parity_gate": {"ok": false, ...} and exits non-zero.
  2. `value`: K steps timed with CUDA events, inputs resident in HBM; `roofline`: the dominant kernel (streaming
     backward) timed live with CUDA events on its stream inside the same region.
  3. extras: the same step as one CUDA graph; `sustained` (>= 2 s of back-to-back steps, clocks recorded);
     `configs`: BASELINE configs[2] (bf16, SynthS2, batch 1024 / N per GPU) and configs[3] (MPI, K=18, batch 512 / N
     per GPU) at this N, eager and as a CUDA graph.
  4. `e2e`: the public API with pinned HOST buffers, H2D/D2H inside the timed region, next to the bare H2D time
     of the same bytes (the PCIe ceiling of this box at this N).
  5. `cpu_baseline`: the oracle port timed on the host cores (rank 0, N=1).
`--impl reference` times that CPU port as the reference arm (the reference is pure Python and is not installable
on the GPU box).
"""
import argparse
import importlib
import json
import math
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
    # B = per-GPU batch of the default (weak-scaling) run; GB = BASELINE's global batch where it names one
    "c2": dict(B=256, GB=None, K=17, R=64, NH=3, NS=15, dtype="f32", w=(3.0, None, None, None), mpi=False,
               workload="HM36_Multi_SurS1 integral head + multi-hyp reprojection loss, batch 256/GPU, 17 joints, 64^3 fp32"),
    "c3": dict(B=256, GB=1024, K=17, R=64, NH=3, NS=15, dtype="bf16", w=(1.0, 0.1, 0.1, 0.0), mpi=False,
               workload="HM36_Multi_SynthS2 finetune-stage loss path, global batch 1024, 17 joints, 64^3 bf16 heatmaps"),
    "c4": dict(B=64, GB=512, K=18, R=64, NH=3, NS=15, dtype="f32", w=(1.0, None, None, None), mpi=True,
               workload="MPI_Multi_SurS1 integral+reproj, global batch 512, 18 joints, 64^3 fp32"),
}
CPU_SAMPLE_B = 32      # BASELINE.json configs[0]: the reference's own CPU-runnable case
K1, K2F, K2B, K3 = "xsup.K1.integral_fwd", "xsup.K2.loss_select_fwd", "xsup.K2.loss_bwd_coef", "xsup.K3.integral_bwd"


def esize(c):
    return 4 if c["dtype"] == "f32" else 2


def bytes_per_sample(c):
    return 3 * c["K"] * c["R"] ** 3 * esize(c)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(config, kernel, B):
    """dram read + write bytes per launch from the committed `ncu --set full` capture of THIS config at THIS per-GPU
    batch (profiles/ncu_traffic.json), else None: a figure from another shape would be a constant, not a measurement."""
    try:
        e = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[config]
        return e[kernel] if int(e.get("batch", -1)) == int(B) else None
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------- CPU port (reference arm)
def cpu_inputs(c, B):
    synth = importlib.import_module("x-as-supervision_b200.synth")
    logits = synth.iid_logits(B, c["K"], c["R"], c["R"], c["R"], seed=0)
    if c["dtype"] == "bf16":
        logits = logits.bfloat16().float()      # the reference has no bf16 path: fp32 math on bf16-rounded logits
    return logits, synth.pseudo_joints(B, c["K"], seed=2), synth.cameras(B, seed=3, mpi=c["mpi"])


def cpu_step_fn(c, B, threads):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    oracle = importlib.import_module("xsup_oracle")
    torch.set_num_threads(threads)
    K, NH, NS = c["K"], c["NH"], c["NS"]
    logits, target, cams = cpu_inputs(c, B)
    w = c["w"]
    keep = {}

    def step():
        x = logits.clone().requires_grad_(True)
        lp, ls, sel, kps, *_ = oracle.fused_forward(x, K, NH, NS, target, cams, w_mse=w[0], w_bone=w[1], w_kp=w[2], w_kp2d=w[3],
                                                    reduction="batch")
        (lp + ls).backward()
        keep.update(lp=float(lp.detach()), ls=float(ls.detach()), sel=sel.tolist(), kps=kps.detach(), grad=x.grad)
        return float((lp + ls).detach())
    return step, keep


def time_cpu(c, steps, warmup, B=CPU_SAMPLE_B, min_seconds=0.0, max_steps=400):
    """Median step time of the CPU port over `steps` steps, continued until `min_seconds` of timed work (bounded by
    `max_steps`).  Returns (samples/s, threads, median seconds per step, steps timed, outputs of the last step)."""
    threads = os.cpu_count() or 1
    step, keep = cpu_step_fn(c, B, threads)
    for _ in range(warmup):
        step()
    ts = []
    while len(ts) < steps or (sum(ts) < min_seconds and len(ts) < max_steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return B / statistics.median(ts), threads, statistics.median(ts), len(ts), keep


def run_reference(args, c):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    v, threads, t, n, _ = time_cpu(c, steps, warmup)
    sample = ("each step = B=%d samples of the workload (%s, K=%d, %d^3, NH=%d; BASELINE configs[0] size) through the torch CPU fp32 port of "
              "the reference's op sequence on %d threads; median of %d steps after %d warm-up; samples/s normalises the batch"
              % (CPU_SAMPLE_B, c["dtype"], c["K"], c["R"], c["NH"], threads, n, warmup))
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 2), "unit": UNIT, "n_gpus": args.gpus, "steps": n,
            "warmup": warmup, "ms_per_step": round(t * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": c["dtype"], "data": "synthetic",
            "config": config_block(c, args.gpus, scope=args.scope),
            "run": {"where": "host cores", "threads": threads},
            "cpu_baseline": {"value": round(v, 2), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": round(v, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def config_block(c, n, scope="global"):
    """The workload only (identical for both arms); how this arm ran it goes into the line's `run` block."""
    return {"workload": c["workload"], "batch_per_gpu": c["B"], "global_batch": c["B"] * n,
            "num_kp": c["K"], "heatmap": [c["R"]] * 3, "num_hypo": c["NH"], "neighbor_size": c["NS"],
            "loss_weights": {"mse": c["w"][0], "bone": c["w"][1], "kp": c["w"][2], "kp_2d": c["w"][3]},
            "reduction": "batch", "scope": scope if n > 1 else "local",
            "parallelism": "sample-sharded x%d" % n,
            "l2": "inputs (%.2f GB of logits per GPU) exceed the 126 MB L2; no explicit flush" % (
                c["B"] * bytes_per_sample(c) / 3 / 1e9)}


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
class Ctx:
    """Process-wide state of the GPU arm (rank, device, package handles, exchange)."""


def setup(args):
    import torch
    import torch.distributed as dist
    x = Ctx()
    x.torch, x.dist = torch, dist
    x.rank = int(os.environ.get("RANK", "0"))
    x.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    x.world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: xsup_b200 has no CPU path")
    torch.cuda.set_device(x.local_rank)
    x.dev = torch.device("cuda", x.local_rank)
    if x.world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner there at VERSION/INFO level
        os.environ["NCCL_DEBUG"] = os.environ.get("XSUP_NCCL_DEBUG", "NONE")
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)                                   # anything NCCL prints while connecting goes to stderr
        try:
            dist.init_process_group("nccl", device_id=x.dev)
            dist.barrier()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        if x.rank == 0:
            ge.build()
        if x.world > 1:
            dist.barrier()
    x.numa = bind_to_gpu_numa(x)
    x.pkg = importlib.import_module("x-as-supervision_b200")
    x.ops, x.synth = x.pkg.load_native(), x.pkg.synth
    x.group, x.exchange = None, "none"
    if x.world > 1 and args.scope == "global":
        if args.exchange == "nvlink":
            try:
                x.group = x.pkg.dist.PeerExchange(dist.group.WORLD, x.dev)
                x.exchange = "nvlink-p2p, inside the loss kernel (xsup_reproj_fused_fwd)"
            except Exception as e:                      # no peer mapping on this box: say so and use NCCL
                sys.stderr.write("[bench] PeerExchange unavailable (%s); using NCCL all_reduce\n" % (e,))
        if x.group is None:
            x.group = dist.group.WORLD
            x.exchange = "nccl all_reduce between two launches"
    return x


def bind_to_gpu_numa(x):
    """Pin this rank to the CPUs local to its GPU (sysfs `local_cpulist` of the PCI device) BEFORE any pinned host buffer is
    allocated: pinned pages are placed by first touch, so the end-to-end copies then read NUMA-local memory.  Returns what it
    did for the bench line; a single-node box (or a container without sysfs) is reported as such and left alone."""
    info = {"bound": False}
    try:
        pr = x.torch.cuda.get_device_properties(x.dev)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read().strip())
        cpus_txt = open(base + "/local_cpulist").read().strip()
        info.update({"pci": bdf, "numa_node": node, "local_cpulist": cpus_txt})
        cpus = set()
        for part in cpus_txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["numa_nodes"] = len(nodes)
        if cpus and (cpus & allowed) and len(nodes) > 1 and (cpus & allowed) != allowed:
            os.sched_setaffinity(0, cpus & allowed)
            info["bound"] = True
    except Exception as e:                                   # no sysfs / no permission: measure unbound
        info["note"] = str(e)[:80]
    return info


def sync_all(x):
    if x.world > 1:
        x.dist.barrier()
    x.torch.cuda.synchronize()


def max_over_ranks(x, vals):
    if x.world == 1:
        return list(vals)
    t = x.torch.tensor(list(vals), device=x.dev, dtype=x.torch.float64)
    x.dist.all_reduce(t, op=x.dist.ReduceOp.MAX)
    return t.tolist()


def make_inputs(x, c, B):
    torch = x.torch
    tdt = torch.float32 if c["dtype"] == "f32" else torch.bfloat16
    gen = torch.Generator(device=x.dev).manual_seed(1234 + x.rank)
    logits = torch.empty(B, c["K"] * c["R"], c["R"], c["R"], device=x.dev, dtype=tdt)
    for i in range(0, B, 32):
        logits[i:i + 32] = torch.randn(min(32, B - i), c["K"] * c["R"], c["R"], c["R"], device=x.dev, generator=gen).to(tdt)
    logits.requires_grad_(True)
    target = x.synth.pseudo_joints(B, c["K"], seed=2 + x.rank).to(x.dev)
    cams = {k: v.to(x.dev) for k, v in x.synth.cameras(B, seed=3 + x.rank, mpi=c["mpi"]).items()}
    return logits, target, cams


def make_step(x, c, logits, target, cams, group):
    w, K, NH, NS = c["w"], c["K"], c["NH"], c["NS"]

    def step():
        logits.grad = None
        lp, ls, sel, kps, world_, dmap, idx = x.ops.integral_reproj_min_loss(
            logits, target, cams, K, NH, NS, w_mse=w[0], w_bone=w[1], w_kp=w[2], w_kp2d=w[3], reduction="batch", group=group)
        (lp + ls).backward()
        return lp, ls, sel, kps
    return step


def make_graph(x, c, logits, target, cams, group):
    if group is not None and not isinstance(group, x.pkg.dist.PeerExchange):
        return None                                      # an NCCL process group is not captured
    w = c["w"]
    return x.ops.GraphedReprojStep(logits, target, cams, c["K"], c["NH"], c["NS"], w_mse=w[0], w_bone=w[1], w_kp=w[2], w_kp2d=w[3],
                                   reduction="batch", group=group)


def time_steps(x, step, steps, warmup):
    """`steps` calls bracketed by barrier + synchronize, CUDA events on the launching stream; ms per step, max over ranks."""
    torch = x.torch
    for _ in range(warmup):
        step()
    sync_all(x)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step()
    t1.record()
    sync_all(x)
    return max_over_ranks(x, [t0.elapsed_time(t1) / steps])[0]


def time_steps_long(x, step, steps, warmup, min_ms=150.0, max_steps=400):
    """For the sub-millisecond extras: a first pass of `steps` gives the step time (max over ranks, so every rank derives the
    same count), then enough steps to fill `min_ms` - one 2 ms host hiccup on any of 8 ranks inside a 12 ms region moved a line
    by 17 %.  Returns (ms per step, steps timed)."""
    est = time_steps(x, step, steps, warmup)
    n = int(min(max_steps, max(steps, min_ms / max(est, 1e-3))))
    if n == steps:
        return est, steps
    return time_steps(x, step, n, 0), n


def time_prefetched(x, steps, copy_in, compute, nbuf=2):
    """The end-to-end step as an input pipeline: two device input sets, the host->device copies of step s+1 issued on a second
    stream before the kernels of step s are launched, so the PCIe transfer and the compute overlap (events `ready` / `free` order
    the two streams per buffer).  Every step still copies its inputs from pinned host memory and reads its results back inside
    the timed region - the first copy is exposed, the others hide behind compute or vice versa.  ms per step, max over ranks."""
    torch = x.torch
    cur = torch.cuda.current_stream()
    cs = torch.cuda.Stream(device=x.dev)
    ready = [torch.cuda.Event() for _ in range(nbuf)]
    free = [torch.cuda.Event() for _ in range(nbuf)]

    def issue(i):
        cs.wait_event(free[i])                               # the step that last read buffer i has finished (no-op before its first use)
        with torch.cuda.stream(cs):
            copy_in(i)
            ready[i].record(cs)

    def run(n):
        issue(0)
        for st in range(n):
            i = st % nbuf
            if st + 1 < n:
                issue((st + 1) % nbuf)
            cur.wait_event(ready[i])
            compute(i)
            free[i].record(cur)
            cur.synchronize()                                # the caller reads the loss every step
    run(2)
    sync_all(x)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    run(steps)
    t1.record()
    sync_all(x)
    return max_over_ranks(x, [t0.elapsed_time(t1) / steps])[0]


# ---- 1. parity gate
def rel_inf(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-300))


def gate_vs_cpu_port(x, c):
    """N=1, rank 0: the CUDA path against the CPU port on the cpu_baseline's own sample (the checker half of the
    cpu_baseline leg; the port is never on the measured GPU path).  Also times the port: returns (gate, cpu_baseline)."""
    torch = x.torch
    v, threads, t, n_cpu, ref = time_cpu(c, 3, 1, min_seconds=10.0)
    logits, target, cams = cpu_inputs(c, CPU_SAMPLE_B)
    tdt = torch.float32 if c["dtype"] == "f32" else torch.bfloat16
    xl = logits.to(x.dev).to(tdt).requires_grad_(True)
    w = c["w"]
    lp, ls, sel, kps, *_ = x.ops.integral_reproj_min_loss(xl, target.to(x.dev), {k: q.to(x.dev) for k, q in cams.items()}, c["K"], c["NH"],
                                                          c["NS"], w_mse=w[0], w_bone=w[1], w_kp=w[2], w_kp2d=w[3], reduction="batch")
    (lp + ls).backward()
    torch.cuda.synchronize()
    e = {"slots_equal": sel.tolist() == ref["sel"],
         "loss_rel": abs(float((lp + ls).detach()) - (ref["lp"] + ref["ls"])) / max(abs(ref["lp"] + ref["ls"]), 1e-30),
         "kps_rel": rel_inf(kps.detach().cpu(), ref["kps"]),
         "grad_rel_inf": rel_inf(xl.grad.float().cpu(), ref["grad"])}
    # the comparator is the reference's own fp32 CPU arithmetic (softmax accurate to ~1e-5 per element)
    tol_grad = 1e-5 if c["dtype"] == "f32" else 2.0 ** -8
    e["ok"] = bool(e["slots_equal"] and e["loss_rel"] < 1e-5 and e["kps_rel"] < 1e-5 and e["grad_rel_inf"] < tol_grad)
    e["what"] = "CUDA path vs the fp32 CPU port of the reference ops, B=%d of the workload: slots exact, loss/kps 1e-5, grad %.0e (norm-wise)" % (
        CPU_SAMPLE_B, tol_grad)
    for k in ("loss_rel", "kps_rel", "grad_rel_inf"):
        e[k] = float("%.3e" % e[k])
    cpu = {"value": round(v, 2), "unit": UNIT, "cores": threads, "kind": "port",
           "sample": "B=%d (BASELINE configs[0]) of the same workload, torch CPU fp32 port of the reference ops, "
                     "median of %d steps (10 s of CPU work) after 1 warm-up, %.2f s per step" % (CPU_SAMPLE_B, n_cpu, t)}
    return e, cpu


def gate_graph_equals_eager(x, c):
    """Single GPU: a CUDA-graph replay of the step is bit-identical to the eager step (small batch)."""
    torch = x.torch
    small = dict(c, R=32)
    logits, target, cams = make_inputs(x, small, 8)
    step = make_step(x, small, logits, target, cams, None)
    lp, ls, sel, kps = step()
    g_eager = logits.grad.clone()
    gs = make_graph(x, small, logits, target, cams, None)
    glp, gls, gsel = gs()
    torch.cuda.synchronize()
    ok = bool(torch.equal(glp, lp.detach()) and torch.equal(gls, ls.detach()) and torch.equal(gsel, sel) and torch.equal(gs.grad, g_eager))
    return {"ok": ok, "what": "CUDA-graph replay == eager step, bit for bit (B=8, 32^3)"}


def gate_multi_gpu(x, c):
    """N>1: the sharded global-scope step against a single-GPU run of the same code on the gathered batch, on every
    rank's own GPU: selected slots identical, loss and this rank's gradient shard within 1e-6."""
    torch, dist = x.torch, x.dist
    small = dict(c, R=32)
    Bs = 8
    logits, target, cams = make_inputs(x, small, Bs)
    step = make_step(x, small, logits, target, cams, x.group)
    lp, ls, sel, kps = step()

    def gather(t):
        out = [torch.empty_like(t) for _ in range(x.world)]
        dist.all_gather(out, t.detach().contiguous())
        return torch.cat(out, 0)
    all_logits = gather(logits).requires_grad_(True)
    all_target = gather(target)
    all_cams = {k: gather(v) for k, v in cams.items()}
    w = small["w"]
    rlp, rls, rsel, rkps, *_ = x.ops.integral_reproj_min_loss(all_logits, all_target, all_cams, small["K"], small["NH"], small["NS"],
                                                              w_mse=w[0], w_bone=w[1], w_kp=w[2], w_kp2d=w[3], reduction="batch", group=None)
    (rlp + rls).backward()
    shard = all_logits.grad[x.rank * Bs:(x.rank + 1) * Bs]
    tol = 1e-6 if c["dtype"] == "f32" else 2.0 ** -8
    e_loss = abs(float((lp + ls).detach()) - float((rlp + rls).detach())) / max(abs(float((rlp + rls).detach())), 1e-30)
    e_grad = rel_inf(logits.grad.float(), shard.float()) if float(shard.float().abs().max()) > 0 else float(logits.grad.float().abs().max())
    ok = sel.tolist() == rsel.tolist() and e_loss < 1e-6 and e_grad < tol and torch.equal(kps.detach(), rkps.detach()[x.rank * Bs:(x.rank + 1) * Bs])
    if isinstance(x.group, x.pkg.dist.PeerExchange):
        x.group.check()                                    # no exchange timed out
    flag = torch.tensor([1.0 if ok else 0.0, -e_loss, -e_grad], device=x.dev, dtype=torch.float64)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return {"ok": bool(flag[0].item() == 1.0), "slots": sel.tolist(), "loss_rel_max": float("%.3e" % -flag[1].item()),
            "grad_rel_max": float("%.3e" % -flag[2].item()),
            "what": "%d ranks x B=%d (32^3), scope=%s via %s vs one GPU on the gathered batch of %d: slots and kps exact, loss 1e-6, "
                    "each rank's gradient shard %.0e" % (x.world, Bs, "global", x.exchange, Bs * x.world, tol)}


# ---- 3. extras
def run_config(x, args, name, steps):
    """One of the other BASELINE configs at this N (its global batch split over the ranks): eager and CUDA-graph."""
    torch = x.torch
    c = dict(CONFIGS[name])
    B = max(1, c["GB"] // x.world)
    c["B"] = B
    out = {"workload": c["workload"], "global_batch": B * x.world, "batch_per_gpu": B, "dtype": c["dtype"], "num_kp": c["K"],
           "scaling": "strong (BASELINE's global batch split over the ranks)", "steps": steps}
    try:
        logits, target, cams = make_inputs(x, c, B)
        step = make_step(x, c, logits, target, cams, x.group)
        ms, n = time_steps_long(x, step, steps, 3)
        gbs = B * bytes_per_sample(c) / (ms * 1e-3) / 1e9
        out["eager"] = {"value": round(B * x.world / (ms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms, 4), "steps": n,
                        "per_gpu_gbs": round(gbs, 1), "frac_of_8TBs": round(gbs / 8000.0, 4)}
        gs = make_graph(x, c, logits, target, cams, x.group)
        if gs is not None:
            ms, n = time_steps_long(x, gs.__call__, steps, 3)
            gbs = B * bytes_per_sample(c) / (ms * 1e-3) / 1e9
            out["cuda_graph"] = {"value": round(B * x.world / (ms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms, 4), "steps": n,
                                 "per_gpu_gbs": round(gbs, 1), "frac_of_8TBs": round(gbs / 8000.0, 4)}
        del gs
        # end to end for this config as well: pinned host logits / target / cameras in, loss / slots / coordinates out
        h_logits = torch.empty(logits.shape, dtype=logits.dtype, pin_memory=True)
        h_logits.copy_(logits.detach())
        h_target = target.cpu().pin_memory()
        h_cams = {k: v.cpu().pin_memory() for k, v in cams.items()}
        h_out = {"loss": torch.empty(2, pin_memory=True), "sel": torch.empty(2, dtype=torch.int64, pin_memory=True),
                 "kps": torch.empty(B, c["NH"], c["K"], 3, pin_memory=True)}

        def e2e_step():
            with torch.no_grad():
                logits.copy_(h_logits, non_blocking=True)
                target.copy_(h_target, non_blocking=True)
                for k in cams:
                    cams[k].copy_(h_cams[k], non_blocking=True)
            lp, ls, sel, kps = step()
            h_out["loss"].copy_(torch.stack((lp.detach(), ls.detach())), non_blocking=True)
            h_out["sel"].copy_(sel, non_blocking=True)
            h_out["kps"].copy_(kps.detach(), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e_ms = time_steps(x, e2e_step, 3, 1)
        h2d = h_logits.numel() * h_logits.element_size() + h_target.numel() * 4 + sum(v.numel() * 4 for v in h_cams.values())
        out["e2e"] = {"value": round(B * x.world / (e_ms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(e_ms, 3),
                      "h2d_bytes_per_step": int(h2d) * x.world, "d2h_bytes_per_step": int(2 * 4 + 2 * 8 + h_out["kps"].numel() * 4) * x.world}
        del h_logits, h_out, step, logits, target, cams
    except torch.cuda.OutOfMemoryError as e:               # every rank sees the same sizes, so every rank lands here together
        out["error"] = "out of memory: %s" % (str(e)[:120],)
    torch.cuda.empty_cache()
    return out


def run_conv_fused(x, args, c, steps):
    """The same per-camera op with the head's final 1x1 conv (deconv_head.py:33-35) pulled in: ops.conv_integral_reproj_min_loss on
    bf16 channels-last activations [B, 256, 64, 64] - the logits never exist, so the host->device traffic of the end-to-end
    step is the activations (0.54 GB at B=256) instead of the logits (4.56 GB).  Device-resident timing, then the end-to-end variant with pinned host
    activations.  Tensor-core roofline: 3 GEMMs of 2*K*D*C*H*W flops per sample are the algorithm (conv fwd, d x, d W); the
    launches execute 5 (the backward recomputes the logit tiles in both of its launches)."""
    torch = x.torch
    B, K, R, NH, NS, C = c["B"], c["K"], c["R"], c["NH"], c["NS"], 256
    out = {"workload": "final Conv2d(%d, %d, 1) + integral head + multi-hyp reprojection loss, batch %d/GPU, activations bf16 channels-last" % (C, K * R, B),
           "batch_per_gpu": B, "channels": C, "steps": steps}
    try:
        gen = torch.Generator(device=x.dev).manual_seed(4321 + x.rank)
        feat = torch.randn(B, C, R, R, device=x.dev, generator=gen).to(dtype=torch.bfloat16, memory_format=torch.channels_last).requires_grad_(True)
        weight = (torch.randn(K * R, C, device=x.dev, generator=gen) / C ** 0.5).requires_grad_(True)
        bias = torch.randn(K * R, device=x.dev, generator=gen).requires_grad_(True)
        target = x.synth.pseudo_joints(B, K, seed=2 + x.rank).to(x.dev)
        cams = {k: v.to(x.dev) for k, v in x.synth.cameras(B, seed=3 + x.rank, mpi=c["mpi"]).items()}
        w = c["w"]
        group = x.group if isinstance(x.group, x.pkg.dist.PeerExchange) else None

        def step():
            feat.grad = weight.grad = bias.grad = None
            lp, ls, sel, kps, *_ = x.ops.conv_integral_reproj_min_loss(feat, weight, bias, target, cams, K, NH, NS, w_mse=w[0], w_bone=w[1],
                                                                       w_kp=w[2], w_kp2d=w[3], reduction="batch", group=group)
            (lp + ls).backward()
            return lp, ls, sel, kps
        n0 = x.ops.launch_count()
        step()
        out["launches_per_step"] = int(x.ops.launch_count() - n0)
        ms = time_steps(x, step, steps, 3)
        gemm = 2.0 * K * R * C * R * R * B
        try:
            peak_tf = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
            peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernels timed inside a step)"
        except Exception:
            peak_tf, peak_src = 1400.0, "fallback"
        out["device"] = {"value": round(B * x.world / (ms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms, 4),
                         "roofline": {"bound": "tensor", "achieved": round(3 * gemm / (ms * 1e-3) / 1e12, 1), "executed": round(5 * gemm / (ms * 1e-3) / 1e12, 1),
                                      "peak": peak_tf, "unit": "TFLOP/s", "frac": round(3 * gemm / (ms * 1e-3) / 1e12 / peak_tf, 4),
                                      "peak_source": peak_src,
                                      "note": "achieved = algorithmic flops (3 GEMMs) / step time; executed counts the 2 recomputed logit GEMMs of the backward too"}}
        # end to end: pinned host activations / target / cameras -> device, step, loss / slots / coordinates back
        h_feat = torch.empty(feat.shape, dtype=torch.bfloat16, pin_memory=True, memory_format=torch.channels_last)
        h_feat.copy_(feat.detach())
        h_target = target.cpu().pin_memory()
        h_cams = {k: v.cpu().pin_memory() for k, v in cams.items()}
        h_out = {"loss": torch.empty(2, pin_memory=True), "sel": torch.empty(2, dtype=torch.int64, pin_memory=True),
                 "kps": torch.empty(B, NH, K, 3, pin_memory=True)}

        def e2e_step():
            with torch.no_grad():
                feat.copy_(h_feat, non_blocking=True)
                target.copy_(h_target, non_blocking=True)
                for k in cams:
                    cams[k].copy_(h_cams[k], non_blocking=True)
            lp, ls, sel, kps = step()
            h_out["loss"].copy_(torch.stack((lp.detach(), ls.detach())), non_blocking=True)
            h_out["sel"].copy_(sel, non_blocking=True)
            h_out["kps"].copy_(kps.detach(), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e_ms = time_steps(x, e2e_step, steps, 2)
        h2d = h_feat.numel() * 2 + h_target.numel() * 4 + sum(v.numel() * 4 for v in h_cams.values())
        out["e2e"] = {"value": round(B * x.world / (e_ms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(e_ms, 3),
                      "h2d_bytes_per_step": int(h2d) * x.world, "d2h_bytes_per_step": int(2 * 4 + 2 * 8 + h_out["kps"].numel() * 4) * x.world,
                      "note": "activations (not logits) cross PCIe: 8.5x fewer bytes per sample than the logit-fed step; gradients d x / d W / d bias stay on the device"}
        try:                                                   # the same steps as a double-buffered input pipeline
            sets = [(feat, target, cams),
                    (torch.empty_like(feat).requires_grad_(True), torch.empty_like(target), {k: torch.empty_like(v) for k, v in cams.items()})]

            def copy_in(i):
                df, dt, dc = sets[i]
                with torch.no_grad():
                    df.copy_(h_feat, non_blocking=True)
                    dt.copy_(h_target, non_blocking=True)
                    for k in dc:
                        dc[k].copy_(h_cams[k], non_blocking=True)

            def compute(i):
                df, dt, dc = sets[i]
                df.grad = weight.grad = bias.grad = None
                lp, ls, sel, kps, *_ = x.ops.conv_integral_reproj_min_loss(df, weight, bias, dt, dc, K, NH, NS, w_mse=w[0], w_bone=w[1],
                                                                           w_kp=w[2], w_kp2d=w[3], reduction="batch", group=group)
                (lp + ls).backward()
                h_out["loss"].copy_(torch.stack((lp.detach(), ls.detach())), non_blocking=True)
                h_out["sel"].copy_(sel, non_blocking=True)
                h_out["kps"].copy_(kps.detach(), non_blocking=True)
            p_ms = time_prefetched(x, steps, copy_in, compute)
            out["e2e"]["prefetch"] = {"value": round(B * x.world / (p_ms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(p_ms, 3), "steps": steps,
                                      "note": "two device input sets; the copies of step s+1 run on a second stream under the kernels of step s"}
            del sets
        except RuntimeError as e:
            out["e2e"]["prefetch"] = {"error": str(e)[:200]}
    except torch.cuda.OutOfMemoryError as e:
        out["error"] = "out of memory: %s" % (str(e)[:120],)
    except RuntimeError as e:
        out["error"] = str(e)[:200]
    torch.cuda.empty_cache()
    return out


def run_ours(args, c):
    x = setup(args)
    torch, ops = x.torch, x.ops
    B, K, R, NH = c["B"], c["K"], c["R"], c["NH"]
    tdt = torch.float32 if c["dtype"] == "f32" else torch.bfloat16
    line_extra = {}

    # ---- 1. parity gate, before any timing
    gate = {"ok": True}
    cpu_baseline = None
    if x.world == 1:
        if not args.no_cpu:
            gate["vs_cpu_port"], cpu_baseline = gate_vs_cpu_port(x, c)
        else:
            gate["vs_cpu_port"] = {"skipped": "--no-cpu"}
        gate["graph_vs_eager"] = gate_graph_equals_eager(x, c)
    else:
        gate["multi_gpu"] = gate_multi_gpu(x, c)
    gate["ok"] = all(v.get("ok", True) for v in gate.values() if isinstance(v, dict))
    if not gate["ok"]:
        if x.rank == 0:
            print(json.dumps({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": x.world, "parity_gate": gate,
                              "error": "parity gate failed: nothing was timed"}), flush=True)
        if x.world > 1:
            x.dist.barrier()
            x.dist.destroy_process_group()
        return 3

    # ---- 2. the timed region
    logits, target, cams = make_inputs(x, c, B)
    step = make_step(x, c, logits, target, cams, x.group)
    graphed = None
    launches_per_step = None
    if args.graph:
        if x.group is not None and not isinstance(x.group, x.pkg.dist.PeerExchange):
            raise SystemExit("--graph with --scope global needs the NVLink exchange (--exchange nvlink), not NCCL")
    ev = {}
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    n_a = ops.launch_count()
    step()
    launches_per_step = ops.launch_count() - n_a
    eager_step = step
    if args.graph:
        ops.set_event_sink(ev)                              # per-kernel events of the eager path (roofline block)
        for _ in range(5):
            eager_step()
        torch.cuda.synchronize()
        ops.set_event_sink(None)
        graphed = make_graph(x, c, logits, target, cams, x.group)
        step = graphed.__call__

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step()
    # rank-0-only set-up goes BEFORE the barrier: every rank must enter the timed region together, otherwise the
    # first exchange of the other ranks waits for rank 0 and that wait is charged to the max-over-ranks time
    sampler = ClockSampler(x.local_rank) if x.rank == 0 else None
    if sampler:
        sampler.start()
    sync_all(x)
    if graphed is None:
        ops.set_event_sink(ev)
    n0 = ops.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step()
    t1.record()
    sync_all(x)
    ops.set_event_sink(None)
    launches = ops.launch_count() - n0
    if graphed is not None:
        launches = launches_per_step * args.steps            # replayed through cudaGraphLaunch: the same kernels, counted per eager step
    clocks = sampler.finish() if sampler else None
    ms = t0.elapsed_time(t1) / args.steps

    def mean_ms(name):
        return statistics.mean(a.elapsed_time(b) for a, b in ev[name]) if ev.get(name) else -1.0
    k_fwd, k_bwd, k_lf, k_lb = mean_ms(K1), mean_ms(K3), mean_ms(K2F), mean_ms(K2B)
    ms, k_fwd, k_bwd, k_lf, k_lb = max_over_ranks(x, [ms, k_fwd, k_bwd, k_lf, k_lb])
    if x.world > 1:
        lt = torch.tensor([launches], device=x.dev, dtype=torch.int64)
        x.dist.all_reduce(lt, op=x.dist.ReduceOp.SUM)
        launches = int(lt.item())
        if isinstance(x.group, x.pkg.dist.PeerExchange):
            x.group.check()
    value = B * x.world / (ms * 1e-3)

    # ---- 3a. the same step replayed as ONE CUDA graph
    if graphed is None:
        try:
            gs = make_graph(x, c, logits, target, cams, x.group)
            if gs is not None:
                gms = time_steps(x, gs.__call__, args.steps, 3)
                line_extra["cuda_graph_replay"] = {"value": round(B * x.world / (gms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(gms, 4),
                                                   "note": "ops.GraphedReprojStep: the step's launches captured once, one cudaGraphLaunch per step"}
            del gs
        except Exception as e:                               # never fail the bench line over the extra measurement
            if x.world > 1:
                raise                                        # ... but a rank that drops out of a collective must not be silent
            line_extra["cuda_graph_replay"] = {"error": str(e)[:200]}

    # ---- 3b. sustained: the same step back to back for >= 2 s, clocks recorded
    if not args.no_sustained:
        n_sus = int(math.ceil(2000.0 / ms))
        s2 = ClockSampler(x.local_rank, period=0.05) if x.rank == 0 else None
        if s2:
            s2.start()
        sms = time_steps(x, step, n_sus, 0)
        line_extra["sustained"] = {"value": round(B * x.world / (sms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(sms, 4), "steps": n_sus,
                                   "seconds": round(sms * n_sus * 1e-3, 2), "vs_value": round((B * x.world / (sms * 1e-3)) / value, 4),
                                   "clocks": s2.finish() if s2 else None}

    # ---- 4. end to end through the public API with pinned host buffers (H2D of every input, D2H of the results)
    e2e = None
    if not args.no_e2e:
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
        w = c["w"]

        def e2e_step():
            with torch.no_grad():
                d_logits.copy_(h_logits, non_blocking=True)
                d_target.copy_(h_target, non_blocking=True)
                for k in d_cams:
                    d_cams[k].copy_(h_cams[k], non_blocking=True)
            d_logits.grad = None
            lp, ls, sel, kps, *_ = ops.integral_reproj_min_loss(d_logits, d_target, d_cams, K, NH, c["NS"], w_mse=w[0], w_bone=w[1],
                                                                w_kp=w[2], w_kp2d=w[3], reduction="batch", group=x.group)
            (lp + ls).backward()
            h_out["loss"].copy_(torch.stack((lp.detach(), ls.detach())), non_blocking=True)
            h_out["sel"].copy_(sel, non_blocking=True)
            h_out["kps"].copy_(kps.detach(), non_blocking=True)
            torch.cuda.current_stream().synchronize()          # the caller reads the loss every step

        def h2d_only():
            with torch.no_grad():
                d_logits.copy_(h_logits, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e2e_ms = time_steps(x, e2e_step, e2e_steps, 1)
        h2d_ms = time_steps(x, h2d_only, 3, 1)                 # all ranks at once: the host-to-device ceiling of this box at this N
        prefetch = None
        try:                                                   # the same steps as a double-buffered input pipeline (extra; `value` above is the serial loop)
            sets = [(d_logits, d_target, d_cams),
                    (torch.empty_like(logits).requires_grad_(True), torch.empty_like(target), {k: torch.empty_like(v) for k, v in cams.items()})]

            def copy_in(i):
                dl, dt, dc = sets[i]
                with torch.no_grad():
                    dl.copy_(h_logits, non_blocking=True)
                    dt.copy_(h_target, non_blocking=True)
                    for k in dc:
                        dc[k].copy_(h_cams[k], non_blocking=True)

            def compute(i):
                dl, dt, dc = sets[i]
                dl.grad = None
                lp, ls, sel, kps, *_ = ops.integral_reproj_min_loss(dl, dt, dc, K, NH, c["NS"], w_mse=w[0], w_bone=w[1], w_kp=w[2],
                                                                    w_kp2d=w[3], reduction="batch", group=x.group)
                (lp + ls).backward()
                h_out["loss"].copy_(torch.stack((lp.detach(), ls.detach())), non_blocking=True)
                h_out["sel"].copy_(sel, non_blocking=True)
                h_out["kps"].copy_(kps.detach(), non_blocking=True)
            p_ms = time_prefetched(x, e2e_steps, copy_in, compute)
            prefetch = {"value": round(B * x.world / (p_ms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(p_ms, 3), "steps": e2e_steps,
                        "note": "two device input sets; the copies of step s+1 run on a second stream under the kernels of step s"}
            del sets
        except RuntimeError as e:
            prefetch = {"error": str(e)[:200]}
        h2d = h_logits.numel() * h_logits.element_size() + h_target.numel() * 4 + sum(v.numel() * 4 for v in h_cams.values())
        d2h = 2 * 4 + 2 * 8 + h_out["kps"].numel() * 4
        e2e = {"value": round(B * x.world / (e2e_ms * 1e-3), 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d) * x.world,
               "d2h_bytes_per_step": int(d2h) * x.world, "bytes_are": "whole job (all %d ranks)" % x.world, "ms_per_step": round(e2e_ms, 3),
               "steps": e2e_steps,
               "h2d_only": {"ms": round(h2d_ms, 3), "gbs_per_gpu": round(h_logits.numel() * h_logits.element_size() / (h2d_ms * 1e-3) / 1e9, 1),
                            "share_of_e2e_step": round(h2d_ms / e2e_ms, 4),
                            "note": "the logits copy alone, all ranks copying at once (max over ranks): the PCIe / host-memory ceiling "
                                    "of this box at this N; the kernels add the rest"},
               "host_binding": x.numa, "prefetch": prefetch,
               "note": "pinned host -> device copy of logits/target/cameras, fused op fwd+bwd, loss/sel/kps read back; PCIe / host-memory "
                       "bound (compare h2d_only): each rank is bound to its GPU's NUMA node before the pinned buffers are allocated"}
        del h_logits, d_logits, h_out, d_target, d_cams
    del logits, target, cams, step, graphed, eager_step
    torch.cuda.empty_cache()

    # ---- 3c. the other BASELINE configs at this N
    if not args.no_configs and args.config == "c2":
        line_extra["configs"] = {name: run_config(x, args, name, max(5, min(args.steps, 20))) for name in ("c3", "c4")}
        line_extra["conv_fused"] = run_conv_fused(x, args, c, max(5, min(args.steps, 20)))

    if x.rank == 0:
        peak, peak_src = measured_peak()
        unit_bytes = K * R ** 3 * esize(c) * B                   # one pass over this rank's volume
        bwd_gbs = 2 * unit_bytes / (k_bwd * 1e-3) / 1e9
        fwd_gbs = unit_bytes / (k_fwd * 1e-3) / 1e9
        step_gbs = 3 * unit_bytes / (ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": x.world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": c["dtype"], "data": "synthetic",
                "config": config_block(c, x.world, scope=args.scope),
                "run": {"exchange": x.exchange, "launch": "cuda graph replay" if args.graph else "eager (one C-ABI call per kernel)"},
                "parity_gate": gate,
                "roofline": {"bound": "hbm", "kernel": "integral_bwd_kernel (read logits + write grad, 2 passes)",
                             "achieved": round(bwd_gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(bwd_gbs / peak, 4),
                             "traffic": ncu_traffic(args.config, "integral_bwd_kernel", B), "peak_source": peak_src,
                             "ms_per_launch": round(k_bwd, 4), "share_of_step": round(k_bwd / ms, 4),
                             "fwd_kernel": {"achieved": round(fwd_gbs, 1), "frac": round(fwd_gbs / peak, 4),
                                            "ms_per_launch": round(k_fwd, 4), "share_of_step": round(k_fwd / ms, 4),
                                            "traffic": ncu_traffic(args.config, "integral_fwd_kernel", B)},
                             "loss_kernels": {"fwd_select_exchange_ms": round(k_lf, 4), "bwd_coef_ms": round(k_lb, 4),
                                              "note": "one launch each (eager: CUDA events include the host gap before the launch)"},
                             "whole_step": {"achieved": round(step_gbs, 1), "frac": round(step_gbs / peak, 4),
                                            "frac_of_8TBs": round(step_gbs / 8000.0, 4)}},
                "e2e": e2e, "gpu_launches": int(launches), "launches_per_step": int(launches_per_step), "clocks": clocks}
        line.update(line_extra)
        line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if x.world > 1:
        x.dist.barrier()
        x.dist.destroy_process_group()
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
                    help="time the step as one CUDA graph replay; the per-kernel events of the roofline block then come from an eager pass before it")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (and with it the GPU-vs-CPU-port half of the parity gate)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (profiling runs)")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[2]/[3] extra")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s sustained extra")
    args = ap.parse_args()
    c = dict(CONFIGS[args.config])
    if args.batch:
        c["B"] = args.batch
    if args.impl == "reference":
        return run_reference(args, c)
    return run_ours(args, c)


if __name__ == "__main__":
    sys.exit(main())
