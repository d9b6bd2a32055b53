"""Time the [4,NH] fp32 all-reduce as the fused op issues it (N ranks, torchrun)."""
import os, torch, torch.distributed as dist
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
t = torch.ones(4, 3, device=dev)
for _ in range(20): dist.all_reduce(t)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): dist.all_reduce(t)
e1.record(); torch.cuda.synchronize()
if rank == 0: print("all_reduce [4,3] fp32: %.1f us per call (back-to-back, device time)" % (e0.elapsed_time(e1) * 1000 / 200))
# with a compute kernel in between, like the real step
x = torch.randn(64 << 20, device=dev)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0.record()
for _ in range(50):
    y = x * 2.0
    dist.all_reduce(t)
    y = x * 3.0
e1.record(); torch.cuda.synchronize()
a = e0.elapsed_time(e1)
e0.record()
for _ in range(50):
    y = x * 2.0
    y = x * 3.0
e1.record(); torch.cuda.synchronize()
b = e0.elapsed_time(e1)
if rank == 0: print("kernel, all_reduce, kernel: %.1f us extra per iteration vs no collective" % ((a - b) * 1000 / 50))
dist.destroy_process_group()
