"""Fault isolation helper: run one kernel family at one size in this process (a fault kills the context)."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("x-as-supervision_b200")
ops = pkg.load_native()
cabi = importlib.import_module("x-as-supervision_b200._cabi")
what, B, K, R = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
dt = torch.bfloat16 if len(sys.argv) > 5 and sys.argv[5] == "bf16" else torch.float32
dev = torch.device("cuda:0")
x = torch.randn(B, K * R, R, R, device=dev).to(dt)
torch.cuda.synchronize()
logits, shape, kps, dmap, idx, stats = ops._head_forward(x, K, 3, 15, cabi.HEAD_MULTI)
torch.cuda.synchronize()
print(what, "fwd ok", B, K, R, "kps finite:", bool(torch.isfinite(kps).all()), flush=True)
if what == "bwd":
    g = torch.randn_like(kps)
    out = ops._head_backward(logits, stats, shape, g)
    torch.cuda.synchronize()
    print("bwd ok; grad finite:", bool(torch.isfinite(out.float()).all()), "unit sums max:", float(out.float().view(B * K, -1).sum(-1).abs().max()), flush=True)
