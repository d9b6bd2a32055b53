"""2-rank NCCL test (needs >= 2 GPUs, `-m gpu`): sample-sharded fused op in 'global' scope must select the
slot, report the loss and produce the heat-map gradient of the single-GPU run on the whole batch."""
import importlib
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu

B, K, R, NS = 8, 18, 32, 15
W = dict(w_mse=1.0, w_bone=0.1, w_kp=0.1, w_kp2d=0.0)


def _inputs(synth):
    return (synth.blob_logits(B, K, R, R, R, seed=101), synth.pseudo_joints(B, K, seed=102), synth.cameras(B, seed=103))


def _worker(rank, world, port, out, transport, NH):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        pkg = importlib.import_module("x-as-supervision_b200")
        ops = pkg.load_native()
        logits, target, cams = _inputs(pkg.synth)
        lo, hi = pkg.dist.shard_range(B, rank, world)
        group = pkg.dist.PeerExchange(dist.group.WORLD, dev) if transport == "nvlink" else dist.group.WORLD
        for _ in range(3):                                   # several steps: mailbox parity / sequence numbers
            x = logits[lo:hi].to(dev).requires_grad_(True)
            lp, ls, sel, kps, world_pts, _, _ = ops.integral_reproj_min_loss(
                x, target[lo:hi].to(dev), {k: v[lo:hi].to(dev) for k, v in cams.items()}, K, NH, NS, reduction="batch",
                group=group, **W)
            (lp + ls).backward()
        res = {"loss": (lp.item(), ls.item()), "sel": sel.cpu(), "grad": x.grad.cpu(), "range": (lo, hi)}
        if transport == "nvlink":
            # the same step captured as ONE CUDA graph (the exchange kernel keeps its sequence number on the device)
            # and replayed three times on every rank: every replay must reproduce the eager result bit for bit
            step = ops.GraphedReprojStep(logits[lo:hi].to(dev), target[lo:hi].to(dev), {k: v[lo:hi].to(dev) for k, v in cams.items()},
                                         K, NH, NS, reduction="batch", group=group, **W)
            ok = True
            for _ in range(3):
                glp, gls, gsel = step()
                torch.cuda.synchronize()
                ok = ok and torch.equal(glp, lp.detach()) and torch.equal(gls, ls.detach()) and torch.equal(gsel, sel) \
                    and torch.equal(step.grad, x.grad)
            res["graph_ok"] = bool(ok)
        torch.save(res, os.path.join(out, "rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("transport,NH", [("nvlink", 3), ("nccl", 3), ("nvlink", 16), ("nvlink", 30)])
def test_two_gpus_global_scope_equals_single_gpu(synth, tmp_path, transport, NH):
    """NH = 16 and D-2 = 30: the exchanged [4, NH] partial sums no longer fit one 64-float line (mailbox slots hold 1023)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import __graft_entry__ as ge
    ge.build()
    world, port = 2, 29500 + (os.getpid() * 4 + (transport == "nccl") + NH) % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path), transport, NH), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r), weights_only=False) for r in range(world)]

    pkg = importlib.import_module("x-as-supervision_b200")
    ops = pkg.load_native()
    dev = torch.device("cuda:0")
    logits, target, cams = _inputs(synth)
    x = logits.to(dev).requires_grad_(True)
    lp, ls, sel, *_ = ops.integral_reproj_min_loss(x, target.to(dev), {k: v.to(dev) for k, v in cams.items()}, K, NH, NS,
                                                   reduction="batch", **W)
    (lp + ls).backward()
    for r in res:
        assert r.get("graph_ok", True), "CUDA-graph replay of the global-scope step differs from the eager step"
        assert torch.equal(r["sel"], sel.cpu())                                   # same slots on every rank
        assert abs(r["loss"][0] - lp.item()) < 1e-6 * abs(lp.item()) and abs(r["loss"][1] - ls.item()) < 1e-6 * abs(ls.item())
        lo, hi = r["range"]
        ref = x.grad[lo:hi].cpu()
        assert (r["grad"] - ref).abs().max().item() <= 1e-6 * ref.abs().max().item()


def _conv_inputs(synth):
    g = torch.Generator().manual_seed(211)
    Bc, Kc, Dc, C = 6, 17, 64, 128
    x = torch.randn(Bc, C, Dc, Dc, generator=g)
    w = torch.randn(Kc * Dc, C, generator=g) / C ** 0.5
    w[::7] *= 3.0
    bias = torch.randn(Kc * Dc, generator=g)
    return x, w, bias, synth.pseudo_joints(Bc, Kc, seed=212), synth.cameras(Bc, seed=213), (Bc, Kc, Dc, C)


def _conv_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        pkg = importlib.import_module("x-as-supervision_b200")
        ops = pkg.load_native()
        x, w, bias, target, cams, (Bc, Kc, Dc, C) = _conv_inputs(pkg.synth)
        lo, hi = pkg.dist.shard_range(Bc, rank, world)
        group = pkg.dist.PeerExchange(dist.group.WORLD, dev)
        for _ in range(2):
            xd = x[lo:hi].to(dev).requires_grad_(True)
            wd = w.to(dev).requires_grad_(True)
            bd = bias.to(dev).requires_grad_(True)
            lp, ls, sel, kps, *_ = ops.conv_integral_reproj_min_loss(xd, wd, bd, target[lo:hi].to(dev), {k: v[lo:hi].to(dev) for k, v in cams.items()},
                                                                     Kc, 3, NS, reduction="batch", group=group, **W)
            (lp + ls).backward()
        torch.save({"loss": (lp.item(), ls.item()), "sel": sel.cpu(), "dx": xd.grad.cpu(), "dw": wd.grad.cpu(), "db": bd.grad.cpu(),
                    "range": (lo, hi)}, os.path.join(out, "conv_rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gpus_conv_fused_loss_global_scope(synth, tmp_path):
    """The conv-fused per-camera op (K7 -> fused loss with the in-kernel NVLink exchange -> K8) sharded over 2 GPUs: slots and loss
    of the single-GPU run on the whole batch, each rank's d x shard bit for bit (a sample's gradient depends on the other
    rank only through the selected slot and the global batch size), d W / d bias = sum of the ranks' partial gradients."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import __graft_entry__ as ge
    ge.build()
    world, port = 2, 29500 + (os.getpid() * 4 + 77) % 2000
    mp.spawn(_conv_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "conv_rank%d.pt" % r), weights_only=False) for r in range(world)]
    pkg = importlib.import_module("x-as-supervision_b200")
    ops = pkg.load_native()
    dev = torch.device("cuda:0")
    x, w, bias, target, cams, (Bc, Kc, Dc, C) = _conv_inputs(synth)
    xd = x.to(dev).requires_grad_(True)
    wd = w.to(dev).requires_grad_(True)
    bd = bias.to(dev).requires_grad_(True)
    lp, ls, sel, *_ = ops.conv_integral_reproj_min_loss(xd, wd, bd, target.to(dev), {k: v.to(dev) for k, v in cams.items()}, Kc, 3, NS,
                                                        reduction="batch", **W)
    (lp + ls).backward()
    dw_sum = sum(r["dw"] for r in res)
    db_sum = sum(r["db"] for r in res)
    for r in res:
        assert torch.equal(r["sel"], sel.cpu())
        assert abs(r["loss"][0] - lp.item()) < 1e-6 * abs(lp.item()) and abs(r["loss"][1] - ls.item()) < 1e-6 * abs(ls.item())
        lo, hi = r["range"]
        assert torch.equal(r["dx"], xd.grad[lo:hi].cpu()), "d x of a shard differs from the single-GPU run"
    assert float((dw_sum - wd.grad.cpu()).abs().max()) <= 1e-5 * float(wd.grad.abs().max())
    assert float((db_sum - bd.grad.cpu()).abs().max()) <= 1e-5 * float(bd.grad.abs().max())
