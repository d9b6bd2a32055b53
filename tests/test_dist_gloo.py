"""world_size-2 `gloo` tests (CPU) of the N>1 host logic: sample sharding, the one [4,NH] all-reduce of
the partial loss sums ('global' scope) and the slot selection that follows it.  Each rank gets its
shard's per-sample terms from the oracle (the CUDA kernel produces the same quantity on a GPU); the
reduced result must select what a single process selects on the whole batch, and the rank-local
selection ('local' scope, what the reference does under DDP, model.py:114,162) must match a
single-process run on the shard."""
import importlib
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

B, K, R, NH, NS = 6, 18, 16, 3, 5
W = dict(w_mse=1.0, w_bone=0.1, w_kp=0.1, w_kp2d=0.2)


def _inputs():
    synth = importlib.import_module("x-as-supervision_b200.synth")
    return synth, synth.iid_logits(B, K, R, R, R, seed=91).double(), synth.pseudo_joints(B, K, seed=92).double(), \
        {k: v.double() for k, v in synth.cameras(B, seed=93).items()}


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        xdist = importlib.import_module("x-as-supervision_b200.dist")
        oracle = importlib.import_module("xsup_oracle")
        synth, logits, target, cams = _inputs()
        lo, hi = xdist.shard_range(B, rank, world)
        kps, _, _ = oracle.integral_multi(logits[lo:hi], K, NH, NS)
        c = {k: v[lo:hi] for k, v in cams.items()}
        world_pts = torch.stack([oracle.patch_to_world(kps[:, h], c) for h in range(NH)], dim=1)
        terms = torch.stack(oracle.per_sample_terms(kps, target[lo:hi], world_pts, 0, 0, 0), dim=1)   # [b,4,NH]
        partial = terms.sum(0).float()
        local = xdist.select_slots(partial, hi - lo, K, W["w_mse"], W["w_bone"], W["w_kp"], W["w_kp2d"])
        n_total = xdist.global_batch(hi - lo, dist.group.WORLD)
        n_exact = xdist.global_batch_exact(hi - lo, dist.group.WORLD)
        xdist.reduce_partials(partial, dist.group.WORLD)
        glob = xdist.select_slots(partial, n_total, K, W["w_mse"], W["w_bone"], W["w_kp"], W["w_kp2d"])
        torch.save({"local": local, "global": glob, "n_total": n_total, "n_exact": n_exact, "range": (lo, hi),
                    "partial": partial}, os.path.join(out, "rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_batch():
    xdist = importlib.import_module("x-as-supervision_b200.dist")
    for n in (0, 1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [xdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_no_group_is_a_noop():
    xdist = importlib.import_module("x-as-supervision_b200.dist")
    p = torch.arange(12.0).view(4, 3)
    assert xdist.reduce_partials(p.clone(), None).equal(p)
    assert xdist.global_batch(5, None) == 5 and xdist.global_batch_exact(5, None) == 5


@pytest.mark.timeout(120)
def test_two_rank_global_scope_matches_single_process(oracle, tmp_path):
    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r), weights_only=False) for r in range(world)]
    synth, logits, target, cams = _inputs()
    # single-process oracle on the whole batch
    lp, ls, sel, *_ = oracle.fused_forward(logits, K, NH, NS, target, cams, reduction="batch", **W)
    for r in res:
        sm, ss, vlp, vls = r["global"]
        assert (sm, ss) == (int(sel[0]), int(sel[1]))                 # bit-exact slots on every rank
        assert abs(vlp - float(lp)) < 1e-6 * abs(float(lp)) and abs(vls - float(ls)) < 1e-6 * abs(float(ls))
        assert r["n_total"] == B and r["n_exact"] == B
    assert torch.equal(res[0]["partial"], res[1]["partial"])           # all-reduce leaves identical sums
    assert res[0]["range"] == (0, 3) and res[1]["range"] == (3, 6)
    # 'local' scope == single-process run on the shard (the reference's DDP behaviour)
    for r in res:
        lo, hi = r["range"]
        lp_l, ls_l, sel_l, *_ = oracle.fused_forward(logits[lo:hi], K, NH, NS, target[lo:hi],
                                                     {k: v[lo:hi] for k, v in cams.items()}, reduction="batch", **W)
        sm, ss, vlp, vls = r["local"]
        assert (sm, ss) == (int(sel_l[0]), int(sel_l[1]))
        assert abs(vlp - float(lp_l)) < 1e-6 * abs(float(lp_l))
