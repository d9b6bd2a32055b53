"""The on-box incumbent: the reference's op sequence (the oracle restatement, plain torch ops) run eagerly in
fp32 ON THE SAME B200, timed beside our kernels.  SURVEY.md 8d asks for this number next to ours; it is a
reported baseline, not a parity gate (parity is test_gpu_parity.py / test_gpu_skeleton.py).  Results go to
gpurun_out/incumbent.json when that directory exists."""
import importlib
import json
import os

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _time(fn, steps=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / steps


def test_incumbent_eager_timings(oracle, synth):
    import __graft_entry__ as ge
    ge.build()
    pkg = importlib.import_module("x-as-supervision_b200")
    ops = pkg.load_native()
    sk = pkg.skeleton
    dev = torch.device("cuda:0")
    out = {}

    # ---- integral head + reprojection min-loss, BASELINE configs[1] shapes at a batch the eager graph fits easily
    B, K, R, NH, NS = 64, 17, 64, 3, 15
    logits = torch.randn(B, K * R, R, R, device=dev)
    target = synth.pseudo_joints(B, K, seed=2).to(dev)
    cams = {k: v.to(dev) for k, v in synth.cameras(B, seed=3).items()}

    def eager():
        x = logits.clone().requires_grad_(True)
        lp, ls, *_ = oracle.fused_forward(x, K, NH, NS, target, cams, w_mse=3.0, reduction="batch")
        (lp + ls).backward()

    def ours():
        x = logits.clone().requires_grad_(True)
        lp, ls, *_ = ops.integral_reproj_min_loss(x, target, cams, K, NH, NS, w_mse=3.0, reduction="batch")
        (lp + ls).backward()

    te, to = _time(eager), _time(ours, steps=20, warmup=3)
    out["integral_reproj"] = {"batch": B, "eager_ms": round(te, 3), "ours_ms": round(to, 3),
                              "eager_samples_per_s": round(B / te * 1e3, 1), "ours_samples_per_s": round(B / to * 1e3, 1),
                              "note": "both include one clone of the logits per step"}
    assert to < te

    # ---- skeleton rasteriser + max + weighted clipped mask loss
    B2, S = 16, 256
    parent, child = sk.cal_links(synth.H36M_PARENTS, synth.LINE_SELECT)
    pose = synth.skeleton_pose2d(B2, 18, seed=60).to(dev)
    gt = synth.silhouette_mask(synth.skeleton_pose2d(B2, 18, seed=61), S).to(dev)
    wmap = synth.geodesic_weight(gt.cpu(), seed=62).to(dev)

    def eager2():
        kp = pose.clone().requires_grad_(True)
        recon = oracle.skeleton_mask(kp, S, parent, child, synth.BODY_WIDTH)
        oracle.mask_recon_loss(recon, gt, weight=wmap, use_clip=True).backward()

    def ours2():
        kp = pose.clone().requires_grad_(True)
        _, loss = sk.skeleton_mask_loss(kp, gt, wmap, S, parent, child, synth.BODY_WIDTH, use_clip=True)
        loss.backward()

    te2, to2 = _time(eager2), _time(ours2, steps=20, warmup=3)
    out["skeleton_mask_loss"] = {"batch": B2, "eager_ms": round(te2, 3), "ours_ms": round(to2, 3),
                                 "eager_samples_per_s": round(B2 / te2 * 1e3, 1), "ours_samples_per_s": round(B2 / to2 * 1e3, 1)}
    assert to2 < te2
    # ---- the whole per-iteration loss graph (Counter3DModel.forward + backward), 4 cameras, batch 32 (HM36_Multi_*.yaml), 64^3
    from torch import nn
    model = importlib.import_module("x-as-supervision_b200.model")
    Bm, Km, Rm, cams4 = 32, 18, 64, (0, 1, 2, 3)
    cfg = synth.model_cfg(cam_ids=cams4, sym=(0.1, 0.1, 0.0), use_dis_map=False)
    xb = {k: v.to(dev) for k, v in synth.model_batch(Bm, Km, Rm, cam_ids=cams4, seed=120).items()}
    lin = nn.Linear(Km * 3, 1).to(dev)
    disc = nn.Sequential(nn.Flatten(), lin)
    det = pkg.detector.KPDetector3DMulti("resnet_multi", Km, Rm, 3, 15, net=nn.Identity())
    ours_model = model.Counter3DModel(cfg, det, None, None)
    parent, child = oracle.skeleton_links(synth.H36M_PARENTS, synth.LINE_SELECT)
    lc = cfg["loss_config"]

    def ours3():
        xs = {k: (v.clone().requires_grad_(True) if k.endswith("_img") else v) for k, v in xb.items()}
        lv, _ = ours_model(xs, disc)
        sum(v.mean() for v in lv.values()).backward()

    def eager3():                                            # the reference's op sequence (oracle pieces), per camera and hypothesis
        xs = {k: (v.clone().requires_grad_(True) if k.endswith("_img") else v) for k, v in xb.items()}
        total = 0
        for c in cams4:
            key = "cam_%d" % c
            cm = {k: xs[key + "_" + k] for k in ("trans_image", "pelvis", "k_mat", "trans_world", "rot_world")}
            kps, _, _ = oracle.integral_multi(xs[key + "_img"], Km, 3, 15)
            _, ls, _, world = oracle.reproj_min_loss(kps, torch.zeros(Bm, Km, 3, device=dev), cm, img_hw=(Rm, Rm), w_mse=0.0, w_bone=0.1,
                                                     w_kp=0.1, w_kp2d=0.0, reduction="batch")
            cen = oracle.root_centre(world, 3).detach()
            lg = oracle.disc_loss(torch.stack([disc(cen[:, i]) for i in range(3)], 1), None) * lc["smpl_gen_loss"]["weight"]
            pk, _, _ = oracle.integral_multi(xs[key + "_pseudo_img"], Km, 3, 15)
            lp, _, _, _ = oracle.reproj_min_loss(pk, xs[key + "_pseudo_joints"], cm, img_hw=(Rm, Rm), w_mse=1.0, reduction="batch")
            recon = oracle.skeleton_mask(kps[:, 0, :, :2], Rm, parent, child, synth.BODY_WIDTH)
            lr = oracle.mask_recon_loss(recon, xs[key + "_mask"], weight=None, use_clip=True).mean() * lc["recons_loss"]["weight"]
            total = total + ls + lg + lp + lr
        total.backward()

    te3, to3 = _time(eager3, steps=3, warmup=1), _time(ours3, steps=10, warmup=2)
    graphed = model.GraphedLossGraph(ours_model, xb, disc)
    tg3 = _time(graphed, steps=20, warmup=3)
    out["loss_graph_4cam_b32"] = {"eager_ms": round(te3, 3), "ours_ms": round(to3, 3), "speedup": round(te3 / to3, 2),
                                  "ours_cuda_graph_ms": round(tg3, 3), "speedup_cuda_graph": round(te3 / tg3, 2),
                                  "note": "Counter3DModel.forward + backward to 8 logits tensors [32, 18*64, 64, 64] (4 cameras x real/pseudo image)"}
    assert to3 < te3
    print("\n[incumbent]", json.dumps(out))
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "incumbent.json"), "w") as f:
            json.dump(out, f, indent=1)
