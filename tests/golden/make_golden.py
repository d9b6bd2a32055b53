"""Generate golden vectors by running the UNMODIFIED reference on seeded inputs.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference's own modules with the two shims of SURVEY.md App. B
(an `easydict` stub; `get_pose_net -> nn.Identity()` so that
`KPDetector3DMulti.forward` runs its lines 67-88 on supplied logits), calls
them on the synthetic inputs of `x-as-supervision_b200/synth.py`, and stores
the outputs as `tests/golden/*.npz`.  Inputs are regenerated from seeds at
test time; an input checksum is stored to catch RNG drift.

Nothing here is read on the GPU box; only the .npz files travel.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("XSUP_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
synth = importlib.import_module("x-as-supervision_b200.synth")


def load_reference():
    ed = types.ModuleType("easydict")

    class EasyDict(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

    ed.EasyDict = EasyDict
    sys.modules["easydict"] = ed
    sys.path.insert(0, REF)
    multi = importlib.import_module("modules.keypoint_detector_integral_multi")
    single = importlib.import_module("modules.keypoint_detector_integral")
    multi.get_pose_net = lambda cfg, num_joints: nn.Identity()
    single.get_pose_net = lambda cfg, num_joints: nn.Identity()
    util = importlib.import_module("modules.util")
    lf = importlib.import_module("modules.base_losses.loss_func")
    return multi, single, util, lf


def checksum(t):
    t = t.double().flatten()
    return np.array([t.sum().item(), (t * torch.arange(1, t.numel() + 1, dtype=torch.float64) % 7.0).sum().item(),
                     t[0].item(), t[-1].item()])


def run_in(dtype, fn):
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)   # the reference builds its aranges with the default dtype (:50-51,57)
    try:
        return fn()
    finally:
        torch.set_default_dtype(old)


def head_case(multi, name, gen, B, K, R, NH, NS, seed, grad_stride):
    logits32 = gen(B, K, R, R, R, seed=seed)
    out = {"meta": np.array([B, K, R, R, R, NH, NS, seed, grad_stride]), "in_checksum": checksum(logits32)}
    gw = torch.randn(B, NH, K, 3, generator=torch.Generator().manual_seed(100 + seed), dtype=torch.float64)
    out["g_kps"] = gw.numpy()
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        def go():
            det = multi.KPDetector3DMulti("resnet_multi", K, R, NH, NS)
            x = logits32.to(dt).clone().requires_grad_(True)
            kps, dmap = det(x)
            # peak bins exactly as the module computes them
            with torch.no_grad():
                p = torch.softmax(x.detach().view(B, K, -1), 2).view(B, K, R, R, R)
                idx = det.find_peak(p.sum(dim=3).sum(dim=3))
            (kps * gw.to(dt)).sum().backward()
            return kps.detach(), dmap.detach(), idx, x.grad.detach()
        kps, dmap, idx, g = run_in(dt, go)
        out["kps_" + tag] = kps.numpy()
        out["dmap_" + tag] = dmap.numpy()
        out["idx_" + tag] = idx.numpy()
        gf = g.flatten()
        out["grad_sub_" + tag] = gf[::grad_stride].numpy().astype(np.float64 if dt == torch.float64 else np.float32)
        out["grad_norms_" + tag] = np.array([gf.double().abs().max().item(), gf.double().norm().item(),
                                             gf.double().sum().item()])
    # number of genuine local maxima per row (fp64), so tests know which slots are defined
    p = torch.softmax(logits32.double().view(B, K, -1), 2).view(B, K, R, R, R)
    pz = p.sum(dim=(3, 4))
    mid = pz[..., 1:-1]
    out["num_peaks"] = ((mid >= pz[..., :-2]) & (mid >= pz[..., 2:])).sum(-1).numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items()})


def single_case(single, name, B, K, R, seed):
    logits32 = synth.iid_logits(B, K, R, R, R, seed=seed)
    out = {"meta": np.array([B, K, R, seed]), "in_checksum": checksum(logits32)}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        def go():
            det = single.KPDetector3D("resnet", K, R)
            kps, dmap = det(logits32.to(dt))
            return kps, dmap
        kps, dmap = run_in(dt, go)
        out["kps_" + tag] = kps.numpy()
        out["dmap_" + tag] = dmap.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name)


def geometry_case(util, name, B, K, seed, mpi):
    cams = synth.cameras(B, seed=seed, mpi=mpi)
    kps = synth.pseudo_joints(B, K, seed=seed + 1)
    out = {"meta": np.array([B, K, seed, int(mpi)]), "in_checksum": checksum(kps)}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        c = {k: v.to(dt) for k, v in cams.items()}
        params = synth.camera_dict(c, "cam_0")
        k = kps.to(dt)
        world = util.convert_patch_to_world(k, params, "cam_0", is_norm=True)
        back = util.convert_world_to_patch(world, params, "cam_0", is_norm=True)
        mono = util.convert_patch_to_world(k, params, "cam_0", is_norm=True, RECT_WIDTH=256, mono=True, patch=False)
        img = util.convert_patch_to_image(k, c["trans_image"].clone(), 256, 256, 256, 2000.0 / 256, c["pelvis"])
        out["world_" + tag] = world.numpy()
        out["back_" + tag] = back.numpy()
        out["mono_" + tag] = mono.numpy()
        out["image_" + tag] = img.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name)


def loss_case(multi, util, lf, name, gen, B, K, R, NH, NS, seed, weights, grad_stride, mpi=False):
    """model.py:71-79 (per-hypothesis world lift), :105-114 (symmetry min), :158-162 (pseudo min),
    driven with the reference's own functions, fwd + bwd to the logits."""
    w_mse, w_bone, w_kp, w_kp2d = weights
    logits32 = gen(B, K, R, R, R, seed=seed)
    target = synth.pseudo_joints(B, K, seed=seed + 2)
    cams = synth.cameras(B, seed=seed + 3, mpi=mpi)
    # every (b,k) row must have >= NH genuine depth peaks, else the reference's loss rides on
    # torch.topk's arbitrary order among tied zeros
    p = torch.softmax(logits32.double().view(B, K, -1), 2).view(B, K, R, R, R)
    pz = p.sum(dim=(3, 4))
    mid = pz[..., 1:-1]
    assert int(((mid >= pz[..., :-2]) & (mid >= pz[..., 2:])).sum(-1).min()) >= NH, name
    out = {"meta": np.array([B, K, R, NH, NS, seed, grad_stride, int(mpi)]),
           "weights": np.array([w if w is not None else np.nan for w in weights], dtype=np.float64),
           "in_checksum": checksum(logits32)}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        def go():
            det = multi.KPDetector3DMulti("resnet_multi", K, R, NH, NS)
            x = logits32.to(dt).clone().requires_grad_(True)
            c = {k: v.to(dt) for k, v in cams.items()}
            params = synth.camera_dict(c, "cam_0")
            kps, _ = det(x)
            world = torch.stack([util.convert_patch_to_world(kps[:, i], params, "cam_0", is_norm=True)
                                 for i in range(NH)], dim=1)
            pseudo = [lf.compute_supervision(kps[:, i], target.to(dt)) for i in range(NH)]
            loss_p = torch.min(torch.stack(pseudo)) * w_mse
            sym = []
            use_sym = any(w is not None for w in (w_bone, w_kp, w_kp2d))
            loss_s = torch.zeros((), dtype=dt)
            if use_sym:
                for i in range(NH):
                    t = 0
                    t = t + lf.compute_bone_sym_loss(world[:, i]) * (w_bone or 0.0)
                    t = t + lf.compute_kp_sym_loss(world[:, i]) * (w_kp or 0.0)
                    if w_kp2d is not None:
                        t = t + lf.compute_kp_sym_loss(kps[:, i, :, :2], is_3D=False) * 1e2 * w_kp2d
                    sym.append(t)
                loss_s = torch.min(torch.stack(sym))
            (loss_p + loss_s).backward()
            return (kps.detach(), world.detach(), torch.stack(pseudo).detach(),
                    torch.stack(sym).detach() if use_sym else torch.zeros(NH, dtype=dt),
                    loss_p.detach(), loss_s.detach(), x.grad.detach())
        kps, world, pseudo, sym, lp, ls, g = run_in(dt, go)
        out["kps_" + tag] = kps.numpy()
        out["world_" + tag] = world.numpy()
        out["pseudo_h_" + tag] = pseudo.numpy()
        out["sym_h_" + tag] = sym.numpy()
        out["loss_" + tag] = np.array([lp.item(), ls.item()])
        gf = g.flatten()
        out["grad_sub_" + tag] = gf[::grad_stride].numpy()
        out["grad_norms_" + tag] = np.array([gf.double().abs().max().item(), gf.double().norm().item(),
                                             gf.double().sum().item()])
    # eval-style per-joint best hypothesis (eval.py:138-145), fp64
    k64 = torch.from_numpy(out["kps_f64"])
    best_idx = (k64 - target.double()[:, None]).pow(2).sum(dim=-1).argmin(dim=1)
    out["best_idx_f64"] = best_idx.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "loss", out["loss_f64"], "pseudo_h", out["pseudo_h_f64"], "sym_h", out["sym_h_f64"])


LOSS_VARIANTS = (("mse", False, False), ("clip", False, True), ("w", True, False), ("wclip", True, True))


def skeleton_case(util, lf, model, name, B, K, S, seed, extension, sub):
    """draw_lines (util.py:21-59) -> channel max (model.py:94) -> compute_mask_reconstruction_loss
    (loss_func.py:4-16) in its four (weight, use_clip) variants, reduced with .mean() as train.py:182 does,
    fwd + bwd to the 2-D keypoints; plus the VJP of the max-map and of the un-maxed heat-maps with seeded
    cotangents."""
    parent, child = model.cal_links(list(synth.H36M_PARENTS), line_select_ids=list(synth.LINE_SELECT), use_root=False,
                                    extension=extension)
    pose = synth.skeleton_pose2d(B, K, seed=seed)
    gt = synth.silhouette_mask(synth.skeleton_pose2d(B, K, seed=seed + 1, jitter=0.03), S)
    wmap = synth.geodesic_weight(gt, seed=seed + 2)
    gen = torch.Generator().manual_seed(300 + seed)
    G = torch.randn(B, 1, S, S, generator=gen, dtype=torch.float64)
    GH = torch.randn(B, len(parent), S, S, generator=gen, dtype=torch.float64)
    out = {"meta": np.array([B, K, S, seed, int(extension), sub, len(parent)]), "parent": np.array(parent), "child": np.array(child),
           "in_checksum": checksum(pose), "gt_checksum": checksum(gt), "w_checksum": checksum(wmap)}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        def go():
            kp = pose.to(dt).clone().requires_grad_(True)
            heat = util.draw_lines(kp, S, parent, child, synth.BODY_WIDTH)
            recon = torch.max(heat.clone(), dim=1, keepdim=True)[0]
            res = {"heat_sub": heat.detach()[:, :, ::sub, ::sub], "recon": recon.detach()}
            res["g_recon"], = torch.autograd.grad((recon * G.to(dt)).sum(), kp, retain_graph=True)
            res["g_heat"], = torch.autograd.grad((heat * GH.to(dt)).sum(), kp, retain_graph=True)
            for vname, use_w, clip in LOSS_VARIANTS:
                loss = lf.compute_mask_reconstruction_loss(recon, gt.to(dt), weight=wmap.to(dt) if use_w else None, use_clip=clip)
                res["loss_shape_" + vname] = torch.tensor(list(loss.shape) or [0])
                res["loss_" + vname] = loss.mean().detach()
                res["g_loss_" + vname], = torch.autograd.grad(loss.mean(), kp, retain_graph=True)
            return res
        for k, v in run_in(dt, go).items():
            out[k + "_" + tag] = v.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: float(out["loss_%s_f64" % k]) for k, _, _ in LOSS_VARIANTS}, "lines", len(parent))


def load_eval_utils():
    """eval_utils.py imports matplotlib and train_util at module level only for plotting helpers: stub them."""
    for m in ("matplotlib", "matplotlib.gridspec", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    tu = types.ModuleType("train_util")
    tu.pose_vis = None
    sys.modules.setdefault("train_util", tu)
    return importlib.import_module("eval_utils")


def eval_case(util, lf, eu, name, B, NH, K, V, seed):
    """eval.py:122-148 (its own lines, driven with eval_utils.switch_points / per_act_mse), util.triangulation on V
    synthetic cameras, and loss_func.compute_disc_loss in its four input-shape combinations."""
    kps32, jp32 = synth.eval_predictions(B, NH, K, seed=seed)
    gen = torch.Generator().manual_seed(400 + seed)
    world = torch.randn(B, K, 3, generator=gen) * 300
    cams = [synth.cameras(B, seed=seed + 10 + i) for i in range(V)]
    noise = [0.002 * torch.randn(B, K, 3, generator=gen) for _ in range(V)]
    logits = {"p2": torch.randn(B, 1, generator=gen), "g2": torch.randn(B, 1, generator=gen),
              "p3": torch.randn(B, NH, 1, generator=gen), "g3": torch.randn(B, NH, 1, generator=gen)}
    out = {"meta": np.array([B, NH, K, V, seed]), "in_checksum": checksum(kps32), "jp_checksum": checksum(jp32),
           "world_checksum": checksum(world)}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        kps, jp = kps32.to(dt), jp32.to(dt)
        kp_gt = jp.clone()
        kp_gt[..., :2] = kp_gt[..., :2] / (256.0 - 1) * 2 - 1
        kp_gt[..., 2] = kp_gt[..., 2] / (256.0 - 1)
        kd, k2 = kps.clone(), kps.clone()[..., :2]
        tr = None
        for h in range(NH):
            k2[:, h, ...], _ = eu.switch_points(k2[:, h, ...], kp_gt[..., :2])
            kd[:, h, ...], tr = eu.switch_points(kd[:, h, ...], kp_gt, switch_all=False)
        best_idx = (kd - kp_gt[:, None, ...]).pow(2).sum(dim=-1).argmin(dim=1)
        kbest = torch.gather(kd, 1, best_idx[:, None, :, None].expand(-1, -1, -1, 3)).squeeze(1)
        best_2d_idx = (k2 - kp_gt[:, None, ..., :2]).pow(2).sum(dim=-1).argmin(dim=1)
        k2best = torch.gather(k2, 1, best_2d_idx[:, None, :, None].expand(-1, -1, -1, 2)).squeeze(1)
        out["kp3d_" + tag], out["kp2d_" + tag] = kbest.numpy(), k2best.numpy()
        out["is_trans_" + tag], out["best_idx_" + tag], out["best_2d_idx_" + tag] = tr.numpy(), best_idx.numpy(), best_2d_idx.numpy()
        out["err2d_" + tag] = eu.per_act_mse(k2best, kp_gt[..., :2]).numpy()
        out["err2d_h0_" + tag] = eu.per_act_mse(k2[:, 0], kp_gt[..., :2]).numpy()          # mode 'confident'
        # triangulation: each camera observes the same world points (projected with the reference's own inverse), + noise
        params, kd3 = {}, {}
        for i, c in enumerate(cams):
            cd = {k: v.to(dt) for k, v in c.items()}
            params.update(synth.camera_dict(cd, "cam_%d" % i))
            kd3["cam_%d" % i] = util.convert_world_to_patch(world.to(dt), params, "cam_%d" % i, is_norm=True) + noise[i].to(dt)
        out["tri_" + tag] = util.triangulation(kd3, params, list(range(V))).numpy()
        if tag == "f32":
            out["tri_inputs_f32"] = torch.stack([kd3["cam_%d" % i] for i in range(V)]).numpy()
        L = {k: v.to(dt) for k, v in logits.items()}
        out["disc_" + tag] = np.array([lf.compute_disc_loss(L["p2"], None).item(), lf.compute_disc_loss(L["p3"], None).item(),
                                       lf.compute_disc_loss(L["p2"], L["g2"]).item(), lf.compute_disc_loss(L["p3"], L["g3"]).item(),
                                       lf.compute_disc_loss(L["p3"], L["g2"]).item()])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "swapped joints", int(out["is_trans_f64"].sum()), "disc", out["disc_f64"])


def stub_discriminator(K, dim, dtype):
    """A seeded stand-in for the GCN discriminator (PyG is not installed here): [N, K, dim] -> [N, 1]."""
    g = torch.Generator().manual_seed(7)
    lin = nn.Linear(K * dim, 1)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(1, K * dim, generator=g, dtype=torch.float32) * 0.5)   # same bits whatever the default dtype
        lin.bias.fill_(0.25)
    net = nn.Sequential(nn.Flatten(), lin).to(dtype)
    net.name = "Linear"                      # Counter3DDisc reads `.name` (model.py:207)
    return net


def model_case(multi, model, name, B, K, R, NH, NS, seed, sym, use_dis_map, grad_stride):
    """The reference's own Counter3DModel.forward and Counter3DDisc.forward (modules/model.py) on a two-camera batch, with
    KPDetector3DMulti(backbone = identity) as regressor and a seeded linear discriminator; loss_values, their sum as the
    trainer forms it (train.py:182-183) and its gradient w.r.t. every logits tensor."""
    cfg = synth.model_cfg(sym=sym, use_dis_map=use_dis_map)
    batch = synth.model_batch(B, K, R, seed=seed)
    out = {"meta": np.array([B, K, R, NH, NS, seed, grad_stride, int(use_dis_map)]),
           "sym": np.array(sym if sym is not None else [np.nan] * 3, dtype=np.float64),
           "in_checksum": checksum(batch["cam_0_img"]), "mask_checksum": checksum(batch["cam_1_mask"])}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        def go():
            x = {k: (v.to(dt).clone() if v.is_floating_point() else v.clone()) for k, v in batch.items()}
            leaves = {}
            for k in list(x):
                if k.endswith("_img"):
                    x[k].requires_grad_(True)
                    leaves[k] = x[k]
            det = multi.KPDetector3DMulti("resnet_multi", K, R, NH, NS)
            disc = stub_discriminator(K, 3, dt)
            m = model.Counter3DModel(cfg, det, None, None, physique_network=None)
            loss_values, output = m(x, disc)
            total = sum(v.mean() for v in loss_values.values())
            total.backward()
            d = model.Counter3DDisc(cfg, disc, None, None)
            loss_disc, _ = d({k: v.detach() for k, v in x.items()}, det)
            res = {"loss_" + k: v.mean().detach() for k, v in loss_values.items()}
            res["loss_total"] = total.detach()
            res["loss_disc"] = loss_disc.detach()
            res["recon_cam_0"] = output["mask_heatmap_line_cam_0"].detach()
            res["kp_gt_world"] = output["kp_gt_world"].detach()
            res["pose_3d_gt_cam_1_pseudo"] = output["pose_3d_gt_cam_1_pseudo"].detach()
            for k, v in leaves.items():
                gf = v.grad.flatten()
                res["grad_sub_" + k] = gf[::grad_stride].clone()
                res["grad_norms_" + k] = torch.tensor([gf.double().abs().max().item(), gf.double().norm().item()])
            return res
        for k, v in run_in(dt, go).items():
            out[k + "_" + tag] = v.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: float(out[k]) for k in out if k.startswith("loss_") and k.endswith("_f64")})


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    multi, single, util, lf = load_reference()
    model = importlib.import_module("modules.model")
    if "--only-model" in sys.argv or "--only-skeleton" not in sys.argv and "--only-eval" not in sys.argv:
        # SynthS2-like (symmetry incl. kp_2d, geodesic-weighted reconstruction) and SurS1-like (no symmetry, plain clip) loss graphs
        model_case(multi, model, "model_synths2_k18_r32", B=3, K=18, R=32, NH=3, NS=15, seed=90, sym=(0.1, 0.1, 0.5), use_dis_map=True, grad_stride=251)
        model_case(multi, model, "model_surs1_k18_r32", B=2, K=18, R=32, NH=3, NS=15, seed=95, sym=None, use_dis_map=False, grad_stride=251)
        if "--only-model" in sys.argv:
            return
    if "--only-eval" in sys.argv:
        eval_case(util, lf, load_eval_utils(), "eval_k18_nh3_v4", B=6, NH=3, K=18, V=4, seed=80)
        return
    if "--only-skeleton" not in sys.argv:
        main_head_and_loss(multi, single, util, lf)
        eval_case(util, lf, load_eval_utils(), "eval_k18_nh3_v4", B=6, NH=3, K=18, V=4, seed=80)
    # skeleton rasteriser + mask loss: 25 lines (17 tree links + 8 braces, arm lines at half width) and the
    # 17-line variant without the braces (below the 21-line threshold of util.py:50)
    skeleton_case(util, lf, model, "skel_h36m_s128", B=2, K=18, S=128, seed=30, extension=True, sub=8)
    skeleton_case(util, lf, model, "skel_l17_s64", B=3, K=18, S=64, seed=33, extension=False, sub=4)


def main_head_and_loss(multi, single, util, lf):
    head_case(multi, "head_iid_k18_r16", synth.iid_logits, B=2, K=18, R=16, NH=3, NS=5, seed=0, grad_stride=7)
    head_case(multi, "head_blob_k17_r32", synth.blob_logits, B=2, K=17, R=32, NH=3, NS=15, seed=1, grad_stride=61)
    head_case(multi, "head_blob_k18_r64", synth.blob_logits, B=1, K=18, R=64, NH=3, NS=15, seed=4, grad_stride=997)
    head_case(multi, "head_iid_k3_r8_nh2", synth.iid_logits, B=3, K=3, R=8, NH=2, NS=3, seed=5, grad_stride=1)
    single_case(single, "single_iid_k18_r16", B=2, K=18, R=16, seed=6)
    geometry_case(util, "geom_h36m", B=8, K=18, seed=7, mpi=False)
    geometry_case(util, "geom_mpi", B=8, K=18, seed=8, mpi=True)
    # SurS1: pseudo-MSE x3.0 only (HM36_Multi_SurS1.yaml:72-73)
    loss_case(multi, util, lf, "loss_surs1_k17_r16", synth.iid_logits, B=4, K=17, R=16, NH=3, NS=5, seed=10,
              weights=(3.0, None, None, None), grad_stride=13)
    # SynthS2: pseudo 1.0, bone 0.1, kp 0.1, kp_2d 0.0 (HM36_Multi_SynthS2.yaml:66-86)
    loss_case(multi, util, lf, "loss_synths2_k18_r32", synth.blob_logits, B=2, K=18, R=32, NH=3, NS=15, seed=12,
              weights=(1.0, 0.1, 0.1, 0.0), grad_stride=53)
    loss_case(multi, util, lf, "loss_synths2_k17_r32_mpi", synth.iid_logits, B=2, K=17, R=32, NH=4, NS=15, seed=12,
              weights=(1.0, 0.1, 0.1, 0.5), grad_stride=127, mpi=True)


if __name__ == "__main__":
    main()
