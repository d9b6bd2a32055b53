"""Randomised pin of the CPU oracle against the reference itself, imported live from /root/reference.

The committed goldens (tests/golden/*.npz) fix a handful of seeds and shapes; this file draws many more and compares,
in fp64, every stage of the path with what the UNMODIFIED reference modules compute on the same tensors.  It only runs
where the reference checkout exists (the build container); on the GPU box it skips — nothing there reads the reference.
CPU only, a few seconds."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_inf

REF = os.environ.get("XSUP_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "modules")), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    spec = importlib.util.spec_from_file_location("xsup_make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    multi, single, util, lf = mg.load_reference()
    return dict(mg=mg, multi=multi, single=single, util=util, lf=lf)


def _shapes(n, seed):
    rng = np.random.RandomState(seed)
    out = []
    for i in range(n):
        R = int(rng.choice([8, 12, 16, 24]))
        NH = int(rng.choice([1, 2, 3, 5]))
        NS = int(rng.choice([1, 3, 7, 15]))
        out.append((int(rng.randint(1, 4)), int(rng.choice([3, 17, 18])), R, min(NH, R - 2), NS, 1000 + 17 * i + seed))
    return out


def _num_peaks(logits, B, K, R):
    p = torch.softmax(logits.double().view(B, K, -1), 2).view(B, K, R, R, R)
    pz = p.sum(dim=(3, 4))
    mid = pz[..., 1:-1]
    return ((mid >= pz[..., :-2]) & (mid >= pz[..., 2:])).sum(-1)


@pytest.mark.parametrize("B,K,R,NH,NS,seed", _shapes(10, 1))
@pytest.mark.parametrize("gen", ["iid_logits", "blob_logits"])
def test_head_forward_and_backward(ref, oracle, synth, gen, B, K, R, NH, NS, seed):
    """…_multi.py:66-88 and its autograd, against oracle.integral_multi and the closed-form backward (App. A.2)."""
    logits = getattr(synth, gen)(B, K, R, R, R, seed=seed).double()
    gw = torch.randn(B, NH, K, 3, generator=torch.Generator().manual_seed(seed), dtype=torch.float64)

    def go():
        det = ref["multi"].KPDetector3DMulti("resnet_multi", K, R, NH, NS)
        x = logits.clone().requires_grad_(True)
        kps, dmap = det(x)
        with torch.no_grad():
            p = torch.softmax(x.detach().view(B, K, -1), 2).view(B, K, R, R, R)
            idx = det.find_peak(p.sum(dim=3).sum(dim=3))
        (kps * gw).sum().backward()
        return kps.detach(), dmap.detach(), idx, x.grad
    r_kps, r_dmap, r_idx, r_grad = ref["mg"].run_in(torch.float64, go)

    kps, dmap, idx = oracle.integral_multi(logits, K, NH, NS)
    defined = (torch.arange(NH)[None, None, :] < _num_peaks(logits, B, K, R)[..., None])      # [B,K,NH]
    assert torch.equal(idx[defined], r_idx[defined])                                           # bit-exact where defined
    assert rel_inf(dmap.numpy(), r_dmap.numpy()) < 1e-12
    m = defined.permute(0, 2, 1)[..., None].expand_as(kps)
    assert (kps - r_kps)[m].abs().max().item() < 1e-12
    assert (kps[..., :2] - r_kps[..., :2]).abs().max().item() < 1e-12
    if bool(defined.all()):
        closed = oracle.integral_multi_backward(logits, gw, K, NH, NS)
        assert rel_inf(closed.numpy(), r_grad.numpy()) < 1e-10


@pytest.mark.parametrize("B,K,seed,mpi", [(1, 17, 5, False), (4, 18, 6, True), (7, 3, 7, False), (2, 25, 8, True)])
def test_geometry_both_directions(ref, oracle, synth, B, K, seed, mpi):
    """util.py:61-168: patch -> world, its inverse, the mono branch."""
    util = ref["util"]
    cams = {k: v.double() for k, v in synth.cameras(B, seed=seed, mpi=mpi).items()}
    params = synth.camera_dict(cams, "cam_0")
    kps = synth.pseudo_joints(B, K, seed=seed + 1).double()
    r_world = util.convert_patch_to_world(kps, params, "cam_0", is_norm=True)
    r_back = util.convert_world_to_patch(r_world, params, "cam_0", is_norm=True)
    world = oracle.patch_to_world(kps, cams)
    assert rel_inf(world.numpy(), r_world.numpy()) < 1e-12
    assert rel_inf(oracle.world_to_patch(r_world, cams).numpy(), r_back.numpy()) < 1e-10
    r_raw = util.convert_patch_to_world(kps, params, "cam_0", is_norm=False)
    assert rel_inf(oracle.patch_to_world(kps, cams, is_norm=False).numpy(), r_raw.numpy()) < 1e-12


@pytest.mark.parametrize("mode", ["mean", "sum", "none"])
@pytest.mark.parametrize("fs,C", [(None, 3), ((64, 48, 32), 3), ((64, 48), 2)])
def test_compute_supervision_all_modes(ref, oracle, synth, mode, fs, C):
    """loss_func.py:38-52 with feature_shape and every nn.MSELoss reduction, values and gradients."""
    kp, gt = synth.pseudo_joints(4, 18, seed=41).double()[..., :C], synth.pseudo_joints(4, 18, seed=42).double()[..., :C]
    a, b = kp.clone().requires_grad_(True), kp.clone().requires_grad_(True)
    ra, ob = ref["lf"].compute_supervision(a, gt, feature_shape=fs, mode=mode), oracle.supervision(b, gt, feature_shape=fs, mode=mode)
    assert ra.shape == ob.shape
    gw = torch.randn(ra.shape, generator=torch.Generator().manual_seed(43), dtype=torch.float64)
    (ra * gw).sum().backward()
    (ob * gw).sum().backward()
    assert rel_inf(ob.detach().numpy(), ra.detach().numpy()) < 1e-12
    assert rel_inf(b.grad.numpy(), a.grad.numpy()) < 1e-12


@pytest.mark.parametrize("B,K,seed,mpi,is_norm", [(3, 18, 15, False, True), (2, 17, 16, True, False), (5, 4, 17, False, False)])
def test_geometry_stage_by_stage_with_gradients(ref, oracle, synth, B, K, seed, mpi, is_norm):
    """util.py:61-125 one stage at a time, with the free parameters the composites fix (image_depth != width, any
    depth_scale, general matrices), values and autograd VJPs: the pin of the oracle functions that the stage-wise GPU API
    (`ops.convert_patch_to_image` ... `ops.convert_world_to_image`) is tested against."""
    util = ref["util"]
    g = torch.Generator().manual_seed(seed)
    cams = {k: v.double() for k, v in synth.cameras(B, seed=seed, mpi=mpi).items()}
    cams["rot_world"] = cams["rot_world"] + 0.1 * torch.randn(B, 3, 3, generator=g, dtype=torch.float64)
    cams["trans_image"][:, :, :2] += 0.05 * torch.randn(B, 2, 2, generator=g, dtype=torch.float64)
    fx, fy, cx, cy = oracle._intrinsics(cams["k_mat"])
    img_d, img_h, img_w, ds = 200, 240, 256, 6.5
    kps = synth.pseudo_joints(B, K, seed=seed + 1).double()
    if not is_norm:
        kps = (kps + 1) * 100
    gw = torch.randn(B, K, 3, generator=g, dtype=torch.float64)

    def pair(rfn, ofn, x0, *args):
        a, b = x0.clone().requires_grad_(True), x0.clone().requires_grad_(True)
        ra, ob = rfn(a, *args), ofn(b, *args)
        (ra * gw).sum().backward()
        (ob * gw).sum().backward()
        assert rel_inf(ob.detach().numpy(), ra.detach().numpy()) < 1e-12, rfn.__name__
        assert rel_inf(b.grad.numpy(), a.grad.numpy()) < 1e-12, rfn.__name__
        return ra.detach()
    img = pair(util.convert_patch_to_image, oracle.patch_to_image, kps, cams["trans_image"], img_d, img_h, img_w, ds, cams["pelvis"], is_norm)
    world = pair(util.convert_image_to_world, oracle.image_to_world, img, fx, fy, cx, cy, cams["trans_world"], cams["rot_world"])
    img2 = pair(util.convert_world_to_image, oracle.world_to_image, world, fx, fy, cx, cy, cams["trans_world"], cams["rot_world"])
    back = pair(util.convert_image_to_patch, oracle.image_to_patch, img2, cams["trans_image"], img_d, img_h, img_w, ds, cams["pelvis"], is_norm)
    assert rel_inf(back.numpy(), kps.numpy()) < 1e-9            # the four stages compose to the identity


@pytest.mark.parametrize("B,K,R,NH,NS,seed", [(2, 17, 32, 3, 15, 31), (3, 18, 32, 2, 7, 32), (1, 17, 24, 3, 3, 33), (4, 18, 12, 1, 15, 34)])
@pytest.mark.parametrize("weights", [(3.0, None, None, None), (1.0, 0.1, 0.1, 0.0), (1.0, 0.1, 0.1, 0.5)])
def test_fused_loss_and_gradient(ref, oracle, synth, B, K, R, NH, NS, seed, weights):
    """model.py:71-79,105-114,158-162 driven with the reference's own functions, against oracle.fused_forward + autograd."""
    multi, util, lf = ref["multi"], ref["util"], ref["lf"]
    w_mse, w_bone, w_kp, w_kp2d = weights
    logits = synth.iid_logits(B, K, R, R, R, seed=seed).double()
    if int(_num_peaks(logits, B, K, R).min()) < NH:
        pytest.skip("a row has fewer than NH depth peaks: the reference's loss rides on topk's order among tied zeros")
    target = synth.pseudo_joints(B, K, seed=seed + 2).double()
    cams = {k: v.double() for k, v in synth.cameras(B, seed=seed + 3).items()}
    params = synth.camera_dict(cams, "cam_0")
    use_sym = any(w is not None for w in (w_bone, w_kp, w_kp2d))

    def go():
        det = multi.KPDetector3DMulti("resnet_multi", K, R, NH, NS)
        x = logits.clone().requires_grad_(True)
        kps, _ = det(x)
        world = torch.stack([util.convert_patch_to_world(kps[:, i], params, "cam_0", is_norm=True) for i in range(NH)], dim=1)
        lp = torch.min(torch.stack([lf.compute_supervision(kps[:, i], target) for i in range(NH)])) * w_mse
        ls = torch.zeros((), dtype=torch.float64)
        if use_sym:
            sym = []
            for i in range(NH):
                t = lf.compute_bone_sym_loss(world[:, i]) * (w_bone or 0.0) + lf.compute_kp_sym_loss(world[:, i]) * (w_kp or 0.0)
                if w_kp2d is not None:
                    t = t + lf.compute_kp_sym_loss(kps[:, i, :, :2], is_3D=False) * 1e2 * w_kp2d
                sym.append(t)
            ls = torch.min(torch.stack(sym))
        (lp + ls).backward()
        return lp.detach(), ls.detach(), x.grad
    r_lp, r_ls, r_grad = ref["mg"].run_in(torch.float64, go)

    x = logits.clone().requires_grad_(True)
    out = oracle.fused_forward(x, K, NH, NS, target, cams, w_mse=w_mse, w_bone=w_bone, w_kp=w_kp, w_kp2d=w_kp2d, reduction="batch")
    lp, ls = out[0], out[1]
    (lp + ls).backward()
    assert abs(lp.item() - r_lp.item()) <= 1e-12 * max(abs(r_lp.item()), 1e-30)
    assert abs(ls.item() - r_ls.item()) <= 1e-12 * max(abs(r_ls.item()), 1e-30)
    assert rel_inf(x.grad.numpy(), r_grad.numpy()) < 1e-10


# --------------------------------------------------------------------------------------------- widening rows
@pytest.mark.parametrize("B,K,S,seed,extension", [(1, 18, 40, 51, True), (2, 18, 64, 52, False), (3, 18, 36, 53, True), (2, 18, 100, 54, True)])
def test_skeleton_rasteriser_and_mask_loss(ref, oracle, synth, B, K, S, seed, extension):
    """util.py:21-59 -> model.py:94 -> loss_func.py:4-16 in its four (weight, use_clip) variants, values and keypoint gradients."""
    util, lf = ref["util"], ref["lf"]
    model = importlib.import_module("modules.model")
    parent, child = model.cal_links(list(synth.H36M_PARENTS), line_select_ids=list(synth.LINE_SELECT), use_root=False, extension=extension)
    assert (list(parent), list(child)) == tuple(map(list, oracle.skeleton_links(synth.H36M_PARENTS, synth.LINE_SELECT, extension=extension)))
    pose = synth.skeleton_pose2d(B, K, seed=seed).double()
    gt = synth.silhouette_mask(synth.skeleton_pose2d(B, K, seed=seed + 1, jitter=0.03), S).double()
    wmap = synth.geodesic_weight(gt.float(), seed=seed + 2).double()
    G = torch.randn(B, 1, S, S, generator=torch.Generator().manual_seed(seed), dtype=torch.float64)

    def run(draw, loss_fn, take_max):
        kp = pose.clone().requires_grad_(True)
        heat = draw(kp)
        recon = take_max(kp, heat)
        res = {"heat": heat.detach(), "recon": recon.detach()}
        res["g_recon"], = torch.autograd.grad((recon * G).sum(), kp, retain_graph=True)
        for vname, use_w, clip in (("mse", False, False), ("clip", False, True), ("w", True, False), ("wclip", True, True)):
            loss = loss_fn(recon, gt, weight=wmap if use_w else None, use_clip=clip)
            res["shape_" + vname] = tuple(loss.shape)
            res["loss_" + vname] = loss.mean().detach()
            res["g_" + vname], = torch.autograd.grad(loss.mean(), kp, retain_graph=True)
        return res
    r = ref["mg"].run_in(torch.float64, lambda: run(lambda kp: util.draw_lines(kp, S, parent, child, synth.BODY_WIDTH),
                                                    lf.compute_mask_reconstruction_loss,
                                                    lambda kp, heat: torch.max(heat.clone(), dim=1, keepdim=True)[0]))
    o = run(lambda kp: oracle.draw_lines(kp, S, parent, child, synth.BODY_WIDTH), oracle.mask_recon_loss,
            lambda kp, heat: oracle.skeleton_mask(kp, S, parent, child, synth.BODY_WIDTH))
    for k in r:
        if k.startswith("shape_"):
            assert o[k] == r[k], k
            continue
        # weight=None + use_clip rides through an fp32 mask in the reference even in its fp64 run (loss_func.py:9-10)
        tol = 2e-5 if k.endswith("_clip") else 1e-10
        assert rel_inf(o[k].numpy(), r[k].numpy()) < tol, k


@pytest.mark.parametrize("B,NH,K,V,seed", [(3, 3, 18, 4, 61), (5, 2, 17, 2, 62), (2, 5, 18, 3, 63)])
def test_eval_selection_triangulation_and_disc_loss(ref, oracle, synth, B, NH, K, V, seed):
    """eval.py:122-148 with eval_utils.switch_points / per_act_mse, util.triangulation, loss_func.compute_disc_loss."""
    util, lf = ref["util"], ref["lf"]
    eu = ref["mg"].load_eval_utils()
    kps, jp = (t.double() for t in synth.eval_predictions(B, NH, K, seed=seed))
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
    o3, o2, otr, oerr, obi, ob2, _ = oracle.eval_select(kps, jp, 256.0, "best")
    assert torch.equal(obi, best_idx) and torch.equal(ob2, best_2d_idx) and torch.equal(otr, tr)
    assert torch.equal(o3, kbest) and torch.equal(o2, k2best)
    assert rel_inf(oerr.numpy(), eu.per_act_mse(k2best, kp_gt[..., :2]).numpy()) < 1e-12

    gen = torch.Generator().manual_seed(400 + seed)
    world = torch.randn(B, K, 3, generator=gen, dtype=torch.float64) * 300
    cams = [{k: v.double() for k, v in synth.cameras(B, seed=seed + 10 + i).items()} for i in range(V)]
    params, kd3 = {}, {}
    for i, c in enumerate(cams):
        params.update(synth.camera_dict(c, "cam_%d" % i))
        kd3["cam_%d" % i] = util.convert_world_to_patch(world, params, "cam_%d" % i, is_norm=True) \
            + 0.002 * torch.randn(B, K, 3, generator=gen, dtype=torch.float64)
    r_tri = ref["mg"].run_in(torch.float64, lambda: util.triangulation(kd3, params, list(range(V))))
    tri = oracle.triangulate([kd3["cam_%d" % i] for i in range(V)], cams)
    assert (tri - r_tri.double()).abs().max().item() < 1e-3        # mm; the reference stores through a float32 buffer (util.py:226)

    p2, g2 = torch.randn(B, 1, generator=gen, dtype=torch.float64), torch.randn(B, 1, generator=gen, dtype=torch.float64)
    p3, g3 = torch.randn(B, NH, 1, generator=gen, dtype=torch.float64), torch.randn(B, NH, 1, generator=gen, dtype=torch.float64)
    for a, b in ((p2, None), (p3, None), (p2, g2), (p3, g3), (p3, g2)):
        assert abs(float(oracle.disc_loss(a, b)) - float(lf.compute_disc_loss(a, b))) < 1e-12
