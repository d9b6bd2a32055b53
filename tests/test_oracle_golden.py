"""Pin the CPU oracle against outputs of the reference itself (tests/golden/*.npz,
made by tests/golden/make_golden.py from /root/reference).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import input_checksum, load_golden, rel_inf, rel_l2

HEAD_CASES = [
    ("head_iid_k18_r16", "iid_logits"),
    ("head_blob_k17_r32", "blob_logits"),
    ("head_blob_k18_r64", "blob_logits"),
    ("head_iid_k3_r8_nh2", "iid_logits"),
]


def _defined_slots(num_peaks, NH):
    """mask[B,K,NH]: slot h is defined by the reference iff the row has > h local maxima."""
    return np.arange(NH)[None, None, :] < num_peaks[..., None]


@pytest.mark.parametrize("name,gen", HEAD_CASES)
# fp64 is the strict pin.  fp32: the reference's CPU F.softmax uses a ~1e-5-accurate vector exp on long
# rows, so its own fp32-vs-fp64 self-error reaches 1.7e-4 on peaked logits -> 5e-4.
@pytest.mark.parametrize("tag,dtype,tol", [("f64", torch.float64, 1e-12), ("f32", torch.float32, 5e-4)])
def test_head_forward_matches_reference(oracle, synth, name, gen, tag, dtype, tol):
    g = load_golden(name)
    B, K, D, H, W, NH, NS, seed, _ = [int(v) for v in g["meta"]]
    logits = getattr(synth, gen)(B, K, D, H, W, seed=seed)
    np.testing.assert_allclose(input_checksum(logits), g["in_checksum"], rtol=1e-12)
    kps, dmap, idx = oracle.integral_multi(logits.to(dtype), K, NH, NS)
    defined = _defined_slots(g["num_peaks"], NH)
    ref_idx = g["idx_" + tag]
    # bit-exact peak bins wherever the reference defines them
    assert np.array_equal(idx.numpy()[defined], ref_idx[defined])
    assert rel_inf(dmap.numpy(), g["dmap_" + tag]) < tol
    ref_kps = g["kps_" + tag]
    mask = np.broadcast_to(defined.transpose(0, 2, 1)[..., None], ref_kps.shape)
    assert np.abs(kps.numpy() - ref_kps)[mask].max() < tol
    # x,y do not depend on the slot at all
    assert np.abs(kps.numpy()[..., :2] - ref_kps[..., :2]).max() < tol


@pytest.mark.parametrize("name,gen", HEAD_CASES)
def test_head_backward_matches_reference(oracle, synth, name, gen):
    """autograd through the oracle AND the closed form (App. A.2) against the reference's autograd, fp64."""
    g = load_golden(name)
    B, K, D, H, W, NH, NS, seed, stride = [int(v) for v in g["meta"]]
    if not _defined_slots(g["num_peaks"], NH).all():
        pytest.skip("filler slots present: reference gradient routes through undefined indices")
    logits = getattr(synth, gen)(B, K, D, H, W, seed=seed).double().requires_grad_(True)
    gw = torch.from_numpy(g["g_kps"])
    kps, _, _ = oracle.integral_multi(logits, K, NH, NS)
    (kps * gw).sum().backward()
    auto = logits.grad.flatten().numpy()
    closed = oracle.integral_multi_backward(logits.detach(), gw, K, NH, NS).flatten().numpy()
    ref = g["grad_sub_f64"]
    assert rel_inf(auto[::stride], ref) < 1e-11
    assert rel_inf(closed[::stride], ref) < 1e-11
    np.testing.assert_allclose([np.abs(closed).max(), np.linalg.norm(closed)], g["grad_norms_f64"][:2], rtol=1e-10)


def test_single_hypothesis_head(oracle, synth):
    g = load_golden("single_iid_k18_r16")
    B, K, R, seed = [int(v) for v in g["meta"]]
    logits = synth.iid_logits(B, K, R, R, R, seed=seed)
    for tag, dt, tol in (("f64", torch.float64, 1e-12), ("f32", torch.float32, 2e-5)):
        kps, dmap = oracle.integral_single(logits.to(dt), K)
        assert kps.shape == (B, 1, K, 3)
        assert np.abs(kps.numpy() - g["kps_" + tag]).max() < tol
        assert rel_inf(dmap.numpy(), g["dmap_" + tag]) < tol


@pytest.mark.parametrize("name", ["geom_h36m", "geom_mpi"])
def test_geometry_matches_reference(oracle, synth, name):
    g = load_golden(name)
    B, K, seed, mpi = [int(v) for v in g["meta"]]
    cams = synth.cameras(B, seed=seed, mpi=bool(mpi))
    kps = synth.pseudo_joints(B, K, seed=seed + 1)
    for tag, dt, tol in (("f64", torch.float64, 1e-11), ("f32", torch.float32, 2e-5)):
        c = {k: v.to(dt) for k, v in cams.items()}
        k = kps.to(dt)
        world = oracle.patch_to_world(k, c)
        assert rel_inf(world.numpy(), g["world_" + tag]) < tol
        img = oracle.patch_to_image(k, c["trans_image"], 256, 256, 256, 2000.0 / 256, c["pelvis"])
        assert rel_inf(img.numpy(), g["image_" + tag]) < tol
        mono = oracle.patch_to_world(k, c, rect_width=256, mono=True, patch=False)
        assert rel_inf(mono.numpy(), g["mono_" + tag]) < tol
        back = oracle.world_to_patch(torch.from_numpy(g["world_" + tag]), c)
        # fp32: the round trip amplifies rounding by Z/f ~ 1e3 px -> looser
        assert np.abs(back.numpy() - g["back_" + tag]).max() < (1e-9 if dt == torch.float64 else 5e-3)
    # projection really is the inverse (fp64)
    c = {k: v.double() for k, v in cams.items()}
    rt = oracle.world_to_patch(oracle.patch_to_world(kps.double(), c), c)
    assert np.abs(rt.numpy() - kps.double().numpy()).max() < 1e-9


LOSS_CASES = [("loss_surs1_k17_r16", "iid_logits"), ("loss_synths2_k18_r32", "blob_logits"),
              ("loss_synths2_k17_r32_mpi", "iid_logits")]


@pytest.mark.parametrize("name,gen", LOSS_CASES)
def test_fused_loss_matches_reference(oracle, synth, name, gen):
    g = load_golden(name)
    B, K, R, NH, NS, seed, stride, mpi = [int(v) for v in g["meta"]]
    w = [None if np.isnan(v) else float(v) for v in g["weights"]]
    logits = getattr(synth, gen)(B, K, R, R, R, seed=seed)
    target = synth.pseudo_joints(B, K, seed=seed + 2)
    cams = synth.cameras(B, seed=seed + 3, mpi=bool(mpi))
    for tag, dt, tol, gtol in (("f64", torch.float64, 1e-11, 1e-10), ("f32", torch.float32, 5e-4, 5e-3)):
        x = logits.to(dt).clone().requires_grad_(True)
        c = {k: v.to(dt) for k, v in cams.items()}
        lp, ls, sel, kps, world, _, _ = oracle.fused_forward(
            x, K, NH, NS, target.to(dt), c, w_mse=w[0], w_bone=w[1], w_kp=w[2], w_kp2d=w[3], reduction="batch")
        (lp + ls).backward()
        assert rel_inf(kps.detach().numpy(), g["kps_" + tag]) < tol
        assert rel_inf(world.detach().numpy(), g["world_" + tag]) < max(tol, 1e-9) * (1 if dt == torch.float64 else 30)
        np.testing.assert_allclose([lp.item(), ls.item()], g["loss_" + tag], rtol=tol * 10, atol=1e-12)
        # selected slots are bit-exact
        assert int(sel[0]) == int(np.argmin(g["pseudo_h_" + tag]))
        if w[1] is not None or w[2] is not None or w[3] is not None:
            assert int(sel[1]) == int(np.argmin(g["sym_h_" + tag]))
        assert rel_inf(x.grad.flatten().numpy()[::stride], g["grad_sub_" + tag]) < gtol
    # eval-style per-joint argmin (eval.py:138-145)
    idx, best = oracle.best_hypothesis(torch.from_numpy(g["kps_f64"]), target.double())
    assert np.array_equal(idx.numpy(), g["best_idx_f64"])
    lj, _, selj, _ = oracle.reproj_min_loss(torch.from_numpy(g["kps_f64"]), target.double(),
                                            {k: v.double() for k, v in cams.items()}, reduction="joint")
    assert np.array_equal(selj.numpy(), g["best_idx_f64"])


def test_filler_slots_are_deterministic(oracle):
    """Rows with fewer than NH local maxima: ours fills with the lowest non-peak interior bins."""
    pz = torch.tensor([[[0.05, 0.1, 0.5, 0.1, 0.05, 0.05, 0.1, 0.05]]], dtype=torch.float64)
    pz = pz / pz.sum()
    idx = oracle.depth_peaks(pz, 4)
    assert idx.tolist() == [[[2, 6, 1, 3]]]
    flat = torch.full((1, 1, 8), 0.125, dtype=torch.float64)          # plateau: every interior bin is a peak
    assert oracle.depth_peaks(flat, 3).tolist() == [[[1, 2, 3]]]


# --------------------------------------------------------------------------- skeleton rasteriser + mask loss
SKEL_VARIANTS = (("mse", False, False), ("clip", False, True), ("w", True, False), ("wclip", True, True))


def skeleton_inputs(synth, g):
    B, K, S, seed = (int(v) for v in g["meta"][:4])
    pose = synth.skeleton_pose2d(B, K, seed=seed)
    gt = synth.silhouette_mask(synth.skeleton_pose2d(B, K, seed=seed + 1, jitter=0.03), S)
    wmap = synth.geodesic_weight(gt, seed=seed + 2)
    np.testing.assert_allclose(input_checksum(pose), g["in_checksum"], rtol=1e-12)
    np.testing.assert_allclose(input_checksum(gt), g["gt_checksum"], rtol=1e-12)
    np.testing.assert_allclose(input_checksum(wmap), g["w_checksum"], rtol=1e-12)
    gen = torch.Generator().manual_seed(300 + seed)
    G = torch.randn(B, 1, S, S, generator=gen, dtype=torch.float64)
    GH = torch.randn(B, int(g["meta"][6]), S, S, generator=gen, dtype=torch.float64)
    return pose, gt, wmap, G, GH, S


@pytest.mark.parametrize("name", ["skel_h36m_s128", "skel_l17_s64"])
@pytest.mark.parametrize("tag,dtype,tol", [("f64", torch.float64, 1e-11), ("f32", torch.float32, 2e-5)])
def test_skeleton_rasteriser_matches_reference(oracle, synth, name, tag, dtype, tol):
    g = load_golden(name)
    pose, gt, wmap, G, GH, S = skeleton_inputs(synth, g)
    parent, child = oracle.skeleton_links(synth.H36M_PARENTS, synth.LINE_SELECT, extension=bool(g["meta"][4]))
    assert parent == g["parent"].tolist() and child == g["child"].tolist()
    sub = int(g["meta"][5])
    kp = pose.to(dtype).requires_grad_(True)
    heat = oracle.draw_lines(kp, S, parent, child, synth.BODY_WIDTH)
    recon = oracle.skeleton_mask(kp, S, parent, child, synth.BODY_WIDTH)
    assert np.abs(heat.detach().numpy()[:, :, ::sub, ::sub] - g["heat_sub_" + tag]).max() < tol
    assert np.abs(recon.detach().numpy() - g["recon_" + tag]).max() < tol
    gr, = torch.autograd.grad((recon * G.to(dtype)).sum(), kp, retain_graph=True)
    gh, = torch.autograd.grad((heat * GH.to(dtype)).sum(), kp, retain_graph=True)
    assert rel_inf(gr.numpy(), g["g_recon_" + tag]) < tol * 10
    assert rel_inf(gh.numpy(), g["g_heat_" + tag]) < tol * 10
    for vname, use_w, clip in SKEL_VARIANTS:
        loss = oracle.mask_recon_loss(recon, gt.to(dtype), weight=wmap.to(dtype) if use_w else None, use_clip=clip)
        assert (list(loss.shape) or [0]) == g["loss_shape_%s_%s" % (vname, tag)].tolist(), vname   # the tensor-valued quirk
        # weight=None + use_clip: the reference multiplies its 0-d fp64 MSE by `(mask > 0.1).float()`, an fp32
        # tensor, so even its fp64 run carries this variant in fp32 (loss_func.py:9-10)
        vtol = max(tol, 2e-6) if vname == "clip" else tol
        ref = float(g["loss_%s_%s" % (vname, tag)])
        assert abs(float(loss.mean().detach()) - ref) < vtol * 10 * abs(ref), vname
        gl, = torch.autograd.grad(loss.mean(), kp, retain_graph=True)
        assert rel_inf(gl.numpy(), g["g_loss_%s_%s" % (vname, tag)]) < vtol * 50, vname


# --------------------------------------------------------------------------- eval selection, triangulation, discriminator loss
@pytest.mark.parametrize("tag,dtype,tol", [("f64", torch.float64, 1e-9), ("f32", torch.float32, 1e-6)])
def test_eval_side_matches_reference(oracle, synth, tag, dtype, tol):
    g = load_golden("eval_k18_nh3_v4")
    B, NH, K, V, seed = (int(v) for v in g["meta"])
    kps, jp = synth.eval_predictions(B, NH, K, seed=seed)
    np.testing.assert_allclose(input_checksum(kps), g["in_checksum"], rtol=1e-12)
    np.testing.assert_allclose(input_checksum(jp), g["jp_checksum"], rtol=1e-12)
    k3, k2, tr, err, bi, b2, gt = oracle.eval_select(kps.to(dtype), jp.to(dtype), 256.0, "best")
    assert np.array_equal(bi.numpy(), g["best_idx_" + tag]) and np.array_equal(b2.numpy(), g["best_2d_idx_" + tag])
    assert np.array_equal(tr.numpy(), g["is_trans_" + tag]) and g["is_trans_" + tag].sum() > 0
    assert np.array_equal(k3.numpy(), g["kp3d_" + tag]) and np.array_equal(k2.numpy(), g["kp2d_" + tag])
    assert np.abs(err.numpy() - g["err2d_" + tag]).max() < tol
    _, _, _, err0, bi0, _, _ = oracle.eval_select(kps.to(dtype), jp.to(dtype), 256.0, "confident")
    assert np.abs(err0.numpy() - g["err2d_h0_" + tag]).max() < tol and int(bi0.abs().max()) == 0
    # triangulation on the reference's own camera inputs
    gen = torch.Generator().manual_seed(400 + seed)
    world = torch.randn(B, K, 3, generator=gen) * 300
    np.testing.assert_allclose(input_checksum(world), g["world_checksum"], rtol=1e-12)
    cams = [{k: v.to(dtype) for k, v in synth.cameras(B, seed=seed + 10 + i).items()} for i in range(V)]
    noise = [0.002 * torch.randn(B, K, 3, generator=gen) for _ in range(V)]
    kpc = [oracle.world_to_patch(world.to(dtype), c) + n.to(dtype) for c, n in zip(cams, noise)]
    tri = oracle.triangulate(kpc, cams)
    # the reference writes its result through a float32 buffer (util.py:226) and its fp32 SVD of the badly scaled
    # DLT matrix is itself only ~1e-3 mm accurate, so: fp64 run to 1e-3 mm of 10^3 mm, fp32 run to 0.5 mm
    assert np.abs(tri.numpy() - g["tri_" + tag]).max() < (1e-3 if tag == "f64" else 0.5)
    logits = {"p2": torch.randn(B, 1, generator=gen), "g2": torch.randn(B, 1, generator=gen),
              "p3": torch.randn(B, NH, 1, generator=gen), "g3": torch.randn(B, NH, 1, generator=gen)}
    L = {k: v.to(dtype) for k, v in logits.items()}
    ours = [oracle.disc_loss(L["p2"], None), oracle.disc_loss(L["p3"], None), oracle.disc_loss(L["p2"], L["g2"]),
            oracle.disc_loss(L["p3"], L["g3"]), oracle.disc_loss(L["p3"], L["g2"])]
    np.testing.assert_allclose(np.array([float(v) for v in ours]), g["disc_" + tag], rtol=tol * 10)


# --------------------------------------------------------------------------- the whole loss graph from the oracle's pieces
@pytest.mark.parametrize("name", ["model_synths2_k18_r32", "model_surs1_k18_r32"])
def test_oracle_pieces_reproduce_the_reference_model(oracle, synth, name):
    """loss_values of the reference's Counter3DModel.forward / Counter3DDisc.forward (golden) rebuilt from the oracle's
    functions in fp64: head, per-hypothesis world lift, symmetry / pseudo min over hypotheses, the literal root-centring,
    the rasterised mask loss.  Pins the composition (which tensor feeds which term) as well as the pieces."""
    from torch import nn
    g = load_golden(name)
    B, K, R, NH, NS, seed, stride, use_dis_map = (int(v) for v in g["meta"])
    sym = None if np.isnan(g["sym"]).all() else tuple(float(v) for v in g["sym"])
    cfg = synth.model_cfg(sym=sym, use_dis_map=bool(use_dis_map))["loss_config"]
    batch = {k: (v.double() if v.is_floating_point() else v) for k, v in synth.model_batch(B, K, R, seed=seed).items()}
    gen = torch.Generator().manual_seed(7)
    lin = nn.Linear(K * 3, 1)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(1, K * 3, generator=gen, dtype=torch.float32) * 0.5)
        lin.bias.fill_(0.25)
    disc = nn.Sequential(nn.Flatten(), lin).double()
    parent, child = oracle.skeleton_links(synth.H36M_PARENTS, synth.LINE_SELECT)
    tot = {"symmetry": 0.0, "smpl_gen": 0.0, "smpl_pseudo_img": 0.0, "reconstruction": 0.0, "disc": 0.0}
    for c in (0, 1):
        key = "cam_%d" % c
        cams = {k: batch[key + "_" + k] for k in ("trans_image", "pelvis", "k_mat", "trans_world", "rot_world")}
        kps, _, _ = oracle.integral_multi(batch[key + "_img"], K, NH, NS)
        if sym is not None:
            _, ls, _, world = oracle.reproj_min_loss(kps, torch.zeros(B, K, 3, dtype=torch.float64), cams, img_hw=(R, R), w_mse=0.0,
                                                     w_bone=sym[0], w_kp=sym[1], w_kp2d=sym[2], reduction="batch")
            tot["symmetry"] += float(ls)
        else:
            world = torch.stack([oracle.patch_to_world(kps[:, i], cams, (R, R), True, 2000.0) for i in range(NH)], 1)
        cen = oracle.root_centre(world, 3)
        tot["smpl_gen"] += float(oracle.disc_loss(torch.stack([disc(cen[:, i]) for i in range(NH)], 1), None)) * cfg["smpl_gen_loss"]["weight"]
        pk, _, _ = oracle.integral_multi(batch[key + "_pseudo_img"], K, NH, NS)
        lp, _, _, _ = oracle.reproj_min_loss(pk, batch[key + "_pseudo_joints"], cams, img_hw=(R, R), w_mse=1.0, reduction="batch")
        tot["smpl_pseudo_img"] += float(lp) * cfg["smpl_pseudo_img_loss"]["weight"]
        recon = oracle.skeleton_mask(kps[:, 0, :, :2], R, parent, child, synth.BODY_WIDTH)
        w = batch[key + "_geodesic_dis"] if use_dis_map else None
        tot["reconstruction"] += float(oracle.mask_recon_loss(recon, batch[key + "_mask"], weight=w, use_clip=True).mean()) * cfg["recons_loss"]["weight"]
        pred = torch.stack([disc(kps[:, i]) for i in range(NH)], 1)
        tot["disc"] += float(oracle.disc_loss(pred, disc(batch[key + "_pseudo_joints"]))) * cfg["smpl_disc_loss"]["weight"]
    for k, v in tot.items():
        ref = g.get("loss_%s_f64" % k)
        if ref is None:
            assert k == "symmetry" and sym is None
            continue
        # the clip variant without a weight map is carried in fp32 by the reference even in its fp64 run (see above)
        tol = 2e-6 if (k == "reconstruction" and not use_dis_map) else 1e-10
        assert abs(v - float(ref)) <= tol * abs(float(ref)), (k, v, float(ref))
