"""GPU parity tests of the skeleton rasteriser + mask-reconstruction loss (`-m gpu`, through the C ABI):
  (1) the golden vectors the reference itself produced (tests/golden/skel_*.npz),
  (2) the fp64 CPU oracle on seeded inputs, incl. degenerate / off-image / coincident-joint skeletons,
  (3) size-independent properties at the BASELINE batch (B=256, 256x256).

Tolerances: heat-map values within 1e-5 absolute (they live in [0,1]); losses 1e-5 relative; keypoint
gradients 1e-5 norm-wise.  Pixels whose value sits within 1e-5 of the 0.1 clip threshold (loss_func.py:9)
may fall on either side in fp32 — the reference's own fp32 run flips them too — so the clipped-loss
comparisons are only made on cases without such pixels (checked on the fp64 oracle's map)."""
import importlib

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_inf
from test_oracle_golden import SKEL_VARIANTS, skeleton_inputs

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def sk():
    import __graft_entry__ as ge
    ge.build()
    pkg = importlib.import_module("x-as-supervision_b200")
    pkg.load_native()
    return pkg.skeleton


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _links(sk, synth, extension=True):
    return sk.cal_links(synth.H36M_PARENTS, synth.LINE_SELECT, use_root=False, extension=extension)


# ------------------------------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("name", ["skel_h36m_s128", "skel_l17_s64"])
def test_rasteriser_against_reference_golden(sk, synth, dev, name):
    g = load_golden(name)
    pose, gt, wmap, G, GH, S = skeleton_inputs(synth, g)
    parent, child = _links(sk, synth, bool(g["meta"][4]))
    assert parent == g["parent"].tolist() and child == g["child"].tolist()
    sub = int(g["meta"][5])
    kp = pose.to(dev).requires_grad_(True)
    heat = sk.draw_lines(kp, S, parent, child, synth.BODY_WIDTH)
    recon = sk.skeleton_mask(kp, S, parent, child, synth.BODY_WIDTH)
    assert np.abs(heat.detach().cpu().numpy()[:, :, ::sub, ::sub] - g["heat_sub_f64"]).max() < TOL
    assert np.abs(recon.detach().cpu().numpy() - g["recon_f64"]).max() < TOL
    gr, = torch.autograd.grad((recon * G.float().to(dev)).sum(), kp, retain_graph=True)
    gh, = torch.autograd.grad((heat * GH.float().to(dev)).sum(), kp, retain_graph=True)
    assert rel_inf(gr.cpu().numpy(), g["g_recon_f64"]) < TOL
    assert rel_inf(gh.cpu().numpy(), g["g_heat_f64"]) < TOL
    # the four (weight, use_clip) variants: stand-alone loss on our recon, and the fused rasterise+loss op
    near = np.abs(g["recon_f64"] - 0.1) < 1e-5
    assert near.sum() == 0, "golden case has pixels on the clip threshold; pick another seed"
    for vname, use_w, clip in SKEL_VARIANTS:
        ref = float(g["loss_%s_f64" % vname])
        tol = 2e-5 if vname == "clip" else TOL        # the reference carries this variant in fp32 even in its fp64 run
        w = wmap.to(dev) if use_w else None
        loss = sk.compute_mask_reconstruction_loss(recon, gt.to(dev), weight=w, use_clip=clip)
        assert (list(loss.shape) or [0]) == g["loss_shape_%s_f64" % vname].tolist(), vname
        assert abs(float(loss.mean()) - ref) < tol * abs(ref), vname
        gl, = torch.autograd.grad(loss.mean(), kp, retain_graph=True)
        assert rel_inf(gl.cpu().numpy(), g["g_loss_%s_f64" % vname]) < tol * 5, vname
        kp2 = pose.to(dev).requires_grad_(True)
        recon2, floss = sk.skeleton_mask_loss(kp2, gt.to(dev), w, S, parent, child, synth.BODY_WIDTH, use_clip=clip)
        assert torch.equal(recon2, recon.detach())
        assert abs(float(floss) - ref) < tol * abs(ref), vname
        floss.backward()
        assert rel_inf(kp2.grad.cpu().numpy(), g["g_loss_%s_f64" % vname]) < tol * 5, vname


# ------------------------------------------------------------------------------------------ fp64 oracle
def _oracle_all(oracle, pose64, S, parent, child, bw, G64, gt64, w64, clip):
    kp = pose64.clone().requires_grad_(True)
    recon = oracle.skeleton_mask(kp, S, parent, child, bw)
    out = {"recon": recon.detach()}
    out["g_recon"], = torch.autograd.grad((recon * G64).sum(), kp, retain_graph=True)
    loss = oracle.mask_recon_loss(recon, gt64, weight=w64, use_clip=clip).mean()
    out["loss"] = float(loss)
    out["g_loss"], = torch.autograd.grad(loss, kp)
    return out


POSES = ["template", "uniform", "tiny", "off_image", "coincident", "huge"]
COINCIDENT = ((1, 2), (11, 12, 13))


def _grad_err(ours, ref64, kind):
    """max |ours - ref| / max |ref|.  Where several joints sit on the same point, the pixels nearest to that point
    are at exactly the same distance from every line touching it, and which of the coincident joints receives their
    gradient is `torch.max`'s tie-break - unspecified in the reference; only the sum over the group is defined."""
    a, b = ours.detach().cpu().double().clone(), ref64.clone()
    if kind == "coincident":
        for grp in COINCIDENT:
            a[:, grp[0]] = a[:, list(grp)].sum(1)
            b[:, grp[0]] = b[:, list(grp)].sum(1)
            a[:, list(grp[1:])] = 0
            b[:, list(grp[1:])] = 0
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


@pytest.mark.parametrize("kind", POSES)
@pytest.mark.parametrize("S", [64, 256, 36, 100])          # 36 and 100: ragged 16x8 edge tiles (S % 16 != 0, S % 8 != 0)
def test_rasteriser_against_oracle(sk, oracle, synth, dev, kind, S):
    B, K = 3, 18
    parent, child = _links(sk, synth)
    g = torch.Generator().manual_seed(40 + len(kind))
    if kind == "template":
        pose = synth.skeleton_pose2d(B, K, seed=41)
    elif kind == "uniform":
        pose = torch.rand(B, K, 2, generator=g) * 1.8 - 0.9
    elif kind == "tiny":                                  # the whole skeleton inside a few pixels
        pose = 0.3 + 0.02 * torch.rand(B, K, 2, generator=g)
    elif kind == "off_image":                             # partly outside the patch
        pose = synth.skeleton_pose2d(B, K, seed=42) * 1.5 + 0.9
    elif kind == "coincident":                            # zero-length segments and shared end points
        pose = synth.skeleton_pose2d(B, K, seed=43)
        pose[:, 2] = pose[:, 1]
        pose[:, 12] = pose[:, 11]
        pose[:, 13] = pose[:, 11]
    else:                                                 # far outside: every pixel underflows to 0
        pose = synth.skeleton_pose2d(B, K, seed=44) + 40.0
    gt = synth.silhouette_mask(synth.skeleton_pose2d(B, K, seed=45), S)
    wmap = synth.geodesic_weight(gt, seed=46)
    G = torch.randn(B, 1, S, S, generator=g)
    kp = pose.to(dev).requires_grad_(True)
    recon = sk.skeleton_mask(kp, S, parent, child, synth.BODY_WIDTH)
    o = _oracle_all(oracle, pose.double(), S, parent, child, synth.BODY_WIDTH, G.double(), gt.double(), wmap.double(), True)
    assert (recon.detach().cpu().double() - o["recon"]).abs().max() < TOL
    gr, = torch.autograd.grad((recon * G.to(dev)).sum(), kp)
    assert _grad_err(gr, o["g_recon"], kind) <= TOL
    # fused loss (SurS1 configuration: weight map + clip), skipping cases with pixels on the clip threshold
    if int(((o["recon"] - 0.1).abs() < 1e-5).sum()) == 0:
        kp2 = pose.to(dev).requires_grad_(True)
        _, loss = sk.skeleton_mask_loss(kp2, gt.to(dev), wmap.to(dev), S, parent, child, synth.BODY_WIDTH, use_clip=True)
        assert abs(float(loss) - o["loss"]) <= TOL * max(abs(o["loss"]), 1e-30)
        loss.backward()
        assert _grad_err(kp2.grad, o["g_loss"], kind) <= TOL
    # un-maxed heat-maps
    heat = sk.draw_lines(pose.to(dev), S, parent, child, synth.BODY_WIDTH)
    oh = oracle.draw_lines(pose.double(), S, parent, child, synth.BODY_WIDTH)
    assert (heat.cpu().double() - oh).abs().max() < TOL
    assert torch.equal(heat.max(dim=1, keepdim=True)[0], recon.detach()), "fused max differs from max of the heat-maps"


def test_strided_view_of_the_head_output_needs_no_copy(sk, oracle, synth, dev):
    """model.py:91 passes kps_ori[cam][:, 0, :, :2], a view of [B,NH,K,3]; gradients flow back into that layout."""
    B, NH, K, S = 2, 3, 18, 64
    parent, child = _links(sk, synth)
    pose = synth.skeleton_pose2d(B, K, seed=47)
    full = torch.zeros(B, NH, K, 3)
    full[:, 0, :, :2] = pose
    full[:, 1:] = 7.0                                     # other hypotheses / z must not be read
    full[..., 2] = -3.0
    kps = full.to(dev).requires_grad_(True)
    recon = sk.skeleton_mask(kps[:, 0, :, :2], S, parent, child, synth.BODY_WIDTH)
    ref = sk.skeleton_mask(pose.to(dev), S, parent, child, synth.BODY_WIDTH)
    assert torch.equal(recon, ref)
    recon.sum().backward()
    assert float(kps.grad[:, 1:].abs().max()) == 0.0 and float(kps.grad[..., 2].abs().max()) == 0.0
    kp = pose.to(dev).requires_grad_(True)
    sk.skeleton_mask(kp, S, parent, child, synth.BODY_WIDTH).sum().backward()
    assert torch.equal(kps.grad[:, 0, :, :2], kp.grad)


@pytest.mark.parametrize("shape", [(2, 1, 64, 64), (3, 5, 7), (1, 1, 10, 10)])
def test_mask_loss_on_arbitrary_tensors(sk, oracle, dev, shape):
    """physique_recons_loss (model.py:176): any mask tensor, all four variants, ragged sizes (n % 4 != 0)."""
    g = torch.Generator().manual_seed(50)
    mask = torch.rand(shape, generator=g)
    gt = (torch.rand(shape, generator=g) > 0.5).float()
    w = torch.rand(shape, generator=g) + 0.5
    mask[(mask - 0.1).abs() < 1e-4] = 0.2
    for vname, use_w, clip in SKEL_VARIANTS:
        m = mask.to(dev).requires_grad_(True)
        loss = sk.compute_mask_reconstruction_loss(m, gt.to(dev), weight=w.to(dev) if use_w else None, use_clip=clip)
        m64 = mask.double().requires_grad_(True)
        ref = oracle.mask_recon_loss(m64, gt.double(), weight=w.double() if use_w else None, use_clip=clip)
        assert tuple(loss.shape) == tuple(ref.shape), vname
        assert rel_inf(loss.detach().cpu().numpy(), ref.detach().numpy()) < TOL, vname
        loss.mean().backward()
        ref.mean().backward()
        assert rel_inf(m.grad.cpu().numpy(), m64.grad.numpy()) < TOL, vname


def test_empty_batch_and_shape_errors(sk, synth, dev):
    parent, child = _links(sk, synth)
    out = sk.skeleton_mask(torch.zeros(0, 18, 2, device=dev), 64, parent, child, synth.BODY_WIDTH)
    assert tuple(out.shape) == (0, 1, 64, 64)
    with pytest.raises(RuntimeError, match="multiple of 4"):
        sk.skeleton_mask(torch.zeros(1, 18, 2, device=dev), 66, parent, child, synth.BODY_WIDTH)
    with pytest.raises(RuntimeError, match="outside"):
        sk.skeleton_mask(torch.zeros(1, 10, 2, device=dev), 64, parent, child, synth.BODY_WIDTH)
    with pytest.raises(ValueError, match="lines"):
        sk.skeleton_mask(torch.zeros(1, 18, 2, device=dev), 64, list(range(18)) * 2, list(range(18)) * 2, synth.BODY_WIDTH)


# ------------------------------------------------------------------------------------------ BASELINE size
def test_full_size_properties(sk, synth, dev):
    """B=256, 256x256 (the batch of BASELINE configs[1]): properties that need no oracle."""
    B, K, S = 256, 18, 256
    parent, child = _links(sk, synth)
    pose = synth.skeleton_pose2d(B, K, seed=60).to(dev)
    gt = synth.silhouette_mask(synth.skeleton_pose2d(32, K, seed=61), S).repeat(B // 32, 1, 1, 1).to(dev)
    wmap = synth.geodesic_weight(gt.cpu()[:32], seed=62).repeat(B // 32, 1, 1, 1).to(dev)

    def run(p):
        kp = p.clone().requires_grad_(True)
        recon, loss = sk.skeleton_mask_loss(kp, gt, wmap, S, parent, child, synth.BODY_WIDTH, use_clip=True)
        loss.backward()
        return recon.detach(), loss.detach(), kp.grad

    recon, loss, grad = run(pose)
    # values are probabilities of a Gaussian profile: in [0,1], exactly 1 only on a joint pixel, max near the limbs
    assert float(recon.min()) >= 0.0 and float(recon.max()) <= 1.0 and float(recon.amax(dim=(1, 2, 3)).min()) > 0.9
    # bit-identical reruns (no float atomics anywhere)
    recon2, loss2, grad2 = run(pose)
    assert torch.equal(recon, recon2) and torch.equal(loss, loss2) and torch.equal(grad, grad2)
    # samples are independent: a permuted batch gives permuted maps and gradients (the loss mean is unchanged up to order)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).to(dev)
    recon_p, _, _ = run(pose[perm])
    assert torch.equal(recon_p, recon[perm])
    # mirror symmetry: flipping x of every joint mirrors the map exactly (the grid is symmetric: x_j = -x_{S-1-j} in fp32
    # up to rounding, so allow 1e-6) and negates the x-gradient of the un-weighted, un-clipped loss
    flip = pose * torch.tensor([-1.0, 1.0], device=dev)
    r1 = sk.skeleton_mask(pose[:8], S, parent, child, synth.BODY_WIDTH)
    r2 = sk.skeleton_mask(flip[:8], S, parent, child, synth.BODY_WIDTH)
    assert float((r1 - r2.flip(-1)).abs().max()) < 2e-5
    # translation equivariance by whole pixels: shifting all joints by k*delta shifts the map by k pixels
    delta = 2.0 / (S - 1)
    r3 = sk.skeleton_mask(pose[:8] + torch.tensor([8 * delta, 0.0], device=dev), S, parent, child, synth.BODY_WIDTH)
    assert float((r3[..., 8:] - r1[..., :-8]).abs().max()) < 2e-5
    # the fused max equals the max of the un-maxed heat-maps, bit for bit
    heat = sk.draw_lines(pose[:4], S, parent, child, synth.BODY_WIDTH)
    assert torch.equal(heat.max(dim=1, keepdim=True)[0], recon[:4])
    # linearity of the backward in the upstream gradient
    kp = pose[:16].clone().requires_grad_(True)
    r = sk.skeleton_mask(kp, S, parent, child, synth.BODY_WIDTH)
    G1 = torch.randn_like(r)
    G2 = torch.randn_like(r)
    g1, = torch.autograd.grad(r, kp, G1, retain_graph=True)
    g2, = torch.autograd.grad(r, kp, G2, retain_graph=True)
    g12, = torch.autograd.grad(r, kp, G1 + 2 * G2)
    assert float((g12 - (g1 + 2 * g2)).abs().max()) <= 1e-4 * float(g12.abs().max())
