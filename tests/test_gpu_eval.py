"""GPU parity tests (`-m gpu`, through the C ABI) of the eval-side selection + triangulation and the
discriminator-side glue against the CPU oracle (itself checked against the reference's eval_utils / util /
loss_func in test_oracle_golden.py).  Index outputs (best hypothesis, swap decisions, argmin slots) are
bit-exact; coordinates 1e-6 absolute (pure selection, no arithmetic); triangulated points 1e-5 relative to
the scene scale against the fp64 oracle; losses and gradients 1e-6 relative."""
import importlib

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ev():
    import __graft_entry__ as ge
    ge.build()
    pkg = importlib.import_module("x-as-supervision_b200")
    pkg.load_native()
    return pkg.evalops


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("B,NH,K,mode", [(7, 3, 18, "best"), (64, 3, 17, "best"), (5, 1, 18, "best"), (9, 4, 18, "confident"),
                                         (256, 8, 17, "best")])
def test_eval_select_matches_oracle(ev, oracle, synth, dev, B, NH, K, mode):
    kps, jp = synth.eval_predictions(B, NH, K, seed=B + NH)
    out = ev.eval_select(kps.to(dev), jp.to(dev), 256.0, mode)
    k3, k2, tr, err, bi, b2, gt = oracle.eval_select(kps, jp, 256.0, mode)
    assert torch.equal(out["best_idx"].cpu(), bi) and torch.equal(out["best_2d_idx"].cpu(), b2)
    assert torch.equal(out["is_trans"].cpu(), tr)
    assert torch.equal(out["kp3d"].cpu(), k3) and torch.equal(out["kp2d"].cpu(), k2)       # pure selection: bit-exact
    assert torch.equal(out["gt"].cpu(), gt)
    assert float((out["err2d"].cpu() - err).abs().max()) < 1e-6
    # fp64 oracle makes the same decisions (no near-ties in these inputs)
    k3d, _, trd, errd, bid, b2d, _ = oracle.eval_select(kps.double(), jp.double(), 256.0, mode)
    assert torch.equal(bid, bi) and torch.equal(b2d, b2) and torch.equal(trd, tr)


def test_switch_points_drop_in(ev, oracle, synth, dev):
    kps, jp = synth.eval_predictions(11, 1, 18, seed=3)
    gt = jp.clone()
    gt[..., :2] = gt[..., :2] / 255 * 2 - 1
    gt[..., 2] = gt[..., 2] / 255
    for C in (3, 2):
        res, tr = ev.switch_points(kps[:, 0, :, :C].to(dev), gt[..., :C].to(dev))
        ores, otr = oracle.switch_points(kps[:, 0, :, :C], gt[..., :C])
        assert torch.equal(tr.cpu(), otr) and torch.equal(res.cpu(), ores)


def test_eval_side_against_reference_golden(ev, synth, dev):
    g = load_golden("eval_k18_nh3_v4")
    B, NH, K, V, seed = (int(v) for v in g["meta"])
    kps, jp = synth.eval_predictions(B, NH, K, seed=seed)
    out = ev.eval_select(kps.to(dev), jp.to(dev), 256.0, "best")
    assert np.array_equal(out["best_idx"].cpu().numpy(), g["best_idx_f64"])
    assert np.array_equal(out["best_2d_idx"].cpu().numpy(), g["best_2d_idx_f64"])
    assert np.array_equal(out["is_trans"].cpu().numpy(), g["is_trans_f64"])
    assert np.array_equal(out["kp3d"].cpu().numpy(), g["kp3d_f32"]) and np.array_equal(out["kp2d"].cpu().numpy(), g["kp2d_f32"])
    assert np.abs(out["err2d"].cpu().numpy() - g["err2d_f64"]).max() < 1e-6
    cams = [{k: v.to(dev) for k, v in synth.cameras(B, seed=seed + 10 + i).items()} for i in range(V)]
    tri = ev.triangulate([torch.from_numpy(g["tri_inputs_f32"][i]).to(dev) for i in range(V)], cams)
    # fp32 inputs (the reference's fp32 projections), fp64 solve: within 0.05 mm of the reference's fp64 run on 10^3 mm scenes
    assert np.abs(tri.cpu().numpy() - g["tri_f64"]).max() < 0.05


@pytest.mark.parametrize("V,B,K", [(4, 6, 18), (2, 33, 17), (8, 3, 18)])
def test_triangulation_matches_oracle(ev, oracle, synth, dev, V, B, K):
    g = torch.Generator().manual_seed(V * 100 + B)
    world = torch.randn(B, K, 3, generator=g) * 300
    cams = [synth.cameras(B, seed=70 + i) for i in range(V)]
    kpc = [oracle.world_to_patch(world.double(), {k: v.double() for k, v in c.items()}).float()
           + 0.002 * torch.randn(B, K, 3, generator=g) for c in cams]
    ref = oracle.triangulate([k.double() for k in kpc], [{k: v.double() for k, v in c.items()} for c in cams])
    out = ev.triangulate([k.to(dev) for k in kpc], [{k: v.to(dev) for k, v in c.items()} for c in cams])
    scale = float(ref.abs().max())
    assert float((out.cpu().double() - ref).abs().max()) < 1e-5 * scale
    assert float((out.cpu() - world).abs().max()) < 25.0            # and it is a sensible reconstruction (mm)
    # dict-keyed drop-in signature (util.py:171)
    params, kd = {}, {}
    for i, c in enumerate(cams):
        params.update({k: v.to(dev) for k, v in synth.camera_dict(c, "cam_%d" % i).items()})
        kd["cam_%d" % i] = kpc[i].to(dev)
    out2 = ev.triangulation(kd, params, list(range(V)))
    assert torch.equal(out2, out)


@pytest.mark.parametrize("dim", [3, 2])
@pytest.mark.parametrize("shape", [(6, 3, 18, 3), (5, 18, 3), (1, 1, 17, 3)])
def test_root_centre_matches_the_reference_expression(ev, oracle, dev, dim, shape):
    """model.py:123-124 on the stacked [B,NH,K,3] tensor (relative to hypothesis 0, the reference's literal behaviour) and on
    a [B,K,3] tensor (relative to the root joint), with the VJP."""
    g = torch.Generator().manual_seed(9)
    world = torch.randn(shape, generator=g) * 400
    w = world.to(dev).requires_grad_(True)
    out = ev.root_centre(w, dim)
    w64 = world.double().requires_grad_(True)
    ref = ((w64 - w64[:, [0], :]) / 1000)[..., :dim]                      # the reference's own expression
    assert torch.equal(ref, oracle.root_centre(w64, dim))
    assert tuple(out.shape) == tuple(ref.shape)
    assert float((out.detach().cpu().double() - ref.detach()).abs().max()) <= 1e-6 * max(float(ref.abs().max()), 1e-30)
    G = torch.randn(out.shape, generator=g)
    out.backward(G.to(dev))
    ref.backward(G.double())
    assert float((w.grad.cpu().double() - w64.grad).abs().max()) <= 1e-6 * float(w64.grad.abs().max())
    assert float(out.detach()[:, 0].abs().max()) == 0.0               # item 0 of axis 1 is exactly the origin


@pytest.mark.parametrize("shape", [(16, 1), (16, 3, 1), (5, 4, 2), (256, 3, 1)])
def test_disc_loss_matches_oracle(ev, oracle, dev, shape):
    g = torch.Generator().manual_seed(sum(shape))
    p, q = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
    for gt in (None, q):
        x = p.to(dev).requires_grad_(True)
        y = gt.to(dev).requires_grad_(True) if gt is not None else None
        loss = ev.compute_disc_loss(x, y)
        x64 = p.double().requires_grad_(True)
        y64 = gt.double().requires_grad_(True) if gt is not None else None
        ref = oracle.disc_loss(x64, y64)
        assert abs(float(loss) - float(ref)) < 1e-6 * abs(float(ref))
        loss.backward()
        ref.backward()
        assert float((x.grad.cpu().double() - x64.grad).abs().max()) < 1e-6 * float(x64.grad.abs().max())
        if gt is not None:
            assert float((y.grad.cpu().double() - y64.grad).abs().max()) < 1e-6 * float(y64.grad.abs().max())
    with pytest.raises(ValueError, match="Invalid dimension"):
        ev.compute_disc_loss(torch.zeros(3, device=dev), None)


def test_standalone_pose_terms_match_oracle(oracle, dev):
    """compute_supervision (mean / sum / feature_shape), compute_bone_sym_loss, compute_kp_sym_loss (3-D, 2-D) and their
    gradients against the fp64 oracle (= the reference's loss_func, checked in test_oracle_golden.py)."""
    import __graft_entry__ as ge
    ge.build()
    pkg = importlib.import_module("x-as-supervision_b200")
    pkg.load_native()
    L = pkg.losses
    g = torch.Generator().manual_seed(11)
    B, K = 9, 18
    world = torch.randn(B, K, 3, generator=g) * 400
    kp = torch.rand(B, K, 3, generator=g) * 2 - 1
    gt = torch.rand(B, K, 3, generator=g) * 2 - 1

    def check(ours_fn, ref_fn, x, tol=1e-5):
        a = x.to(dev).requires_grad_(True)
        b = x.double().requires_grad_(True)
        la, lb = ours_fn(a), ref_fn(b)
        assert abs(float(la) - float(lb)) <= tol * abs(float(lb)), (float(la), float(lb))
        la.backward()
        lb.backward()
        assert float((a.grad.cpu().double() - b.grad).abs().max()) <= tol * float(b.grad.abs().max())

    check(lambda a: L.compute_supervision(a, gt.to(dev)), lambda b: oracle.supervision_mse(b, gt.double()), kp)
    check(lambda a: L.compute_supervision(a, gt.to(dev), mode="sum"), lambda b: ((b - gt.double()) ** 2).sum() / B, kp)

    def ref_fs(b):
        k = b.clone()
        k[:, :, :2] = (k[:, :, :2] + 1) / 2.0
        k[:, :, 0] = k[:, :, 0] * (64 - 1)
        k[:, :, 1] = k[:, :, 1] * (48 - 1)
        k[:, :, 2] = k[:, :, 2] * (32 - 1)
        return ((k - gt.double() * 30) ** 2).mean()
    check(lambda a: L.compute_supervision(a, gt.to(dev) * 30, feature_shape=(64, 48, 32)), ref_fs, kp)
    check(L.compute_bone_sym_loss, oracle.bone_sym, world)
    check(lambda a: L.compute_kp_sym_loss(a), lambda b: oracle.kp_sym(b, True), world)
    check(lambda a: L.compute_kp_sym_loss(a, is_3D=False), lambda b: oracle.kp_sym(b, False), kp[..., :2].contiguous())
    with pytest.raises(RuntimeError, match="joints up to 16"):
        L.compute_bone_sym_loss(torch.zeros(2, 10, 3, device=dev))
    # a zero-length bone: the reference's torch.norm has a zero subgradient there (the oracle's sqrt would give NaN), so the
    # gradient must stay finite and equal torch.norm's
    w0 = world.clone()
    w0[0, 16] = w0[0, 15]
    a = w0.to(dev).requires_grad_(True)
    L.compute_bone_sym_loss(a).backward()
    b = w0.double().requires_grad_(True)
    bone = torch.norm(b[:, [16, 15, 13, 12, 3, 2, 6, 5], :] - b[:, [15, 14, 12, 11, 2, 1, 5, 4], :], dim=2) * 1e-3
    torch.nn.functional.mse_loss(bone[:, [0, 2, 4, 6]], bone[:, [1, 3, 5, 7]]).backward()
    assert torch.isfinite(a.grad).all() and torch.isfinite(b.grad).all()
    assert float((a.grad.cpu().double() - b.grad).abs().max()) <= 1e-5 * float(b.grad.abs().max())


def test_exact_ties_follow_the_reference_operators(ev, oracle, dev):
    """Strict comparisons and first-minimum argmins, on inputs built to tie exactly: `torch.lt` in switch_points
    (eval_utils.py:26), `argmin(dim=1)` in the best-hypothesis selection (eval.py:139), `min(dim=1)` in compute_disc_loss
    (loss_func.py:59), `mask > 0.1` in the clip filter (loss_func.py:9)."""
    B, NH, K = 3, 3, 18
    jp = torch.full((B, K, 3), 127.5)                          # gt at the patch centre: normalised (0, 0, 0.5)
    kps = torch.zeros(B, NH, K, 3)
    kps[..., 2] = 0.5
    kps[:, :, 1, 0] = 0.25                                      # joint 1 and its mirror joint 4 are equally far from gt: no swap
    kps[:, :, 4, 0] = -0.25
    kps[:, 1] = kps[:, 0]                                       # hypotheses 0 and 1 identical: argmin must return 0
    kps[:, 2, :, 2] = 0.75
    out = ev.eval_select(kps.to(dev), jp.to(dev), 256.0, "best")
    k3, k2, tr, err, bi, b2, _ = oracle.eval_select(kps, jp, 256.0, "best")
    assert not bool(out["is_trans"].any()) and not bool(tr.any())
    assert int(out["best_idx"].max()) == 0 and torch.equal(out["best_idx"].cpu(), bi) and torch.equal(out["best_2d_idx"].cpu(), b2)
    # discriminator term: equal logits in two slots -> the gradient goes to the first
    x = torch.tensor([[[0.5], [0.5], [2.0]]], device=dev, requires_grad=True)
    ev.compute_disc_loss(x, None).backward()
    assert x.grad[0, 0, 0] != 0 and x.grad[0, 1, 0] == 0 and x.grad[0, 2, 0] == 0
    # clip filter: a mask value of exactly 0.1 is filtered out (strict >)
    sk = importlib.import_module("x-as-supervision_b200.skeleton")
    m = torch.full((1, 1, 4, 4), 0.1, device=dev)
    m[0, 0, 0, 0] = 0.5
    loss = sk.compute_mask_reconstruction_loss(m, torch.zeros_like(m), weight=torch.ones_like(m), use_clip=True)
    assert abs(float(loss) - 0.25 / 16) < 1e-7
