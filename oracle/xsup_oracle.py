"""CPU oracle for the integral + multi-hypothesis reprojection-loss path.

TEST INFRASTRUCTURE ONLY.  Nothing under `x-as-supervision_b200/` imports this
file; only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may.  It is the checker, never the thing
shipped or measured as the product.

It is a from-scratch restatement, in plain torch CPU ops, of the algorithm in
  * modules/keypoint_detector_integral_multi.py:24-88   (multi-hypothesis head tail)
  * modules/keypoint_detector_integral.py:21-65         (single-hypothesis variant)
  * modules/util.py:61-168                              (camera geometry, both directions)
  * modules/base_losses/loss_func.py:18-52              (MSE / bone / keypoint symmetry)
  * modules/model.py:71-79,105-114,158-162, eval.py:138-145 (min / argmin over hypotheses)
of /root/reference.  Every function works in the dtype of its inputs, so the
same code is the fp32 comparator and the fp64 ground truth.

Parity pin: the reference has no tests or golden vectors of its own
(SURVEY.md §4), so this oracle is pinned against *outputs of the reference
itself*, generated in the build container by `tests/golden/make_golden.py`
(which imports /root/reference) and committed under `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks every function here against them.

One place is deliberately *defined* where the reference is undefined: when a
depth row has fewer than `num_hypo` local maxima `torch.topk` returns arbitrary
tied indices (all candidates are 0.0).  Here ties are broken by ascending bin
index, which is also what the CUDA kernels do.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch

# bone / midpoint index tables of loss_func.py:20,28,33
BONE_CHILD = (16, 15, 13, 12, 3, 2, 6, 5)
BONE_PARENT = (15, 14, 12, 11, 2, 1, 5, 4)
MID_A = (11, 1)
MID_B = (14, 4)


# --------------------------------------------------------------------------- head tail
def softmax_volume(logits: torch.Tensor, num_kp: int) -> torch.Tensor:
    """`[B, K*D, H, W]` logits -> probabilities `[B, K, D, H, W]`, softmax over each
    joint's whole D*H*W volume (keypoint_detector_integral_multi.py:69-74)."""
    B, C, H, W = logits.shape
    D = C // num_kp
    # the same ATen kernel the reference calls (F.softmax, :71), so the fp32 mode of this oracle has the
    # reference's own numerics and CPU cost; fp64 mode is the ground truth
    return torch.softmax(logits.reshape(B, num_kp, D * H * W), dim=2).reshape(B, num_kp, D, H, W)


def marginals(p: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(ax[B,K,W], ay[B,K,H], pz[B,K,D]) as in :39-44."""
    over_d = p.sum(dim=2)                                             # [B,K,H,W]
    return over_d.sum(dim=2), over_d.sum(dim=3), p.sum(dim=3).sum(dim=3)


def depth_peaks(pz: torch.Tensor, num_hypo: int) -> torch.Tensor:
    """Bins of the `num_hypo` largest non-strict interior local maxima of `pz[B,K,D]`,
    value-descending, as original bin indices (int64) — find_peak, :24-34.  Ties
    (including the all-zero filler candidates) resolve to the lowest bin."""
    mid = pz[..., 1:-1]
    is_peak = (mid >= pz[..., :-2]) & (mid >= pz[..., 2:])
    cand = torch.where(is_peak, mid, torch.zeros_like(mid))
    order = torch.sort(cand, dim=-1, descending=True, stable=True).indices
    return order[..., :num_hypo] + 1


def window_depth(pz: torch.Tensor, idx: torch.Tensor, neighbor_size: int) -> torch.Tensor:
    """Depth expectation inside a `neighbor_size` window centred on each peak bin:
    the two zero-padded, count-include-pad average pools of :57-62 gathered at `idx`."""
    D = pz.shape[-1]
    half = neighbor_size // 2
    bins = torch.arange(D, dtype=pz.dtype, device=pz.device)
    pad = torch.zeros(pz.shape[:-1] + (half,), dtype=pz.dtype, device=pz.device)
    num = torch.cat((pad, pz * bins, pad), dim=-1).unfold(-1, neighbor_size, 1).sum(-1) / neighbor_size
    den = torch.cat((pad, pz, pad), dim=-1).unfold(-1, neighbor_size, 1).sum(-1) / neighbor_size
    return torch.gather(num, -1, idx) / torch.gather(den, -1, idx)


def integral_multi(logits: torch.Tensor, num_kp: int, num_hypo: int, neighbor_size: int
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Tail of KPDetector3DMulti.forward (:69-88).

    Returns `kps[B,NH,K,3]`, `depth_prob_map[K,D]` (sample 0, :48) and the peak bins
    `[B,K,NH]` int64.  The reference's quirks are kept: the x expectation is
    normalised by H and y by W (:78-79), the normaliser is the size, not size-1,
    and the depth axis must equal the width (`arange(z_dim=W)` multiplies a
    length-D vector, :57)."""
    B, C, H, W = logits.shape
    D = C // num_kp
    if D != W:
        raise ValueError("reference semantics need depth_dim == width (got D=%d, W=%d)" % (D, W))
    p = softmax_volume(logits, num_kp)
    ax, ay, pz = marginals(p)
    xbar = (ax * torch.arange(W, dtype=p.dtype, device=p.device)).sum(-1, keepdim=True)
    ybar = (ay * torch.arange(H, dtype=p.dtype, device=p.device)).sum(-1, keepdim=True)
    idx = depth_peaks(pz, num_hypo)
    zwin = window_depth(pz, idx, neighbor_size)                      # [B,K,NH]
    x = xbar / H * 2 - 1
    y = ybar / W * 2 - 1
    z = zwin / D * 2 - 1
    kps = torch.stack((x.expand(-1, -1, num_hypo), y.expand(-1, -1, num_hypo), z), dim=-1)  # [B,K,NH,3]
    return kps.permute(0, 2, 1, 3).contiguous(), pz[0].clone(), idx


def integral_single(logits: torch.Tensor, num_kp: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Tail of KPDetector3D.forward (keypoint_detector_integral.py:45-65): global depth
    expectation, output `[B,1,K,3]`."""
    B, C, H, W = logits.shape
    D = C // num_kp
    if D != W:
        raise ValueError("reference semantics need depth_dim == width")
    p = softmax_volume(logits, num_kp)
    ax, ay, pz = marginals(p)
    x = (ax * torch.arange(W, dtype=p.dtype, device=p.device)).sum(-1) / H * 2 - 1
    y = (ay * torch.arange(H, dtype=p.dtype, device=p.device)).sum(-1) / W * 2 - 1
    z = (pz * torch.arange(D, dtype=p.dtype, device=p.device)).sum(-1) / D * 2 - 1
    return torch.stack((x, y, z), dim=-1).unsqueeze(1), pz[0].clone()


def integral_multi_backward(logits: torch.Tensor, g_kps: torch.Tensor, num_kp: int, num_hypo: int,
                            neighbor_size: int) -> torch.Tensor:
    """Closed-form d loss / d logits for `integral_multi` given `g_kps[B,NH,K,3]`
    (SURVEY.md App. A.2).  Independent of autograd; used to cross-check both the
    autograd of this file and the CUDA backward kernel."""
    B, C, H, W = logits.shape
    D = C // num_kp
    half = neighbor_size // 2
    p = softmax_volume(logits, num_kp)
    ax, ay, pz = marginals(p)
    idx = depth_peaks(pz, num_hypo)                                   # [B,K,NH]
    dt = p.dtype
    bins = torch.arange(D, dtype=dt, device=logits.device)
    g = g_kps.to(dt).permute(0, 2, 1, 3)                              # [B,K,NH,3]
    a = g[..., 0].sum(-1) * (2.0 / H)                                 # on w
    b = g[..., 1].sum(-1) * (2.0 / W)                                 # on h
    inwin = ((bins.view(1, 1, 1, D) - idx.unsqueeze(-1).to(dt)).abs() <= half).to(dt)   # [B,K,NH,D]
    swin = (inwin * pz.unsqueeze(2)).sum(-1)                          # [B,K,NH]
    zbar = (inwin * (pz * bins).unsqueeze(2)).sum(-1) / swin
    c = (g[..., 2].unsqueeze(-1) * (2.0 / D) * inwin * (bins.view(1, 1, 1, D) - zbar.unsqueeze(-1))
         / swin.unsqueeze(-1)).sum(2)                                 # [B,K,D]
    xbar = (ax * torch.arange(W, dtype=dt, device=logits.device)).sum(-1)
    ybar = (ay * torch.arange(H, dtype=dt, device=logits.device)).sum(-1)
    gbar = a * xbar + b * ybar + (c * pz).sum(-1)
    field = (a.view(B, num_kp, 1, 1, 1) * torch.arange(W, dtype=dt, device=logits.device).view(1, 1, 1, 1, W)
             + b.view(B, num_kp, 1, 1, 1) * torch.arange(H, dtype=dt, device=logits.device).view(1, 1, 1, H, 1)
             + c.view(B, num_kp, D, 1, 1) - gbar.view(B, num_kp, 1, 1, 1))
    return (p * field).reshape(B, C, H, W)


# --------------------------------------------------------------------------- geometry
def _affine_parts(trans_image: torch.Tensor):
    return trans_image[..., :, :2], trans_image[..., :, 2]


def patch_to_image(kps, trans_image, img_d, img_h, img_w, depth_scale, pelvis, is_norm=True):
    """Normalised patch coords -> image px / depth mm (util.py:61-82): undo the [-1,1]
    normalisation with (size-1), apply the inverse crop affine, px-depth -> mm + pelvis depth."""
    x, y, z = kps[..., 0], kps[..., 1], kps[..., 2]
    if is_norm:
        x = (x + 1) / 2.0 * (img_w - 1)
        y = (y + 1) / 2.0 * (img_h - 1)
        z = z * (img_d - 1)
    A, t = _affine_parts(trans_image)
    Ainv = torch.linalg.inv(A)                                        # [B,2,2]
    uv = torch.stack((x, y), dim=-1) - t.unsqueeze(1)                 # [B,K,2]
    uv = torch.einsum("bij,bkj->bki", Ainv, uv)
    Z = z * depth_scale + pelvis[..., 2].unsqueeze(1)
    return torch.cat((uv, Z.unsqueeze(-1)), dim=-1)


def image_to_world(kps, fx, fy, cx, cy, trans_world, rot_world):
    """Pinhole back-projection then inverse extrinsics (util.py:85-95)."""
    Z = kps[..., 2]
    X = (kps[..., 0] - cx) / fx * Z
    Y = (kps[..., 1] - cy) / fy * Z
    cam = torch.stack((X, Y, Z), dim=-1) - trans_world.unsqueeze(1)
    return torch.einsum("bij,bkj->bki", torch.linalg.inv(rot_world), cam)


def world_to_image(kps, fx, fy, cx, cy, trans_world, rot_world):
    """Extrinsics then perspective projection (util.py:116-125)."""
    cam = torch.einsum("bij,bkj->bki", rot_world, kps) + trans_world.unsqueeze(1)
    u = cam[..., 0] / cam[..., 2] * fx + cx
    v = cam[..., 1] / cam[..., 2] * fy + cy
    return torch.stack((u, v, cam[..., 2]), dim=-1)


def image_to_patch(kps, trans_image, img_d, img_h, img_w, depth_scale, pelvis, is_norm=True):
    """Image px / depth mm -> normalised patch coords (util.py:98-113)."""
    z = (kps[..., 2] - pelvis[..., 2].unsqueeze(1)) / depth_scale
    A, t = _affine_parts(trans_image)
    uv = torch.einsum("bij,bkj->bki", A, kps[..., :2]) + t.unsqueeze(1)
    x, y = uv[..., 0], uv[..., 1]
    if is_norm:
        x = x / (img_w - 1) * 2 - 1
        y = y / (img_h - 1) * 2 - 1
        z = z / (img_d - 1)
    return torch.stack((x, y, z), dim=-1)


def _intrinsics(k_mat):
    return k_mat[..., 0, [0]], k_mat[..., 1, [1]], k_mat[..., 0, [2]], k_mat[..., 1, [2]]


def patch_to_world(kps, cams: Dict[str, torch.Tensor], img_hw: Sequence[int] = (256, 256), is_norm=True,
                   rect_width=2000.0, mono=False, patch=True):
    """convert_patch_to_world (util.py:128-152).  `cams` is keyed by
    trans_image/pelvis/k_mat/trans_world/rot_world; `img_hw` is the image's
    (H, W) — the reference reads it off `params['{mode}_img'].shape` and uses
    the *width* as the depth extent (:137)."""
    img_h, img_w = img_hw
    if patch:
        k_img = patch_to_image(kps, cams["trans_image"], img_w, img_h, img_w, 1.0 / img_w * rect_width,
                               cams["pelvis"], is_norm=is_norm)
    else:
        k_img = kps
    if mono:
        out = k_img.clone()
        out[..., 2] = out[..., 2] + 128
        return -out[..., [0, 2, 1]]
    fx, fy, cx, cy = _intrinsics(cams["k_mat"])
    return image_to_world(k_img, fx, fy, cx, cy, cams["trans_world"], cams["rot_world"])


def world_to_patch(kps, cams: Dict[str, torch.Tensor], img_hw: Sequence[int] = (256, 256), is_norm=True,
                   rect_width=2000.0):
    """convert_world_to_patch (util.py:155-168), the forward perspective projection."""
    img_h, img_w = img_hw
    fx, fy, cx, cy = _intrinsics(cams["k_mat"])
    k_img = world_to_image(kps, fx, fy, cx, cy, cams["trans_world"], cams["rot_world"])
    return image_to_patch(k_img, cams["trans_image"], img_w, img_h, img_w, 1.0 / img_w * rect_width,
                          cams["pelvis"], is_norm=is_norm)


# --------------------------------------------------------------------------- losses
def supervision_mse(pred, gt):
    """compute_supervision with the default arguments (loss_func.py:38-52): mean squared error."""
    return ((pred - gt) ** 2).mean()


def supervision(keypoint, keypoint_gt, feature_shape=None, mode="mean"):
    """compute_supervision with all its arguments (loss_func.py:38-52): optional rescale of the prediction to a feature
    grid (:39-45), then nn.MSELoss(reduction=mode) with mode in mean / sum (divided by the batch size, :50-51) / none."""
    k = keypoint
    if feature_shape is not None:
        cols = [(k[..., 0] + 1) / 2.0 * (feature_shape[0] - 1), (k[..., 1] + 1) / 2.0 * (feature_shape[1] - 1)]
        if k.shape[-1] == 3:
            cols.append(k[..., 2] * (feature_shape[2] - 1))
        k = torch.stack(cols, dim=-1)
    e = (k - keypoint_gt) ** 2
    if mode == "none":
        return e
    if mode == "sum":
        return e.sum() / k.shape[0]
    if mode == "mean":
        return e.mean()
    raise ValueError("unknown mode %r" % (mode,))


def bone_sym(world):
    """compute_bone_sym_loss (loss_func.py:18-25) on `[B,K,3]` world mm."""
    v = world[:, list(BONE_CHILD), :] - world[:, list(BONE_PARENT), :]
    n = v.pow(2).sum(-1).sqrt() * 1e-3
    return ((n[:, 0::2] - n[:, 1::2]) ** 2).mean()


def kp_sym(k, is_3d=True):
    """compute_kp_sym_loss (loss_func.py:27-35): midpoints of (11,14) and (1,4) against
    joints K-1 and 0."""
    mid = (k[:, list(MID_A), :] + k[:, list(MID_B), :]) / 2
    ref = k[:, [-1, 0], :]
    if is_3d:
        return ((mid * 1e-3 - ref * 1e-3) ** 2).mean()
    return ((mid - ref) ** 2).mean()


def per_sample_terms(kps, target, world, w_bone, w_kp, w_kp2d):
    """Per-(b,h) un-normalised sums behind the four loss terms.
    Returns mse_sum[B,NH] (over k,c), bone_sum[B,NH] (over 4 pairs),
    kp_sum[B,NH] (over 2x3), kp2d_sum[B,NH] (over 2x2)."""
    B, NH, K, _ = kps.shape
    mse = ((kps - target.unsqueeze(1)) ** 2).sum(dim=(2, 3))
    v = world[:, :, list(BONE_CHILD), :] - world[:, :, list(BONE_PARENT), :]
    n = v.pow(2).sum(-1).sqrt() * 1e-3
    bone = ((n[..., 0::2] - n[..., 1::2]) ** 2).sum(-1)
    mid = (world[:, :, list(MID_A), :] + world[:, :, list(MID_B), :]) / 2
    ref = world[:, :, [-1, 0], :]
    kp3 = ((mid * 1e-3 - ref * 1e-3) ** 2).sum(dim=(2, 3))
    k2 = kps[..., :2]
    mid2 = (k2[:, :, list(MID_A), :] + k2[:, :, list(MID_B), :]) / 2
    kp2 = ((mid2 - k2[:, :, [-1, 0], :]) ** 2).sum(dim=(2, 3))
    return mse, bone, kp3, kp2


def reproj_min_loss(kps, target, cams, img_hw=(256, 256), rect_width=2000.0, w_mse=1.0,
                    w_bone=None, w_kp=None, w_kp2d=None, reduction="batch", batch_total=None):
    """The per-camera loss graph around the head (model.py:71-79,105-114,158-162).

    kps `[B,NH,K,3]` (differentiable), target `[B,K,3]`.
    reduction:
      'batch'  - model.py semantics: one scalar per hypothesis over the batch, `torch.min`
                 picks the slot; pseudo-MSE and symmetry pick independently.
      'sample' - per-sample min over hypotheses, then batch mean (loss_func.py:59 style).
      'joint'  - per-(sample, joint) argmin of the squared error (eval.py:138-145); MSE only.
    Returns (loss_pseudo, loss_sym, sel, world) where sel is
      batch: int64[2] (pseudo slot, symmetry slot; -1 when the term is off)
      sample: int64[2,B]     joint: int64[B,K].
    `batch_total` overrides B in the mean denominators (global-batch scope)."""
    B, NH, K, _ = kps.shape
    world = torch.stack([patch_to_world(kps[:, h], cams, img_hw, True, rect_width) for h in range(NH)], dim=1)
    use_sym = any(w is not None for w in (w_bone, w_kp, w_kp2d))
    wb = 0.0 if w_bone is None else w_bone
    wk = 0.0 if w_kp is None else w_kp
    wk2 = 0.0 if w_kp2d is None else w_kp2d
    mse, bone, kp3, kp2 = per_sample_terms(kps, target, world, wb, wk, wk2)
    n = float(B if batch_total is None else batch_total)
    zero = kps.new_zeros(())
    if reduction == "batch":
        mse_h = mse.sum(0) / (n * K * 3)
        sel_m = int(torch.argmin(mse_h))
        loss_p = w_mse * mse_h[sel_m]
        sel_s, loss_s = -1, zero
        if use_sym:
            sym_h = wb * bone.sum(0) / (n * 4) + wk * kp3.sum(0) / (n * 6) + wk2 * 1e2 * kp2.sum(0) / (n * 4)
            sel_s = int(torch.argmin(sym_h))
            loss_s = sym_h[sel_s]
        return loss_p, loss_s, torch.tensor([sel_m, sel_s]), world
    if reduction == "sample":
        mse_bh = w_mse * mse / (K * 3)
        vm, im = mse_bh.min(dim=1)
        loss_p = vm.sum() / n
        loss_s, isym = zero, torch.full((B,), -1, dtype=torch.long, device=kps.device)
        if use_sym:
            sym_bh = wb * bone / 4 + wk * kp3 / 6 + wk2 * 1e2 * kp2 / 4
            vs, isym = sym_bh.min(dim=1)
            loss_s = vs.sum() / n
        return loss_p, loss_s, torch.stack((im, isym)), world
    if reduction == "joint":
        e = ((kps - target.unsqueeze(1)) ** 2).sum(-1)               # [B,NH,K]
        v, i = e.min(dim=1)
        return w_mse * v.sum() / (n * K * 3), zero, i, world
    raise ValueError("unknown reduction %r" % (reduction,))


def fused_forward(logits, num_kp, num_hypo, neighbor_size, target, cams, **kw):
    """Head tail + per-camera loss graph in one call (what IntegralReprojMinLoss computes)."""
    kps, dmap, idx = integral_multi(logits, num_kp, num_hypo, neighbor_size)
    loss_p, loss_s, sel, world = reproj_min_loss(kps, target, cams, **kw)
    return loss_p, loss_s, sel, kps, world, dmap, idx


def best_hypothesis(kps, gt):
    """eval.py:138-145: per-(b,k) argmin over hypotheses of the squared error, and the gather."""
    idx = (kps - gt[:, None]).pow(2).sum(-1).argmin(dim=1)             # [B,K]
    best = torch.gather(kps, 1, idx[:, None, :, None].expand(-1, -1, -1, kps.shape[-1])).squeeze(1)
    return idx, best


# --------------------------------------------------------------------------- skeleton rasteriser + mask loss
ARM_LINES = (11, 12, 14, 15)      # util.py:53: lines drawn with half the body width when there are >= 21 of them


def skeleton_links(parent_ids: Sequence[int], line_select_ids: Optional[Sequence[int]] = None,
                   use_root: bool = False, extension: bool = True):
    """cal_links (model.py:8-22): (parent, child) joint pairs of the drawn lines; the eight
    extension links are the torso/limb cross braces appended at model.py:19-20."""
    parent_ids = list(parent_ids)
    if use_root:
        child = list(range(len(parent_ids)))
        parent = parent_ids
    else:
        child = list(range(1, len(parent_ids)))
        parent = parent_ids[1:]
    if line_select_ids is None:
        line_select_ids = range(len(parent))
    parent = [parent[i] for i in line_select_ids]
    child = [child[i] for i in line_select_ids]
    if extension:
        parent += [7, 7, 7, 7, 0, 0, 1, 4]
        child += [1, 4, 11, 14, 2, 5, 14, 11]
    return parent, child


def pixel_grid(size: int, dtype) -> torch.Tensor:
    """make_coordinate_grid (util.py:3-19) flattened to `[size*size, 2]`, x fastest."""
    c = 2 * (torch.arange(size).to(dtype) / (size - 1)) - 1
    return torch.stack((c.repeat(size), c.repeat_interleave(size)), dim=-1)


def segment_sqdist(kp2d: torch.Tensor, size: int, parent_ids, child_ids) -> torch.Tensor:
    """Squared distance of every pixel centre to every line segment, `[B, L, size*size]`
    (util.py:34-47).  The segment runs from the child joint (`start`) to the parent joint (`end`)."""
    s = kp2d[:, list(child_ids), :]                       # [B,L,2]
    e = kp2d[:, list(parent_ids), :]
    d = e - s
    g = pixel_grid(size, kp2d.dtype).to(kp2d.device)      # [P,2]
    a = g[None, None] - s[:, :, None]                     # [B,L,P,2]
    t = (a * d[:, :, None]).sum(-1) / (1e-8 + (d * d).sum(-1, keepdim=True))
    to_end = g[None, None] - e[:, :, None]
    foot = g[None, None] - (s[:, :, None] + t[..., None] * d[:, :, None])
    zero = torch.zeros((), dtype=kp2d.dtype, device=kp2d.device)
    q = torch.where(t <= 0, (a * a).sum(-1), zero) + torch.where(t >= 1, (to_end * to_end).sum(-1), zero) \
        + torch.where((t > 0) & (t < 1), (foot * foot).sum(-1), zero)
    return q


def draw_lines(kp2d: torch.Tensor, size: int, parent_ids, child_ids, body_width: float) -> torch.Tensor:
    """draw_lines (util.py:21-59): Gaussian-profile line heat-maps `[B, L, size, size]`."""
    B = kp2d.shape[0]
    u = -segment_sqdist(kp2d, size, parent_ids, child_ids) / body_width
    if u.shape[1] >= 21:
        scale = torch.ones(u.shape[1], dtype=u.dtype, device=u.device)
        scale[list(ARM_LINES)] = 2
        u = u * scale[None, :, None]
    return torch.exp(u).reshape(B, -1, size, size)


def skeleton_mask(kp2d, size, parent_ids, child_ids, body_width):
    """model.py:91-94: channel-max of the line heat-maps, `[B, 1, size, size]`."""
    return draw_lines(kp2d, size, parent_ids, child_ids, body_width).max(dim=1, keepdim=True)[0]


def mask_recon_loss(mask, gt, weight=None, use_clip=False):
    """compute_mask_reconstruction_loss (loss_func.py:4-16).  Note the reference's asymmetry: without a
    weight map the MSE is reduced to a scalar *before* the clip filter multiplies it, so with
    `use_clip` and no weight the result is a tensor shaped like `mask` (the trainer `.mean()`s it,
    train.py:182)."""
    sq = (mask - gt) ** 2
    loss = sq.mean() if weight is None else sq
    if use_clip:
        loss = loss * (mask > 0.1).to(mask.dtype)
    if weight is not None:
        loss = (loss * weight).mean()
    return loss


# --------------------------------------------------------------------------- eval-side selection + triangulation
SWITCH_LIST = ((1, 4), (2, 5), (3, 6), (14, 11), (15, 12), (16, 13))     # left/right pairs, eval_utils.py:8


def switch_points(points, gt, switch_all=False, switch_list=SWITCH_LIST):
    """eval_utils.py:7-29: left/right-swapped copy of `points [B,K,C]`, kept wherever its L1 error in
    (x, y) against `gt` is strictly smaller - per joint, or per sample with `switch_all`."""
    perm = list(range(points.shape[1]))
    for a, b in switch_list:
        perm[a], perm[b] = b, a
    swapped = points[:, perm, :]
    e_sw = (swapped - gt).abs()[..., :2]
    e = (points - gt).abs()[..., :2]
    dims = (1, 2) if switch_all else (2,)
    is_trans = e_sw.sum(dim=dims, keepdim=True) < e.sum(dim=dims, keepdim=True)
    return torch.where(is_trans, swapped, points), is_trans


def per_act_mse(pred, gt):
    """eval_utils.py:31-41: mean over joints of the 2-D Euclidean error in [0,1] patch units -> `[B]`."""
    return ((((pred + 1) / 2 - (gt + 1) / 2) ** 2).sum(dim=2)).sqrt().mean(dim=1)


def eval_select(kps, joints_px, img_size=256.0, mode="best"):
    """eval.py:117-148 for one camera: normalise the pixel-space ground truth, undo left/right swaps per
    hypothesis (2-D copy and 3-D prediction separately), then per joint keep the hypothesis closest to the
    ground truth ('best') or hypothesis 0 ('confident').
    Returns (kp3d [B,K,3], kp2d [B,K,2], is_trans [B,K,1] of the LAST hypothesis (the reference overwrites
    `trans_dict[cam_key]` in its loop), err2d [B], best_idx [B,K], best_2d_idx [B,K], gt_norm [B,K,3])."""
    B, NH, K, _ = kps.shape
    gt = joints_px.clone()
    gt[..., :2] = gt[..., :2] / (img_size - 1) * 2 - 1
    gt[..., 2] = gt[..., 2] / (img_size - 1)
    k3 = kps.clone()
    k2 = kps[..., :2].clone()
    is_trans = None
    for h in range(NH):
        k2[:, h], _ = switch_points(k2[:, h], gt[..., :2])
        k3[:, h], is_trans = switch_points(k3[:, h], gt, switch_all=False)
    if mode == "best" and NH > 1:
        bi = (k3 - gt[:, None]).pow(2).sum(-1).argmin(dim=1)
        k3 = torch.gather(k3, 1, bi[:, None, :, None].expand(-1, -1, -1, 3)).squeeze(1)
        b2 = (k2 - gt[:, None, :, :2]).pow(2).sum(-1).argmin(dim=1)
        k2 = torch.gather(k2, 1, b2[:, None, :, None].expand(-1, -1, -1, 2)).squeeze(1)
    else:
        bi = torch.zeros(B, K, dtype=torch.long, device=kps.device)
        b2 = bi
        k3, k2 = k3[:, 0], k2[:, 0]
    return k3, k2, is_trans, per_act_mse(k2, gt[..., :2]), bi, b2, gt


def triangulate(kps_by_cam, cams_by_cam, img_hw=(256, 256), is_norm=True, rect_width=2000.0):
    """triangulation + batch_triangulate (util.py:171-230): every camera's patch keypoints `[B,K,3]` to image
    pixels (+ metric depth, which the reference then uses as the per-point confidence weight), P = K [R|T],
    the two DLT rows per view, and the right singular vector of the smallest singular value, dehomogenised."""
    img_h, img_w = img_hw
    pts, Ps = [], []
    for kps, cams in zip(kps_by_cam, cams_by_cam):
        pts.append(patch_to_image(kps, cams["trans_image"], img_w, img_h, img_w, 1.0 / img_w * rect_width,
                                  cams["pelvis"], is_norm=is_norm))
        Ps.append(cams["k_mat"] @ torch.cat((cams["rot_world"], cams["trans_world"].unsqueeze(-1)), dim=-1))
    pts = torch.stack(pts, dim=1)                                    # [B,V,K,3]
    P = torch.stack(Ps, dim=1)                                       # [B,V,3,4]
    u = pts[..., 0].permute(0, 2, 1)[..., None]                      # [B,K,V,1]
    v = pts[..., 1].permute(0, 2, 1)[..., None]
    conf = pts[..., 2].permute(0, 2, 1)[..., None]
    P0, P1, P2 = (P[:, None, :, i, :] for i in range(3))             # [B,1,V,4]
    A = torch.cat((conf * (u * P2 - P0), conf * (v * P2 - P1)), dim=2)   # [B,K,2V,4]
    X = torch.linalg.svd(A)[2][:, :, -1, :]
    return (X / X[..., 3:])[..., :3]


# --------------------------------------------------------------------------- discriminator-side glue
def root_centre(world, dim=3):
    """model.py:123-124, literally: `(w - w[:, [0], :]) / 1000`, then the first `dim` coordinates.  On the stacked
    `[B, NH, K, 3]` world joints the reference passes, `[:, [0], :]` indexes the HYPOTHESIS axis: every hypothesis is
    expressed relative to hypothesis 0 (which becomes all zeros), not relative to the root joint; on a `[B, K, 3]`
    tensor the same expression is the usual root-joint centring."""
    return ((world - world[:, [0]]) / 1000)[..., :dim]


def disc_loss(pred_logits, gt_logits=None):
    """compute_disc_loss (loss_func.py:54-76): least-squares GAN terms; a `[B,NH,1]` input takes the
    per-sample min over hypotheses before the batch mean."""
    def term(x, target):
        e = (x - target).pow(2)
        if x.dim() == 3:
            e = e.min(dim=1)[0]
        elif x.dim() != 2:
            raise ValueError("Invalid dimension of logits")
        return e.mean()
    if gt_logits is None:
        return term(pred_logits, 1.0)
    return 0.5 * term(gt_logits, 1.0) + 0.5 * term(pred_logits, 0.0)
