"""Deterministic synthetic inputs for the integral + reprojection-loss path.

The shapes and value ranges follow the input contract the reference's data
loader defines: logits as `[B, K*D, H, W]`
(`modules/keypoint_detector_integral_multi.py:67-70`), pseudo joints with
x,y in [-1,1] and z in metres/2 (`human_utils/dataloader/dataloader.py:226-227`),
H36M-like cameras (`human_utils/dataset/hm36.py:163-186,284-304`).

Everything is generated on the CPU from a seeded `torch.Generator` in fp32 so
that the CPU oracle and the GPU kernels receive identical bits.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

__all__ = ["iid_logits", "blob_logits", "pseudo_joints", "cameras", "camera_dict", "CAM_FIELDS",
           "H36M_PARENTS", "LINE_SELECT", "BODY_WIDTH", "skeleton_pose2d", "silhouette_mask", "geodesic_weight",
           "eval_predictions", "SWITCH_LIST", "model_batch", "model_cfg"]

# order of the per-sample camera tensors everywhere in this package
CAM_FIELDS = ("trans_image", "pelvis", "k_mat", "trans_world", "rot_world")


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def iid_logits(B: int, K: int, D: int, H: int, W: int, seed: int = 0, scale: float = 1.0) -> torch.Tensor:
    """iid N(0, scale^2) logits, `[B, K*D, H, W]` fp32 (stress case: many depth peaks)."""
    return torch.randn(B, K * D, H, W, generator=_gen(seed), dtype=torch.float32) * scale


def blob_logits(B: int, K: int, D: int, H: int, W: int, seed: int = 1, modes: int = 3) -> torch.Tensor:
    """Multi-modal depth "blob" logits: `modes` Gaussian depth
    modes at one shared (h, w) centre per (b, k) on top of N(0,1) noise."""
    g = _gen(seed)
    L = torch.randn(B, K, D, H, W, generator=g, dtype=torch.float32)
    cx = (0.2 + 0.6 * torch.rand(B, K, generator=g)) * W
    cy = (0.2 + 0.6 * torch.rand(B, K, generator=g)) * H
    hh = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1)
    ww = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    gxy = -((hh - cy.view(B, K, 1, 1)) ** 2 + (ww - cx.view(B, K, 1, 1)) ** 2) / (2 * 2.5 ** 2)
    dd = torch.arange(D, dtype=torch.float32).view(1, 1, D)
    for _ in range(modes):
        cz = (0.1 + 0.8 * torch.rand(B, K, generator=g)) * D
        amp = 12.0 * (0.5 + 0.5 * torch.rand(B, K, generator=g))
        gz = -((dd - cz.view(B, K, 1)) ** 2) / (2 * 2.0 ** 2)
        L += amp.view(B, K, 1, 1, 1) * torch.exp(gz.view(B, K, D, 1, 1) + gxy.view(B, K, 1, H, W))
    return L.view(B, K * D, H, W).contiguous()


def pseudo_joints(B: int, K: int, seed: int = 2) -> torch.Tensor:
    """Pseudo-GT joints `[B, K, 3]`: x,y ~ U(-1,1), z ~ U(-0.5,0.5)."""
    g = _gen(seed)
    xy = torch.rand(B, K, 2, generator=g) * 2 - 1
    z = torch.rand(B, K, 1, generator=g) - 0.5
    return torch.cat((xy, z), dim=-1).contiguous()


def cameras(B: int, seed: int = 3, mpi: bool = False, img: int = 256, rect_width: float = 2000.0) -> Dict[str, torch.Tensor]:
    """Per-sample camera tensors keyed by `CAM_FIELDS` (H36M-like; `mpi=True`
    uses MPI-INF-3DHP intrinsics, `human_utils/dataset/mpi_inf_3dhp.py:53,176-187`)."""
    g = _gen(seed)
    if mpi:
        f = 1497.7 + torch.randn(B, 1, generator=g) * 0.5
        fx, fy = f, f.clone()
        cx = 1024.0 + torch.randn(B, 1, generator=g)
        cy = 1024.0 + torch.randn(B, 1, generator=g)
    else:
        fx = 1140.0 + 10.0 * torch.rand(B, 1, generator=g)
        fy = 1140.0 + 10.0 * torch.rand(B, 1, generator=g)
        cx = 500.0 + 20.0 * torch.rand(B, 1, generator=g)
        cy = 500.0 + 20.0 * torch.rand(B, 1, generator=g)
    k_mat = torch.zeros(B, 3, 3)
    k_mat[:, 0, 0] = fx[:, 0]
    k_mat[:, 1, 1] = fy[:, 0]
    k_mat[:, 0, 2] = cx[:, 0]
    k_mat[:, 1, 2] = cy[:, 0]
    k_mat[:, 2, 2] = 1.0

    pelvis = torch.empty(B, 3)
    pelvis[:, :2] = torch.randn(B, 2, generator=g) * 300.0
    pelvis[:, 2] = 3000.0 + 3000.0 * torch.rand(B, generator=g)

    # crop affine: scale so that rect_width mm at the pelvis depth spans the patch,
    # translation puts the pelvis projection at the patch centre
    # patch-px per image-px: rect_width mm at depth Z span rect_width*fx/Z image px -> img patch px
    s = img * pelvis[:, 2] / (rect_width * fx[:, 0])  # ~0.3-0.7 for H36M
    u0 = pelvis[:, 0] / pelvis[:, 2] * fx[:, 0] + cx[:, 0]
    v0 = pelvis[:, 1] / pelvis[:, 2] * fy[:, 0] + cy[:, 0]
    trans_image = torch.zeros(B, 2, 3)
    trans_image[:, 0, 0] = s
    trans_image[:, 1, 1] = s
    trans_image[:, :, :2] += torch.randn(B, 2, 2, generator=g) * 1e-3
    trans_image[:, 0, 2] = img / 2.0 - s * u0
    trans_image[:, 1, 2] = img / 2.0 - s * v0

    q, r = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))
    q = q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1)).unsqueeze(-2)
    det = torch.linalg.det(q)
    q[:, :, 2] *= det.sign().unsqueeze(-1)
    rot_world = q.contiguous()
    trans_world = torch.randn(B, 3, generator=g) * 2000.0

    return {
        "trans_image": trans_image.contiguous(),
        "pelvis": pelvis.contiguous(),
        "k_mat": k_mat.contiguous(),
        "trans_world": trans_world.contiguous(),
        "rot_world": rot_world,
    }


def camera_dict(cams: Dict[str, torch.Tensor], mode: str = "cam_0", img: int = 256) -> Dict[str, torch.Tensor]:
    """The reference's dict-keyed parameter bundle (`modules/util.py:129-134`):
    `'{mode}_trans_image'`, `'{mode}_img'` (only `.shape` is read), ..."""
    B = cams["pelvis"].shape[0]
    out = {"{}_{}".format(mode, k): v for k, v in cams.items()}
    # a zero-stride view: only the shape is ever read (util.py:130,137-138)
    out["{}_img".format(mode)] = torch.zeros(1, dtype=cams["pelvis"].dtype, device=cams["pelvis"].device).expand(B, 3, img, img)
    return out


# --------------------------------------------------------------------------------------- skeleton rasteriser inputs
# config/HM36_Multi_SurS1.yaml:56,60,61: 18-joint H36M tree (joint 17 = thorax), the 17 drawn child->parent
# links and the body width (x 1e-3 at model.py:32)
H36M_PARENTS = (0, 0, 1, 2, 0, 4, 5, 0, 17, 8, 9, 17, 11, 12, 17, 14, 15, 7)
LINE_SELECT = tuple(range(17))
BODY_WIDTH = 3.0e-3

# a standing figure in patch coordinates (x right, y down, [-1,1]); joint order as H36M_PARENTS
_TEMPLATE = ((0.0, 0.05), (-0.12, 0.05), (-0.13, 0.40), (-0.14, 0.75), (0.12, 0.05), (0.13, 0.40), (0.14, 0.75),
             (0.0, -0.20), (0.0, -0.50), (0.0, -0.60), (0.0, -0.72), (0.20, -0.45), (0.30, -0.20), (0.33, 0.05),
             (-0.20, -0.45), (-0.30, -0.20), (-0.33, 0.05), (0.0, -0.45))


def skeleton_pose2d(B: int, K: int = 18, seed: int = 20, jitter: float = 0.06) -> torch.Tensor:
    """2-D patch poses `[B, K, 2]`: the template figure with per-joint jitter, a global scale U(0.7,1.1)
    and a shift N(0, 0.08).  Joints beyond the 18 of the template are placed uniformly in [-0.8, 0.8]."""
    g = _gen(seed)
    base = torch.tensor(_TEMPLATE, dtype=torch.float32)
    if K <= base.shape[0]:
        base = base[:K].expand(B, K, 2)
    else:
        extra = torch.rand(K - base.shape[0], 2, generator=g) * 1.6 - 0.8
        base = torch.cat((base, extra)).expand(B, K, 2)
    pose = base + jitter * torch.randn(B, K, 2, generator=g)
    scale = 0.7 + 0.4 * torch.rand(B, 1, 1, generator=g)
    shift = 0.08 * torch.randn(B, 1, 2, generator=g)
    return (pose * scale + shift).contiguous()


def _pixel_centres(size: int) -> torch.Tensor:
    c = 2 * (torch.arange(size, dtype=torch.float32) / (size - 1)) - 1
    return torch.stack((c.view(1, size).expand(size, size), c.view(size, 1).expand(size, size)), dim=-1)   # [S,S,2] (x,y)


def silhouette_mask(pose: torch.Tensor, size: int = 256, radius: float = 0.075) -> torch.Tensor:
    """Binary person mask `[B,1,S,S]` (floats 0/1, like `mask_patch / 255`, dataloader.py:184): the union of
    discs around the joints and the points at 1/3 and 2/3 of every tree edge of `pose`."""
    B, K, _ = pose.shape
    par = torch.tensor([H36M_PARENTS[k] if k < len(H36M_PARENTS) else 0 for k in range(K)])
    pts = torch.cat((pose, pose + (pose[:, par] - pose) / 3, pose + (pose[:, par] - pose) * 2 / 3), dim=1)   # [B,3K,2]
    g = _pixel_centres(size).view(1, 1, size * size, 2)
    out = torch.zeros(B, size * size, dtype=torch.bool)
    for i in range(0, pts.shape[1], 6):                                       # bounded temporaries
        d2 = (g - pts[:, i:i + 6, None, :]).pow(2).sum(-1)
        out |= (d2 < radius * radius).any(dim=1)
    return out.view(B, 1, size, size).to(torch.float32)


def geodesic_weight(mask: torch.Tensor, seed: int = 21) -> torch.Tensor:
    """A positive per-pixel weight map `[B,1,S,S]` standing in for `geodesic_dis` (geodesic.py:14-53:
    exp of a normalised in-mask distance plus a scaled background distance): larger inside the mask and
    growing away from its centroid, plus a background ramp."""
    B, _, S, _ = mask.shape
    g = _pixel_centres(S)
    gen = _gen(seed)
    a = 0.5 + torch.rand(B, 1, 1, 1, generator=gen)
    m = mask[:, 0]
    cnt = m.sum(dim=(1, 2)).clamp_min(1.0)
    cx = (m * g[..., 0]).sum(dim=(1, 2)) / cnt
    cy = (m * g[..., 1]).sum(dim=(1, 2)) / cnt
    r = ((g[None, ..., 0] - cx.view(B, 1, 1)) ** 2 + (g[None, ..., 1] - cy.view(B, 1, 1)) ** 2).sqrt()
    r = r / r.amax(dim=(1, 2), keepdim=True)
    w = mask * torch.exp(a * r[:, None]) + 0.3 + 0.7 * r[:, None]
    return w.contiguous()


# --------------------------------------------------------------------------------------- eval-side inputs
SWITCH_LIST = ((1, 4), (2, 5), (3, 6), (14, 11), (15, 12), (16, 13))      # eval_utils.py:8


def eval_predictions(B: int, NH: int, K: int, seed: int = 80, img_size: float = 256.0, noise: float = 0.05):
    """(kps [B,NH,K,3], joints_px [B,K,3]): pixel-space ground-truth joints (`{cam}_joints`, dataloader.py:166) and
    multi-hypothesis predictions scattered around their normalised positions; every other sample predicts the
    left/right-swapped skeleton, the case `switch_points` exists for."""
    g = _gen(seed)
    jp = torch.rand(B, K, 3, generator=g) * (img_size - 1)
    gt = jp.clone()
    gt[..., :2] = gt[..., :2] / (img_size - 1) * 2 - 1
    gt[..., 2] = gt[..., 2] / (img_size - 1)
    kps = gt[:, None] + noise * (torch.rand(B, NH, K, 3, generator=g) * 2 - 1)
    perm = list(range(K))
    for a, b in SWITCH_LIST:
        if a < K and b < K:
            perm[a], perm[b] = b, a
    kps[1::2] = kps[1::2][:, :, perm]
    return kps.contiguous(), jp.contiguous()


# --------------------------------------------------------------------------------------- whole loss-graph inputs
def model_cfg(cam_ids=(0, 1), sym=(0.1, 0.1, 0.5), w_pseudo=1.0, w_gen=0.5, w_rec=0.02, use_dis_map=True, w_disc=0.5):
    """`model_params` of config/HM36_Multi_*.yaml reduced to what Counter3DModel / Counter3DDisc read (model.py:24-48,195-216)."""
    loss = {"smpl_pseudo_img_loss": {"weight": w_pseudo}, "smpl_gen_loss": {"weight": w_gen},
            "recons_loss": {"use_dis_map": use_dis_map, "weight": w_rec}, "smpl_disc_loss": {"weight": w_disc, "update_interval": 1}}
    if sym is not None:
        loss["symmetry_loss"] = {"weight": {"bone": sym[0], "kp": sym[1], "kp_2d": sym[2]}}
    return {"cam_id_list": list(cam_ids), "parent_ids": list(H36M_PARENTS), "line_select_ids": list(LINE_SELECT), "body_width": 3.0,
            "loss_config": loss, "smpl_disc_params": {"disc_sup_dim": 3}}


def model_batch(B: int, K: int, R: int, cam_ids=(0, 1), seed: int = 90) -> Dict[str, torch.Tensor]:
    """The dict the data loader hands to Counter3DModel.forward (dataloader.py:160-230), with the detector's backbone
    replaced by the identity: `{cam}_img` / `{cam}_pseudo_img` ARE the `[B, K*R, R, R]` logits (so the "image" is RxR)."""
    x = {}
    for i, c in enumerate(cam_ids):
        key = "cam_%d" % c
        x[key + "_img"] = blob_logits(B, K, R, R, R, seed=seed + 10 * i)
        x[key + "_pseudo_img"] = blob_logits(B, K, R, R, R, seed=seed + 10 * i + 1)
        x[key + "_pseudo_joints"] = pseudo_joints(B, K, seed=seed + 10 * i + 2)
        x[key + "_joints"] = torch.rand(B, K, 3, generator=_gen(seed + 10 * i + 3)) * (R - 1)
        for k, v in cameras(B, seed=seed + 10 * i + 4, img=R).items():
            x[key + "_" + k] = v
        x[key + "_mask"] = silhouette_mask(skeleton_pose2d(B, K, seed=seed + 10 * i + 5, jitter=0.03), R)
        x[key + "_geodesic_dis"] = geodesic_weight(x[key + "_mask"], seed=seed + 10 * i + 6)
    return x
