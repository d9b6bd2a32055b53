"""Eval-side selection + triangulation and the discriminator-side glue over the C ABI
(SURVEY.md section 8f rows 3 and 4).  Drop-ins with the reference's signatures:

* `switch_points`        eval_utils.py:7-29 (per-joint mode only goes through the fused kernel; see below)
* `eval_select`          eval.py:122-148 for one camera in ONE launch
* `per_act_mse`          eval_utils.py:31-41 (returned by `eval_select`)
* `triangulation`        modules/util.py:171-198 (+ batch_triangulate :201-230), dict-keyed like the reference
* `root_centre`          modules/model.py:123-124
* `compute_disc_loss`    modules/base_losses/loss_func.py:54-76

All arithmetic is in `csrc/eval_disc.cu`; there is no CPU path.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import _cabi as cabi

__all__ = ["SWITCH_LIST", "eval_select", "switch_points", "triangulation", "triangulate", "root_centre",
           "compute_disc_loss", "RootCentre", "DiscMinLoss"]

SWITCH_LIST = ((1, 4), (2, 5), (3, 6), (14, 11), (15, 12), (16, 13))      # eval_utils.py:8


def _perm(K: int, switch_list) -> list:
    perm = list(range(K))
    for a, b in switch_list:
        perm[a], perm[b] = b, a
    return perm


def eval_select(kps: torch.Tensor, joints_px: torch.Tensor, img_size: float = 256.0, mode: str = "best",
                switch_list=SWITCH_LIST) -> Dict[str, torch.Tensor]:
    """kps `[B,NH,K,3]` (detector output) and the pixel-space ground truth `joints_px [B,K,3]` ->
    dict(kp3d [B,K,3], kp2d [B,K,2], is_trans [B,K,1] bool, err2d [B], best_idx [B,K], best_2d_idx [B,K], gt [B,K,3])
    exactly as eval.py:122-148 leaves `kp_pred_dict[cam]`, `kp_pred_2d`, `trans_dict[cam]`, `error_2d`, `kp_gt`."""
    cabi.require_cuda(kps, "kps")
    if mode not in ("best", "confident"):
        raise ValueError("Unknown mode: {}".format(mode))
    kps = kps.detach().to(torch.float32).contiguous()
    B, NH, K, _ = kps.shape
    dev = kps.device
    jp = joints_px.detach().to(device=dev, dtype=torch.float32).contiguous()
    if tuple(jp.shape) != (B, K, 3):
        raise ValueError("joints must be [B,K,3] = %s, got %s" % ((B, K, 3), tuple(jp.shape)))
    cfg = cabi.Eval(B, NH, K, float(img_size), int(mode == "best" and NH > 1))
    for k, v in enumerate(_perm(K, switch_list)):
        cfg.perm[k] = v
    out = {"kp3d": torch.empty(B, K, 3, device=dev), "kp2d": torch.empty(B, K, 2, device=dev),
           "is_trans": torch.empty(B, K, 1, dtype=torch.uint8, device=dev), "err2d": torch.empty(B, device=dev),
           "best_idx": torch.empty(B, K, dtype=torch.int64, device=dev),
           "best_2d_idx": torch.empty(B, K, dtype=torch.int64, device=dev), "gt": torch.empty(B, K, 3, device=dev)}
    with torch.cuda.device(dev):
        cabi.check(cabi.lib.xsup_eval_select(kps.data_ptr(), jp.data_ptr(), cfg, out["kp3d"].data_ptr(), out["kp2d"].data_ptr(),
                                             out["is_trans"].data_ptr(), out["err2d"].data_ptr(), out["best_idx"].data_ptr(),
                                             out["best_2d_idx"].data_ptr(), out["gt"].data_ptr(), cabi.stream_ptr(dev)),
                   "xsup_eval_select")
    out["is_trans"] = out["is_trans"].bool()
    return out


def switch_points(points, gt, switch_all=False, switch_list=SWITCH_LIST):
    """eval_utils.py:7 for the per-joint mode the eval loop uses (`switch_all=False`): the fused kernel run on one
    hypothesis with an already-normalised ground truth (`img_size=0`), so the swap decisions are bit-exact.  Returns (points with swaps undone, is_trans [B,K,1])."""
    if switch_all:
        raise NotImplementedError("switch_all=True is not on the eval path (eval.py:135-136 pass False)")
    C = points.shape[-1]
    p3 = points if C == 3 else torch.cat((points, points.new_zeros(points.shape[:-1] + (3 - C,))), dim=-1)
    g3 = gt if gt.shape[-1] == 3 else torch.cat((gt, gt.new_zeros(gt.shape[:-1] + (3 - gt.shape[-1],))), dim=-1)
    # img_size = 0 tells the kernel the ground truth is already normalised: its bits reach the strict `es < e` swap
    # decision untouched, as in the reference
    out = eval_select(p3.unsqueeze(1), g3.to(torch.float32), img_size=0.0, mode="confident", switch_list=switch_list)
    return out["kp3d"][..., :C], out["is_trans"]


def triangulate(kps_by_cam: Sequence[torch.Tensor], cams_by_cam: Sequence[Dict[str, torch.Tensor]], img_hw=(256, 256),
                is_norm: bool = True, rect_width: float = 2000.0) -> torch.Tensor:
    """V cameras' patch keypoints `[B,K,3]` + camera tensor dicts -> world `[B,K,3]` (util.py:171-230)."""
    V = len(kps_by_cam)
    if V != len(cams_by_cam) or not 2 <= V <= cabi.MAX_VIEWS:
        raise ValueError("need 2..%d cameras with one keypoint tensor each, got %d / %d" % (cabi.MAX_VIEWS, V, len(cams_by_cam)))
    cabi.require_cuda(kps_by_cam[0], "keypoints")
    dev = kps_by_cam[0].device
    B, K, _ = kps_by_cam[0].shape
    t = cabi.Tri(V, B, K, int(img_hw[0]), int(img_hw[1]), int(bool(is_norm)), float(rect_width))
    keep = []
    for v in range(V):
        kp = kps_by_cam[v].detach().to(torch.float32).contiguous()
        if tuple(kp.shape) != (B, K, 3):
            raise ValueError("camera %d keypoints are %s, expected %s" % (v, tuple(kp.shape), (B, K, 3)))
        cam_t = [cams_by_cam[v][k].to(torch.float32).contiguous() for k in ("trans_image", "pelvis", "k_mat", "trans_world", "rot_world")]
        keep.append((kp, cam_t))
        t.kps[v] = kp.data_ptr()
        t.cam[v] = cabi.make_cam(*cam_t, B)
    world = torch.empty(B, K, 3, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        cabi.check(cabi.lib.xsup_triangulate(t, world.data_ptr(), cabi.stream_ptr(dev)), "xsup_triangulate")
    return world


def triangulation(keypoints: Dict[str, torch.Tensor], params: Dict[str, torch.Tensor], cam_id_list, is_norm=True, RECT_WIDTH=2000):
    """Same dict-keyed signature as modules/util.py:171."""
    kps, cams, hw = [], [], None
    for cam_id in cam_id_list:
        mode = "cam_{}".format(cam_id)
        kps.append(keypoints[mode])
        cams.append({k: params["{}_{}".format(mode, k)] for k in ("trans_image", "pelvis", "k_mat", "trans_world", "rot_world")})
        shp = params["{}_img".format(mode)].shape
        hw = (int(shp[-2]), int(shp[-1]))
    return triangulate(kps, cams, hw, is_norm, float(RECT_WIDTH))


class RootCentre(torch.autograd.Function):
    """`(world - world[:, [0], :]) / 1000` sliced to `dim` coordinates, exactly as model.py:123-124 writes it."""

    @staticmethod
    def forward(ctx, world, dim):
        cabi.require_cuda(world, "world joints")
        w = world.detach().to(torch.float32).contiguous()
        if w.dim() not in (3, 4) or w.shape[-1] != 3:
            raise ValueError("world joints must be [B, K, 3] or [B, NH, K, 3], got %s" % (tuple(world.shape),))
        N, M = w.shape[0], w.shape[1]
        R = w.numel() // max(N * M, 1) if w.numel() else 3 * (w.shape[2] if w.dim() == 4 else 1)
        out = torch.empty(w.shape[:-1] + (dim,), dtype=torch.float32, device=w.device)
        with torch.cuda.device(w.device):
            cabi.check(cabi.lib.xsup_root_centre_fwd(w.data_ptr(), out.data_ptr(), N, M, R, dim, cabi.stream_ptr(w.device)),
                       "xsup_root_centre_fwd")
        ctx.meta = (N, M, R, dim, tuple(w.shape), world.dtype)
        return out

    @staticmethod
    def backward(ctx, g_out):
        N, M, R, dim, shape, in_dtype = ctx.meta
        g = g_out.to(torch.float32).contiguous()
        gw = torch.empty(shape, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            cabi.check(cabi.lib.xsup_root_centre_bwd(g.data_ptr(), gw.data_ptr(), N, M, R, dim, cabi.stream_ptr(g.device)),
                       "xsup_root_centre_bwd")
        return gw.to(in_dtype), None


def root_centre(world: torch.Tensor, dim: int = 3) -> torch.Tensor:
    """model.py:123-124 kept literally: item 0 along axis 1 is subtracted.  On the stacked `[B, NH, K, 3]` world
    joints the reference passes that axis is the HYPOTHESIS axis (hypothesis 0 becomes all zeros, the others are
    expressed relative to it); pass a `[B, K, 3]` tensor for the usual root-joint centring.  The result
    `out.flatten(0, 1)` feeds ONE batched discriminator call instead of the reference's NH calls (model.py:126-129)."""
    return RootCentre.apply(world, int(dim))


class DiscMinLoss(torch.autograd.Function):
    """logits `[B,NH,C]` -> mean over (b,c) of min over h of (x - target)^2, differentiable in the logits."""

    @staticmethod
    def forward(ctx, logits, target):
        cabi.require_cuda(logits, "logits")
        x = logits.detach().to(torch.float32).contiguous()
        B, NH, C = x.shape
        dev = x.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        sel = torch.empty(B, C, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            cabi.check(cabi.lib.xsup_disc_min_loss_fwd(x.data_ptr(), B, NH, C, float(target), loss.data_ptr(), sel.data_ptr(),
                                                       cabi.stream_ptr(dev)), "xsup_disc_min_loss_fwd")
        ctx.save_for_backward(x, sel)
        ctx.target, ctx.in_dtype = float(target), logits.dtype
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        x, sel = ctx.saved_tensors
        B, NH, C = x.shape
        g = torch.empty_like(x)
        gl = g_loss.to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(x.device):
            cabi.check(cabi.lib.xsup_disc_min_loss_bwd(x.data_ptr(), sel.data_ptr(), gl.data_ptr(), B, NH, C, ctx.target,
                                                       g.data_ptr(), cabi.stream_ptr(x.device)), "xsup_disc_min_loss_bwd")
        return g.to(ctx.in_dtype), None


def _term(x: torch.Tensor, target: float) -> torch.Tensor:
    if x.dim() == 2:
        return DiscMinLoss.apply(x.unsqueeze(1), target)
    if x.dim() == 3:
        return DiscMinLoss.apply(x, target)
    raise ValueError("Invalid dimension of logits")


def compute_disc_loss(pred_logits, gt_logits):
    """Same signature as modules/base_losses/loss_func.py:54."""
    if gt_logits is None:
        return _term(pred_logits, 1.0)
    return 0.5 * _term(gt_logits, 1.0) + 0.5 * _term(pred_logits, 0.0)
