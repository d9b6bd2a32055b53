"""Skeleton rasteriser + mask-reconstruction loss over the C ABI.

Drop-ins with the reference's signatures:

* `cal_links`                         modules/model.py:8-22 (host logic)
* `draw_lines`                        modules/util.py:21-59            -> `[B, L, S, S]`
* `skeleton_mask`                     modules/model.py:91-94           -> `[B, 1, S, S]` (max over lines, never
                                      materialising the L heat-maps)
* `compute_mask_reconstruction_loss`  modules/base_losses/loss_func.py:4-16
* `skeleton_mask_loss`                model.py:91-94 + :185-188 fused: rasterise, max, loss in one pass

All arithmetic is in `csrc/skeleton_mask.cu`; there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _cabi as cabi

__all__ = ["cal_links", "draw_lines", "skeleton_mask", "compute_mask_reconstruction_loss", "skeleton_mask_loss",
           "DrawLines", "SkeletonMask", "MaskReconLoss", "SkeletonMaskLoss"]


def cal_links(parent_ids, line_select_ids=None, use_root=False, extension=True):
    """(parent_ids, child_ids) of the drawn lines, as modules/model.py:8-22."""
    parent_ids = list(parent_ids)
    if use_root:
        child = list(range(len(parent_ids)))
        parent = parent_ids
    else:
        child = list(range(1, len(parent_ids)))
        parent = parent_ids[1:]
    if line_select_ids is None:
        raise TypeError("line_select_ids is required (the reference indexes with it unconditionally, model.py:15)")
    parent = [parent[i] for i in line_select_ids]
    child = [child[i] for i in line_select_ids]
    if extension:
        parent.extend([7, 7, 7, 7, 0, 0, 1, 4])
        child.extend([1, 4, 11, 14, 2, 5, 14, 11])
    return parent, child


def _skel(kp: torch.Tensor, image_size: int, parent_ids: Sequence[int], child_ids: Sequence[int], body_width: float):
    """-> (kp usable in place, xsup_skel_t).  `kp` is `[B, N, 2]` fp32, possibly a strided view such as
    `kps[:, 0, :, :2]` of the head output, which is passed without a copy."""
    cabi.require_cuda(kp, "keypoints")
    if kp.dim() != 3 or kp.shape[-1] != 2:
        raise ValueError("keypoints must be [B, N, 2], got %s" % (tuple(kp.shape),))
    if len(parent_ids) != len(child_ids):
        raise ValueError("parent_ids and child_ids differ in length")
    L = len(parent_ids)
    if not 1 <= L <= cabi.MAX_LINES:
        raise ValueError("%d lines, the kernels take 1..%d" % (L, cabi.MAX_LINES))
    if kp.dtype != torch.float32 or kp.stride(2) != 1 or (kp.shape[0] > 1 and kp.stride(0) < 0) or kp.stride(1) < 2 \
            or kp.data_ptr() % 4:
        kp = kp.to(torch.float32).contiguous()
    B, N, _ = kp.shape
    s = cabi.Skel(B, N, int(image_size), L, kp.stride(0) if B > 1 else 0, kp.stride(1), float(body_width))
    for i, (a, b) in enumerate(zip(parent_ids, child_ids)):
        s.parent[i] = int(a)
        s.child[i] = int(b)
    return kp, s


def _ws(n_floats: int, dev) -> torch.Tensor:
    return torch.empty(max(int(n_floats), 4), dtype=torch.float32, device=dev)


def _mask_cfg(n: int, weight, use_clip: bool) -> cabi.MaskLoss:
    if weight is not None:
        return cabi.MaskLoss(n, cabi.MASK_WEIGHTED, int(bool(use_clip)))
    return cabi.MaskLoss(n, cabi.MASK_CLIP_MEAN if use_clip else cabi.MASK_MSE, int(bool(use_clip)))


def _map_like(t: Optional[torch.Tensor], shape, name: str, dev) -> Optional[torch.Tensor]:
    if t is None:
        return None
    t = t.detach().to(device=dev, dtype=torch.float32)
    if tuple(t.shape) != tuple(shape):
        t = t.expand(shape)
    return t.contiguous()


class DrawLines(torch.autograd.Function):
    """keypoints `[B,N,2]` -> heat-maps `[B,L,S,S]` (util.py:21-59), differentiable in the keypoints."""

    @staticmethod
    def forward(ctx, keypoints, image_size, parent_ids, child_ids, body_width):
        kp, s = _skel(keypoints.detach(), image_size, parent_ids, child_ids, body_width)
        heat = torch.empty(s.B, s.L, s.S, s.S, dtype=torch.float32, device=kp.device)
        with torch.cuda.device(kp.device):
            cabi.check(cabi.lib.xsup_draw_lines_fwd(kp.data_ptr(), s, heat.data_ptr(), cabi.stream_ptr(kp.device)),
                       "xsup_draw_lines_fwd")
        ctx.save_for_backward(kp, heat)
        ctx.skel = s
        return heat

    @staticmethod
    def backward(ctx, g_heat):
        kp, heat = ctx.saved_tensors
        s = ctx.skel
        g_heat = g_heat.to(torch.float32).contiguous()
        g_kp = torch.empty(s.B, s.K, 2, dtype=torch.float32, device=kp.device)
        ws = _ws(cabi.lib.xsup_draw_lines_ws_floats(s), kp.device)
        with torch.cuda.device(kp.device):
            cabi.check(cabi.lib.xsup_draw_lines_bwd(kp.data_ptr(), s, heat.data_ptr(), g_heat.data_ptr(), g_kp.data_ptr(),
                                                    ws.data_ptr(), cabi.stream_ptr(kp.device)), "xsup_draw_lines_bwd")
        return g_kp, None, None, None, None


class SkeletonMaskLoss(torch.autograd.Function):
    """Rasterise the skeleton, take the max over lines and (optionally) the mask-reconstruction loss
    against `gt` / `weight` in one pass (model.py:91-94 + :185-188).

    forward(keypoints [B,N,2], gt [B,1,S,S] | None, weight [B,1,S,S] | None, image_size, parent_ids,
            child_ids, body_width, use_clip) -> (recon [B,1,S,S], loss 0-d)
    `loss` is the value the trainer optimises: the reference's result followed by `.mean()`
    (train.py:182), which only matters for weight=None with use_clip=True where the reference returns a
    tensor.  Without `gt` the loss output is a constant zero.  Gradients reach the keypoints from the
    loss and from any other consumer of `recon` (the physique network, model.py:170)."""

    @staticmethod
    def forward(ctx, keypoints, gt, weight, image_size, parent_ids, child_ids, body_width, use_clip):
        kp, s = _skel(keypoints.detach(), image_size, parent_ids, child_ids, body_width)
        dev = kp.device
        recon = torch.empty(s.B, 1, s.S, s.S, dtype=torch.float32, device=dev)
        line_idx = torch.empty(s.B, s.S, s.S, dtype=torch.uint8, device=dev)
        sums = torch.zeros(cabi.MASK_SUMS, dtype=torch.float32, device=dev)
        ws = _ws(cabi.lib.xsup_skel_ws_floats(s), dev)
        cfg = None
        if gt is not None:
            if s.B == 0:
                raise ValueError("the mask loss needs a non-empty batch")
            gt = _map_like(gt, recon.shape, "gt", dev)
            weight = _map_like(weight, recon.shape, "weight", dev)
            cfg = _mask_cfg(recon.numel(), weight, use_clip)
        with torch.cuda.device(dev):
            cabi.check(cabi.lib.xsup_skeleton_mask_fwd(kp.data_ptr(), s, recon.data_ptr(), line_idx.data_ptr(),
                                                       gt.data_ptr() if cfg else None,
                                                       weight.data_ptr() if cfg and weight is not None else None,
                                                       cfg, sums.data_ptr(), ws.data_ptr(), cabi.stream_ptr(dev)),
                       "xsup_skeleton_mask_fwd")
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(kp, recon, line_idx, sums, gt if cfg else None, weight if cfg else None)
        ctx.skel, ctx.cfg = s, cfg
        return recon, sums[3].clone()

    @staticmethod
    def backward(ctx, g_recon, g_loss):
        kp, recon, line_idx, sums, gt, weight = ctx.saved_tensors
        s, cfg = ctx.skel, ctx.cfg
        dev = kp.device
        if g_recon is None and (g_loss is None or cfg is None):
            return (None,) * 8
        use_loss = cfg is not None and g_loss is not None
        g_kp = torch.empty(s.B, s.K, 2, dtype=torch.float32, device=dev)
        ws = _ws(cabi.lib.xsup_skel_ws_floats(s), dev)
        if g_recon is not None:
            g_recon = g_recon.to(torch.float32).contiguous()
        if use_loss:
            g_loss = g_loss.to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(dev):
            cabi.check(cabi.lib.xsup_skeleton_mask_bwd(
                kp.data_ptr(), s, recon.data_ptr(), line_idx.data_ptr(), g_recon.data_ptr() if g_recon is not None else None,
                gt.data_ptr() if use_loss else None, weight.data_ptr() if use_loss and weight is not None else None,
                cfg if use_loss else None, sums.data_ptr(), g_loss.data_ptr() if use_loss else None, g_kp.data_ptr(),
                ws.data_ptr(), cabi.stream_ptr(dev)), "xsup_skeleton_mask_bwd")
        return (g_kp,) + (None,) * 7


class SkeletonMask(torch.autograd.Function):
    """keypoints `[B,N,2]` -> `[B,1,S,S]` = max over the line heat-maps (model.py:91-94)."""

    @staticmethod
    def forward(ctx, keypoints, image_size, parent_ids, child_ids, body_width):
        recon, _ = SkeletonMaskLoss.forward(ctx, keypoints, None, None, image_size, parent_ids, child_ids, body_width, False)
        return recon

    @staticmethod
    def backward(ctx, g_recon):
        return SkeletonMaskLoss.backward(ctx, g_recon, None)[:5]


class MaskReconLoss(torch.autograd.Function):
    """compute_mask_reconstruction_loss on an arbitrary mask tensor (loss_func.py:4-16).
    forward(mask, gt, weight | None, use_clip, want_filter) -> (loss 0-d, filter | None)"""

    @staticmethod
    def forward(ctx, mask, gt, weight, mode, use_clip, want_filter):
        cabi.require_cuda(mask, "mask")
        dev = mask.device
        m = mask.detach().to(torch.float32).contiguous()
        if m.numel() == 0:
            raise ValueError("the mask loss needs a non-empty mask")
        gt = _map_like(gt, m.shape, "gt", dev)
        weight = _map_like(weight, m.shape, "weight", dev)
        cfg = cabi.MaskLoss(m.numel(), mode, int(bool(use_clip)))
        sums = torch.empty(cabi.MASK_SUMS, dtype=torch.float32, device=dev)
        filt = torch.empty_like(m) if want_filter else None
        ws = _ws(cabi.lib.xsup_mask_loss_ws_floats(m.numel()), dev)
        with torch.cuda.device(dev):
            cabi.check(cabi.lib.xsup_mask_loss_fwd(m.data_ptr(), gt.data_ptr(), weight.data_ptr() if weight is not None else None,
                                                   filt.data_ptr() if want_filter else None, cfg, sums.data_ptr(), ws.data_ptr(),
                                                   cabi.stream_ptr(dev)), "xsup_mask_loss_fwd")
        ctx.save_for_backward(m, gt, weight, sums)
        ctx.cfg, ctx.in_dtype = cfg, mask.dtype
        if want_filter:
            ctx.mark_non_differentiable(filt)
        return sums[3].clone(), filt

    @staticmethod
    def backward(ctx, g_loss, _g_filter):
        m, gt, weight, sums = ctx.saved_tensors
        dev = m.device
        g_loss = g_loss.to(torch.float32).reshape(1).contiguous()
        g_mask = torch.empty_like(m)
        with torch.cuda.device(dev):
            cabi.check(cabi.lib.xsup_mask_loss_bwd(m.data_ptr(), gt.data_ptr(), weight.data_ptr() if weight is not None else None,
                                                   ctx.cfg, sums.data_ptr(), g_loss.data_ptr(), g_mask.data_ptr(),
                                                   cabi.stream_ptr(dev)), "xsup_mask_loss_bwd")
        return g_mask.to(ctx.in_dtype), None, None, None, None, None


# --------------------------------------------------------------------------------------- reference-signature functions
def draw_lines(keypoints, image_size, parent_ids, child_ids, body_width):
    """Same signature as modules/util.py:21."""
    return DrawLines.apply(keypoints, int(image_size), tuple(parent_ids), tuple(child_ids), float(body_width))


def skeleton_mask(keypoints, image_size, parent_ids, child_ids, body_width):
    """`torch.max(draw_lines(...), dim=1, keepdim=True)[0]` (model.py:91-94) without the L heat-maps."""
    return SkeletonMask.apply(keypoints, int(image_size), tuple(parent_ids), tuple(child_ids), float(body_width))


def compute_mask_reconstruction_loss(mask, gt, weight=None, use_clip=False):
    """Same signature and return value as modules/base_losses/loss_func.py:4.  With `weight=None` and
    `use_clip=True` the reference returns a tensor (scalar MSE times the `mask > 0.1` map); that one product is
    formed here by broadcasting our scalar against the kernel-written filter map so that the result and its
    autograd are the reference's."""
    if weight is not None:
        return MaskReconLoss.apply(mask, gt, weight, cabi.MASK_WEIGHTED, use_clip, False)[0]
    if not use_clip:
        return MaskReconLoss.apply(mask, gt, None, cabi.MASK_MSE, False, False)[0]
    mse, filt = MaskReconLoss.apply(mask, gt, None, cabi.MASK_MSE, True, True)
    return mse * filt


def skeleton_mask_loss(keypoints, gt, weight, image_size, parent_ids, child_ids, body_width, use_clip=True):
    """Fused `draw_lines` -> max -> `compute_mask_reconstruction_loss(..., use_clip).mean()`; returns (recon, loss)."""
    return SkeletonMaskLoss.apply(keypoints, gt, weight, int(image_size), tuple(parent_ids), tuple(child_ids),
                                  float(body_width), bool(use_clip))
