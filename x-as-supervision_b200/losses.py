"""Stand-alone pose loss terms over the C ABI, with the reference's signatures
(modules/base_losses/loss_func.py:18-52): `compute_supervision`, `compute_bone_sym_loss`, `compute_kp_sym_loss`.

Inside the per-camera loss graph these are evaluated per hypothesis by the fused op
(`ops.integral_reproj_min_loss`); the functions here are for callers that use a term on its own.
All arithmetic is in `csrc/eval_disc.cu`; there is no CPU path."""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi as cabi

__all__ = ["compute_supervision", "compute_bone_sym_loss", "compute_kp_sym_loss", "PoseTerm", "PoseSqErr"]


class PoseTerm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gt, term, flag, feature_shape):
        cabi.require_cuda(x, "keypoints")
        xf = x.detach().to(torch.float32).contiguous()
        if xf.dim() != 3:
            raise ValueError("keypoints must be [B, K, C], got %s" % (tuple(x.shape),))
        B, K, Cc = xf.shape
        dev = xf.device
        g = None
        if gt is not None:
            g = gt.detach().to(device=dev, dtype=torch.float32).expand_as(xf).contiguous()
        fs = (C.c_float * 3)(*[float(v) for v in (list(feature_shape) + [1.0, 1.0, 1.0])[:3]]) if feature_shape is not None else None
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ws = torch.empty(max(B, 1), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            cabi.check(cabi.lib.xsup_pose_term_fwd(xf.data_ptr(), g.data_ptr() if g is not None else None, fs, term, flag, B, K, Cc,
                                                   ws.data_ptr(), loss.data_ptr(), cabi.stream_ptr(dev)), "xsup_pose_term_fwd")
        ctx.save_for_backward(xf, g)
        ctx.meta = (term, flag, feature_shape, x.dtype)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        xf, g = ctx.saved_tensors
        term, flag, feature_shape, in_dtype = ctx.meta
        B, K, Cc = xf.shape
        dev = xf.device
        fs = (C.c_float * 3)(*[float(v) for v in (list(feature_shape) + [1.0, 1.0, 1.0])[:3]]) if feature_shape is not None else None
        gl = g_loss.to(torch.float32).reshape(1).contiguous()
        gx = torch.empty_like(xf)
        with torch.cuda.device(dev):
            cabi.check(cabi.lib.xsup_pose_term_bwd(xf.data_ptr(), g.data_ptr() if g is not None else None, fs, term, flag, B, K, Cc,
                                                   gl.data_ptr(), gx.data_ptr(), cabi.stream_ptr(dev)), "xsup_pose_term_bwd")
        return gx.to(in_dtype), None, None, None, None


class PoseSqErr(torch.autograd.Function):
    """Element-wise squared error `[B,K,C]` (nn.MSELoss(reduction='none'), loss_func.py:46-47) and its VJP."""

    @staticmethod
    def forward(ctx, x, gt, feature_shape):
        cabi.require_cuda(x, "keypoints")
        xf = x.detach().to(torch.float32).contiguous()
        if xf.dim() != 3:
            raise ValueError("keypoints must be [B, K, C], got %s" % (tuple(x.shape),))
        B, K, Cc = xf.shape
        dev = xf.device
        g = gt.detach().to(device=dev, dtype=torch.float32).expand_as(xf).contiguous()
        fs = (C.c_float * 3)(*[float(v) for v in (list(feature_shape) + [1.0, 1.0, 1.0])[:3]]) if feature_shape is not None else None
        out = torch.empty_like(xf)
        with torch.cuda.device(dev):
            cabi.check(cabi.lib.xsup_pose_sqerr(xf.data_ptr(), g.data_ptr(), fs, B, K, Cc, None, out.data_ptr(), cabi.stream_ptr(dev)),
                       "xsup_pose_sqerr")
        ctx.save_for_backward(xf, g)
        ctx.meta = (feature_shape, x.dtype)
        return out

    @staticmethod
    def backward(ctx, g_out):
        xf, g = ctx.saved_tensors
        feature_shape, in_dtype = ctx.meta
        B, K, Cc = xf.shape
        dev = xf.device
        fs = (C.c_float * 3)(*[float(v) for v in (list(feature_shape) + [1.0, 1.0, 1.0])[:3]]) if feature_shape is not None else None
        go = g_out.to(torch.float32).contiguous()
        gx = torch.empty_like(xf)
        with torch.cuda.device(dev):
            cabi.check(cabi.lib.xsup_pose_sqerr(xf.data_ptr(), g.data_ptr(), fs, B, K, Cc, go.data_ptr(), gx.data_ptr(),
                                                cabi.stream_ptr(dev)), "xsup_pose_sqerr (vjp)")
        return gx.to(in_dtype), None, None


def compute_supervision(keypoint, keypoint_gt, feature_shape=None, mode="mean"):
    """Same signature as loss_func.py:38.  `mode` is nn.MSELoss's reduction: 'mean', 'sum' (then divided by the batch
    size, :50-51) or 'none' (the element-wise `[B,K,C]` tensor)."""
    if mode == "none":
        return PoseSqErr.apply(keypoint, keypoint_gt, feature_shape)
    if mode not in ("mean", "sum"):
        raise ValueError("mode must be 'mean', 'sum' or 'none', got %r" % (mode,))
    return PoseTerm.apply(keypoint, keypoint_gt, cabi.TERM_MSE, int(mode == "sum"), feature_shape)


def compute_bone_sym_loss(keypoints):
    """Same signature as loss_func.py:18: arm / leg bone pairs (8 bones) of equal length, world mm -> m."""
    return PoseTerm.apply(keypoints, None, cabi.TERM_BONE, 0, None)


def compute_kp_sym_loss(keypoints, is_3D=True):
    """Same signature as loss_func.py:27: midpoints of (11,14) and (1,4) against joints K-1 and 0."""
    return PoseTerm.apply(keypoints, None, cabi.TERM_KP, int(bool(is_3D)), None)
