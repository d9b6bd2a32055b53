"""Drop-in `torch.autograd.Function`s over the C ABI (include/xsup_b200.h).

Tensor signatures follow the reference so its detector, physique network, GCN
discriminator and trainer can call these unchanged:

* `IntegralMultiHead`      replaces modules/keypoint_detector_integral_multi.py:69-88
* `IntegralSingleHead`     replaces modules/keypoint_detector_integral.py:45-65
* `PatchToWorld` / `convert_patch_to_world` / `convert_world_to_patch`
                           replace modules/util.py:128-152 / :155-168; `convert_patch_to_image`, `convert_image_to_world`,
                           `convert_image_to_patch`, `convert_world_to_image` replace :61-125 (all differentiable)
* `IntegralReprojMinLoss`  replaces, for one camera, modules/model.py:64,71-79,105-114,158-162
                           (head + per-hypothesis world lift + loss terms + torch.min over slots)

PyTorch is used for device memory, streams and `torch.distributed` only; all
arithmetic happens in the CUDA kernels.  Nothing here falls back to the CPU.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _cabi as cabi
from . import dist as xdist

__all__ = ["IntegralMultiHead", "IntegralSingleHead", "PatchToWorld", "IntegralReprojMinLoss", "integral_multi_head",
           "integral_single_head", "convert_patch_to_world", "convert_world_to_patch", "convert_patch_to_image",
           "convert_image_to_world", "convert_image_to_patch", "convert_world_to_image", "find_peak",
           "integral_reproj_min_loss", "launch_count", "set_event_sink", "GraphedReprojStep", "conv_integral_head", "ConvIntegralHead",
           "conv_integral_head_train", "ConvIntegralReprojMinLoss", "conv_integral_reproj_min_loss"]

launch_count = cabi.launch_count
_nvtx = torch.cuda.nvtx          # ranges around K1 / K2 (+ exchange) / K3: visible in nsys / ncu --nvtx, free otherwise
_event_sink = None               # bench.py: dict name -> [(start, stop) CUDA events] recorded on the launching stream


def set_event_sink(sink) -> None:
    """Measurement hook: with a dict, every kernel range below also records a CUDA-event pair on the current stream
    into `sink[range name]`; `None` switches it off.  Not for use under CUDA-graph capture."""
    global _event_sink
    _event_sink = sink


class _range:
    """NVTX range (and, when a sink is set, a CUDA-event pair) around the launches of one kernel group."""
    __slots__ = ("name", "ev")

    def __init__(self, name):
        self.name = name
        self.ev = None

    def __enter__(self):
        _nvtx.range_push(self.name)
        if _event_sink is not None:
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()

    def __exit__(self, *exc):
        if self.ev is not None:
            self.ev[1].record()
            _event_sink.setdefault(self.name, []).append(self.ev)
        _nvtx.range_pop()
        return False


def _volume_dims(logits: torch.Tensor, num_kp: int) -> Tuple[int, int, int, int, int]:
    if logits.dim() != 4:
        raise ValueError("logits must be [B, K*D, H, W], got %s" % (tuple(logits.shape),))
    B, C, H, W = logits.shape
    if C % num_kp:
        raise ValueError("channel count %d is not a multiple of num_kp %d" % (C, num_kp))
    return B, C // num_kp, H, W, C


def _head_forward(logits, num_kp, num_hypo, neighbor_size, head):
    cabi.require_cuda(logits, "logits")
    if not logits.is_contiguous():
        logits = logits.contiguous()
    B, D, H, W, _ = _volume_dims(logits, num_kp)
    shape = cabi.make_shape(B, num_kp, D, H, W, num_hypo, neighbor_size, logits.dtype, head)
    dev = logits.device
    kps = torch.empty(B, num_hypo, num_kp, 3, dtype=torch.float32, device=dev)
    dmap = torch.zeros(num_kp, D, dtype=torch.float32, device=dev) if B == 0 else \
        torch.empty(num_kp, D, dtype=torch.float32, device=dev)
    idx = torch.empty(B, num_kp, num_hypo, dtype=torch.int64, device=dev)
    stats = torch.empty(cabi.lib.xsup_stats_floats(shape), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev), _range("xsup.K1.integral_fwd"):
        cabi.check(cabi.lib.xsup_integral_fwd(logits.data_ptr(), kps.data_ptr(), dmap.data_ptr(),
                                              idx.data_ptr() if head == cabi.HEAD_MULTI else None,
                                              stats.data_ptr(), shape, cabi.stream_ptr(dev)), "xsup_integral_fwd")
    return logits, shape, kps, dmap, idx, stats


def _head_backward(logits, stats, shape, g_kps, inplace=False):
    dev = logits.device
    g_kps = g_kps.to(torch.float32).contiguous()
    g_logits = logits if inplace else torch.empty_like(logits)
    coef = torch.empty(cabi.lib.xsup_coef_floats(shape), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev), _range("xsup.K3.integral_bwd"):
        cabi.check(cabi.lib.xsup_integral_bwd(logits.data_ptr(), stats.data_ptr(), g_kps.data_ptr(), g_logits.data_ptr(),
                                              coef.data_ptr(), shape, cabi.stream_ptr(dev)), "xsup_integral_bwd")
    return g_logits


class IntegralMultiHead(torch.autograd.Function):
    """logits `[B, K*D, H, W]` (fp32 | bf16) -> kps `[B,NH,K,3]` fp32, depth_prob_map `[K,D]` fp32,
    peak_idx `[B,K,NH]` int64.  Only `kps` is differentiable (depth_prob_map is a TensorBoard
    side output in the reference, train_util.py:296-299)."""

    @staticmethod
    def forward(ctx, logits, num_kp, num_hypo, neighbor_size):
        logits, shape, kps, dmap, idx, stats = _head_forward(logits, num_kp, num_hypo, neighbor_size, cabi.HEAD_MULTI)
        ctx.save_for_backward(logits, stats)
        ctx.shape = shape
        ctx.mark_non_differentiable(dmap, idx)
        return kps, dmap, idx

    @staticmethod
    def backward(ctx, g_kps, _g_dmap, _g_idx):
        logits, stats = ctx.saved_tensors
        if g_kps is None:
            return None, None, None, None
        return _head_backward(logits, stats, ctx.shape, g_kps), None, None, None


class IntegralSingleHead(torch.autograd.Function):
    """Single-hypothesis variant: kps `[B,1,K,3]`, depth_prob_map `[K,D]`."""

    @staticmethod
    def forward(ctx, logits, num_kp):
        logits, shape, kps, dmap, _, stats = _head_forward(logits, num_kp, 1, 1, cabi.HEAD_SINGLE)
        ctx.save_for_backward(logits, stats)
        ctx.shape = shape
        ctx.mark_non_differentiable(dmap)
        return kps, dmap

    @staticmethod
    def backward(ctx, g_kps, _g_dmap):
        logits, stats = ctx.saved_tensors
        return _head_backward(logits, stats, ctx.shape, g_kps), None


def integral_multi_head(logits, num_kp, num_hypo, neighbor_size):
    return IntegralMultiHead.apply(logits, num_kp, num_hypo, neighbor_size)


def integral_single_head(logits, num_kp):
    return IntegralSingleHead.apply(logits, num_kp)


def find_peak(pz: torch.Tensor, num_hypo: int) -> torch.Tensor:
    """KPDetector3DMulti.find_peak (…_multi.py:24-34) on `[..., D]` -> int64 `[..., NH]`."""
    cabi.require_cuda(pz, "heatmap")
    flat = pz.detach().to(torch.float32).contiguous().view(-1, pz.shape[-1])
    idx = torch.empty(flat.shape[0], num_hypo, dtype=torch.int64, device=pz.device)
    with torch.cuda.device(pz.device):
        cabi.check(cabi.lib.xsup_find_peak(flat.data_ptr(), idx.data_ptr(), flat.shape[0], flat.shape[1], num_hypo,
                                           cabi.stream_ptr(pz.device)), "xsup_find_peak")
    return idx.view(*pz.shape[:-1], num_hypo)


# --------------------------------------------------------------------------------------- geometry
def _cam_args(cams: Dict[str, torch.Tensor], B: int):
    t = [cams[k].to(torch.float32).contiguous() for k in ("trans_image", "pelvis", "k_mat", "trans_world", "rot_world")]
    return t, cabi.make_cam(*t, B)


def _flags(is_norm, mono, patch):
    return (cabi.FLAG_NORM if is_norm else 0) | (cabi.FLAG_MONO if mono else 0) | (cabi.FLAG_PATCH if patch else 0)


def _f32c(t, dev, shape=None, name="tensor"):
    t = t.detach().to(device=dev, dtype=torch.float32)
    if shape is not None:
        if t.numel() != int(torch.Size(shape).numel()):
            raise ValueError("%s has %d elements, expected shape %s" % (name, t.numel(), tuple(shape)))
        t = t.reshape(shape)
    return t.contiguous()


class _GeomStage(torch.autograd.Function):
    """One call of `xsup_geom_patch_to_world` (direction 0) / `xsup_geom_world_to_patch` (direction 1) with the stages
    `flags` select; differentiable in the points (the camera tensors are data-loader constants, as in the reference).

    `intr` is a `[B,3,3]` k_mat (stride 9) or a tuple (fx, fy, cx, cy) of `[B]`-sized tensors (stride 1), or None."""

    @staticmethod
    def forward(ctx, pts, direction, flags, img_dhw, depth_scale, trans_image, pelvis, intr, trans_world, rot_world):
        cabi.require_cuda(pts, "keypoints")
        if pts.dim() != 3 or pts.shape[-1] != 3:
            raise ValueError("keypoints must be [B, J, 3], got %s" % (tuple(pts.shape),))
        dev = pts.device
        x = pts.detach().to(torch.float32).contiguous()
        B, J, _ = x.shape
        keep = []
        g = cabi.Geom(B, J, int(img_dhw[0]), int(img_dhw[1]), int(img_dhw[2]), float(depth_scale), int(flags), 1)
        if flags & cabi.GEOM_PATCH_STAGE:
            ti, pv = _f32c(trans_image, dev, (B, 2, 3), "trans"), _f32c(pelvis, dev, (B, 3), "pelvis")
            keep += [ti, pv]
            g.trans_image, g.pelvis = ti.data_ptr(), pv.data_ptr()
        if (flags & cabi.GEOM_CAMERA_STAGE) and not (flags & cabi.GEOM_MONO):
            if isinstance(intr, (tuple, list)):
                fx, fy, cx, cy = (_f32c(t, dev, (B,), "intrinsic") for t in intr)
                keep += [fx, fy, cx, cy]
                g.fx, g.fy, g.cx, g.cy, g.intr_stride = fx.data_ptr(), fy.data_ptr(), cx.data_ptr(), cy.data_ptr(), 1
            else:
                km = _f32c(intr, dev, (B, 3, 3), "k_mat")
                keep.append(km)
                base = km.data_ptr()
                g.fx, g.fy, g.cx, g.cy, g.intr_stride = base, base + 16, base + 8, base + 20, 9
            tw, rw = _f32c(trans_world, dev, (B, 3), "trans_world"), _f32c(rot_world, dev, (B, 3, 3), "rot_world")
            keep += [tw, rw]
            g.trans_world, g.rot_world = tw.data_ptr(), rw.data_ptr()
        out = torch.empty_like(x)
        fn = cabi.lib.xsup_geom_world_to_patch if direction else cabi.lib.xsup_geom_patch_to_world
        with torch.cuda.device(dev):
            cabi.check(fn(x.data_ptr(), out.data_ptr(), g, cabi.stream_ptr(dev)), "xsup_geom")
        ctx.save_for_backward(x, *keep)
        ctx.geom, ctx.direction, ctx.in_dtype = g, direction, pts.dtype
        return out.to(pts.dtype) if pts.dtype != torch.float32 else out

    @staticmethod
    def backward(ctx, g_out):
        x, *_keep = ctx.saved_tensors                       # `_keep` holds the tensors ctx.geom points into
        dev = x.device
        go = g_out.to(torch.float32).contiguous()
        gi = torch.empty_like(x)
        fn = cabi.lib.xsup_geom_world_to_patch_vjp if ctx.direction else cabi.lib.xsup_geom_patch_to_world_vjp
        with torch.cuda.device(dev):
            cabi.check(fn(x.data_ptr(), go.data_ptr(), gi.data_ptr(), ctx.geom, cabi.stream_ptr(dev)), "xsup_geom_vjp")
        return (gi.to(ctx.in_dtype),) + (None,) * 9


class PatchToWorld:
    """kps `[B,J,3]` + per-sample camera tensors -> world `[B,J,3]` (util.py:128-152), differentiable in kps:
    `_GeomStage` with both stages selected and the data loader's tensors (`k_mat`, depth extent = image width)."""

    @staticmethod
    def apply(kps, trans_image, pelvis, k_mat, trans_world, rot_world, img_h, img_w, rect_width, flags):
        return _GeomStage.apply(kps, 0, (flags & 7) | cabi.GEOM_CAMERA_STAGE, (img_w, img_h, img_w), float(rect_width) / img_w,
                                trans_image, pelvis, k_mat, trans_world, rot_world)


def _unpack_params(params: Dict[str, torch.Tensor], mode: str):
    cams = {k: params["{}_{}".format(mode, k)] for k in ("trans_image", "pelvis", "k_mat", "trans_world", "rot_world")}
    shape_img = params["{}_img".format(mode)].shape
    return cams, int(shape_img[-2]), int(shape_img[-1])


def convert_patch_to_world(keypoints, params, mode, is_norm=True, RECT_WIDTH=2000, mono=False, patch=True):
    """Same dict-keyed signature as modules/util.py:128 so model.py / eval.py can import this one."""
    cams, img_h, img_w = _unpack_params(params, mode)
    return PatchToWorld.apply(keypoints, cams["trans_image"], cams["pelvis"], cams["k_mat"], cams["trans_world"],
                              cams["rot_world"], img_h, img_w, float(RECT_WIDTH), _flags(is_norm, mono, patch))


def convert_world_to_patch(keypoints, params, mode, is_norm=True, RECT_WIDTH=2000):
    """modules/util.py:155-168 (forward perspective projection, the inverse of `convert_patch_to_world`);
    differentiable in the keypoints."""
    cams, img_h, img_w = _unpack_params(params, mode)
    flags = (cabi.GEOM_NORM if is_norm else 0) | cabi.GEOM_PATCH_STAGE | cabi.GEOM_CAMERA_STAGE
    return _GeomStage.apply(keypoints, 1, flags, (img_w, img_h, img_w), float(RECT_WIDTH) / img_w, cams["trans_image"],
                            cams["pelvis"], cams["k_mat"], cams["trans_world"], cams["rot_world"])


def convert_patch_to_image(kps, trans, image_depth, image_height, image_width, depth_scale, pelvis, is_norm=True):
    """modules/util.py:61-82, same signature: inverse crop affine, depth px -> mm + pelvis depth."""
    flags = (cabi.GEOM_NORM if is_norm else 0) | cabi.GEOM_PATCH_STAGE
    return _GeomStage.apply(kps, 0, flags, (image_depth, image_height, image_width), depth_scale, trans, pelvis, None, None, None)


def convert_image_to_world(kps, fx, fy, u, v, trans, rot):
    """modules/util.py:85-95, same signature: pinhole back-projection, inverse extrinsics (general 3x3 inverse)."""
    return _GeomStage.apply(kps, 0, cabi.GEOM_CAMERA_STAGE, (2, 2, 2), 1.0, None, None, (fx, fy, u, v), trans, rot)


def convert_image_to_patch(kps, trans, image_depth, image_height, image_width, depth_scale, pelvis, is_norm=True):
    """modules/util.py:98-113, same signature: mm -> depth px about the pelvis, crop affine, optional normalisation."""
    flags = (cabi.GEOM_NORM if is_norm else 0) | cabi.GEOM_PATCH_STAGE
    return _GeomStage.apply(kps, 1, flags, (image_depth, image_height, image_width), depth_scale, trans, pelvis, None, None, None)


def convert_world_to_image(kps, fx, fy, u, v, trans, rot):
    """modules/util.py:116-125, same signature: extrinsics then perspective division."""
    return _GeomStage.apply(kps, 1, cabi.GEOM_CAMERA_STAGE, (2, 2, 2), 1.0, None, None, (fx, fy, u, v), trans, rot)


# --------------------------------------------------------------------------------------- fused head + loss
class IntegralReprojMinLoss(torch.autograd.Function):
    """Head + per-hypothesis world lift + loss terms + min over hypotheses for one camera.

    forward(logits, target, trans_image, pelvis, k_mat, trans_world, rot_world,
            num_kp, num_hypo, neighbor_size, img_hw, rect_width,
            w_mse, w_bone, w_kp, w_kp2d, reduction, group)
      -> loss_pseudo (0-d), loss_sym (0-d), sel_idx (int64), kps [B,NH,K,3], kps_world [B,NH,K,3],
         depth_prob_map [K,D], peak_idx [B,K,NH]

    `w_bone/w_kp/w_kp2d = None` means the symmetry term is absent (SurS1 configs).
    `group`: a `dist.PeerExchange` (in-kernel NVLink exchange, preferred) or a torch.distributed
    process group (NCCL) -> 'global' scope: the per-hypothesis partial sums are all-reduced (one
    [4,NH] fp32 message) so every rank selects the slot the single-process reference would select on
    the global batch.  `None` -> rank-local min, which is what the reference does under DDP
    (model.py:114,162).
    Gradients flow to `logits` from both losses and from anything downstream of `kps` / `kps_world`."""

    @staticmethod
    def forward(ctx, logits, target, trans_image, pelvis, k_mat, trans_world, rot_world, num_kp, num_hypo, neighbor_size,
                img_hw, rect_width, w_mse, w_bone, w_kp, w_kp2d, reduction, group):
        logits, shape, kps, dmap, idx, stats = _head_forward(logits, num_kp, num_hypo, neighbor_size, cabi.HEAD_MULTI)
        ctx.set_materialize_grads(False)          # unused outputs (kps, kps_world, loss_sym) arrive as None, not zeros
        dev = logits.device
        B, K, NH = shape.B, num_kp, num_hypo
        if B == 0:
            raise ValueError("IntegralReprojMinLoss needs a non-empty batch")
        target = target.to(device=dev, dtype=torch.float32).contiguous()
        if tuple(target.shape) != (B, K, 3):
            raise ValueError("target must be [B,K,3] = %s, got %s" % ((B, K, 3), tuple(target.shape)))
        keep, cam = _cam_args(dict(trans_image=trans_image, pelvis=pelvis, k_mat=k_mat, trans_world=trans_world,
                                   rot_world=rot_world), B)
        use_sym = any(w is not None for w in (w_bone, w_kp, w_kp2d))
        n_total = xdist.global_batch(B, group)
        cfg = cabi.LossCfg(B, K, NH, int(img_hw[0]), int(img_hw[1]), float(rect_width), float(w_mse),
                           float(w_bone or 0.0), float(w_kp or 0.0), float(w_kp2d or 0.0), int(use_sym),
                           cabi.REDUCE[reduction], n_total)
        world = torch.empty_like(kps)
        sample_terms = torch.empty(B, cabi.LOSS_TERMS, NH, dtype=torch.float32, device=dev)
        partial = torch.empty(cabi.LOSS_TERMS, NH, dtype=torch.float32, device=dev)
        loss = torch.empty(2, dtype=torch.float32, device=dev)
        sel_shape = {"batch": (2,), "sample": (2, B), "joint": (B, K)}[reduction]
        sel = torch.empty(sel_shape, dtype=torch.int64, device=dev)
        st = cabi.stream_ptr(dev)
        nccl_group = xdist.is_active(group) and not isinstance(group, xdist.PeerExchange)
        with torch.cuda.device(dev), _range("xsup.K2.loss_select_fwd"):
            if not nccl_group:
                # ONE launch: world lift + loss terms, last CTA done -> fixed-order batch sums -> NVLink exchange -> slots
                xchg = group.descriptor() if xdist.is_active(group) else None
                ticket = stats.data_ptr() + 4 * (B * K * int(cabi.lib.xsup_stats_stride(shape)) + 1)
                cabi.check(cabi.lib.xsup_reproj_fused_fwd(kps.data_ptr(), target.data_ptr(), cam, world.data_ptr(),
                                                          sample_terms.data_ptr(), partial.data_ptr(), loss.data_ptr(),
                                                          sel.data_ptr(), cfg, xchg, ticket, st), "xsup_reproj_fused_fwd")
            else:
                # a torch.distributed process group (NCCL): the all-reduce is a library call between two of our launches
                cabi.check(cabi.lib.xsup_reproj_loss_fwd(kps.data_ptr(), target.data_ptr(), cam, world.data_ptr(),
                                                         sample_terms.data_ptr(), partial.data_ptr(), cfg, st), "xsup_reproj_loss_fwd")
                if reduction == "batch":
                    xdist.reduce_partials(partial, group)        # the one exchange step of the path
                cabi.check(cabi.lib.xsup_reproj_select(kps.data_ptr(), target.data_ptr(), sample_terms.data_ptr(),
                                                       partial.data_ptr(), loss.data_ptr(), sel.data_ptr(), cfg, st), "xsup_reproj_select")
                if reduction != "batch":
                    xdist.reduce_partials(loss, group)           # reporting only: the gradient needs just n_total
        ctx.save_for_backward(logits, stats, kps, target, sel, *keep)
        ctx.shape, ctx.cfg, ctx.cam = shape, cfg, cam          # `cam` holds raw pointers into `keep`, which the line above keeps alive
        ctx.mark_non_differentiable(sel, dmap, idx)
        return loss[0], loss[1], sel, kps, world, dmap, idx

    @staticmethod
    def backward(ctx, g_lp, g_ls, _g_sel, g_kps_out, g_world, _g_dmap, _g_idx):
        logits, stats, kps, target, sel, *keep = ctx.saved_tensors
        shape, cfg, cam = ctx.shape, ctx.cfg, ctx.cam
        dev = logits.device
        if g_lp is None and g_ls is None and g_kps_out is None and g_world is None:
            return (None,) * 18

        def ptr(t):                                   # fp32 contiguous device tensor or NULL; no kernels for fp32 inputs
            if t is None:
                return None, None
            t = t.to(torch.float32).contiguous()
            return t, t.data_ptr()
        g_lp, p_lp = ptr(g_lp)
        g_ls, p_ls = ptr(g_ls)
        g_kps_out, p_gk = ptr(g_kps_out)              # downstream use of kps (draw_lines, model.py:91)
        g_world, p_gw = ptr(g_world)                  # downstream use of kps_world (generator loss with use_aug, model.py:138)
        coef = torch.empty(cabi.lib.xsup_coef_floats(shape), dtype=torch.float32, device=dev)
        g_logits = torch.empty_like(logits)
        st = cabi.stream_ptr(dev)
        with torch.cuda.device(dev):
            with _range("xsup.K2.loss_bwd_coef"):
                # ONE launch: loss VJP of every hypothesis + upstream gradients -> coefficient blocks; d loss / d kps stays on chip
                cabi.check(cabi.lib.xsup_reproj_fused_bwd(kps.data_ptr(), target.data_ptr(), cam, sel.data_ptr(), p_lp, p_ls, p_gk,
                                                          p_gw, stats.data_ptr(), coef.data_ptr(), None, cfg, shape, st),
                           "xsup_reproj_fused_bwd")
            with _range("xsup.K3.integral_bwd"):
                cabi.check(cabi.lib.xsup_integral_bwd_apply(logits.data_ptr(), coef.data_ptr(), g_logits.data_ptr(), shape, st),
                           "xsup_integral_bwd_apply")
        return (g_logits,) + (None,) * 17


def integral_reproj_min_loss(logits, target, cams: Dict[str, torch.Tensor], num_kp, num_hypo, neighbor_size,
                             img_hw: Sequence[int] = (256, 256), rect_width: float = 2000.0, w_mse: float = 1.0,
                             w_bone: Optional[float] = None, w_kp: Optional[float] = None, w_kp2d: Optional[float] = None,
                             reduction: str = "batch", group=None):
    """Functional form of `IntegralReprojMinLoss` taking the camera tensors as a dict keyed
    trans_image / pelvis / k_mat / trans_world / rot_world."""
    return IntegralReprojMinLoss.apply(logits, target, cams["trans_image"], cams["pelvis"], cams["k_mat"],
                                       cams["trans_world"], cams["rot_world"], num_kp, num_hypo, neighbor_size,
                                       tuple(img_hw), rect_width, w_mse, w_bone, w_kp, w_kp2d, reduction, group)


class GraphedReprojStep:
    """The fused head + reprojection min-loss forward AND backward captured once into a CUDA graph.

    Every C-ABI call is capture-safe (caller's stream, no allocation, no synchronisation), so for fixed shapes the
    4 launches (+ the autograd glue) and the Python/autograd bookkeeping of a step collapse into one `cudaGraphLaunch`: this is what
    removes the host-side floor (~0.3 ms) that small batches / small volumes otherwise sit on.

        step = GraphedReprojStep(logits, target, cams, num_kp, num_hypo, neighbor_size, **loss_kwargs)
        step.logits.copy_(new_logits); ...                # or produce the inputs directly into the static buffers
        loss_pseudo, loss_sym, sel = step()               # replays; step.grad holds d(loss_pseudo + loss_sym)/d logits

    Static tensors: `logits`, `target`, `cams[...]` (inputs), `grad`, `kps`, `kps_world`, `loss_pseudo`, `loss_sym`,
    `sel` (outputs).  `group` may be None (rank-local selection) or a `dist.PeerExchange` (global selection: its
    in-kernel NVLink all-reduce keeps the call sequence number on the device, so it replays correctly; every rank
    must then replay the same number of times).  A `torch.distributed` process group is not captured here."""

    def __init__(self, logits, target, cams, num_kp, num_hypo, neighbor_size, warmup: int = 3, **kw):
        if kw.get("group") is not None and not isinstance(kw["group"], xdist.PeerExchange):
            raise ValueError("GraphedReprojStep takes group=None or a dist.PeerExchange")
        cabi.require_cuda(logits, "logits")
        dev = logits.device
        self.logits = logits.detach().clone().requires_grad_(True)
        self.target = target.detach().to(dev).clone()
        self.cams = {k: v.detach().to(dev).clone() for k, v in cams.items()}
        args = (num_kp, num_hypo, neighbor_size)

        def run():
            self.logits.grad = None
            out = integral_reproj_min_loss(self.logits, self.target, self.cams, *args, **kw)
            (out[0] + out[1]).backward()
            return out
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                       # warm-up off the capture stream, as torch.cuda.graph requires
            for _ in range(max(1, warmup)):
                run()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        self.logits.grad = None
        with torch.cuda.graph(self.graph):
            out = run()
        self.loss_pseudo, self.loss_sym, self.sel, self.kps, self.kps_world = (o.detach() for o in out[:5])
        self.grad = self.logits.grad

    def __call__(self):
        self.graph.replay()
        return self.loss_pseudo, self.loss_sym, self.sel


def _nhwc_bf16(x: torch.Tensor) -> torch.Tensor:
    """`x [B,C,H,W]` as bf16 with channels-last storage (the K-major operand of the conv-fused kernels): used in place when
    it already is; a contiguous fp32 NCHW tensor goes through `xsup_pack_nhwc_bf16` (transpose + rounding in one pass);
    anything else through torch's conversion."""
    xb = x.detach()
    if xb.dtype == torch.bfloat16 and xb.is_contiguous(memory_format=torch.channels_last):
        return xb
    B, C, H, W = xb.shape
    if xb.dtype == torch.float32 and xb.is_contiguous() and C % 64 == 0 and (H * W) % 64 == 0 and B > 0:
        packed = torch.empty((B, C, H, W), dtype=torch.bfloat16, device=xb.device, memory_format=torch.channels_last)
        with torch.cuda.device(xb.device):
            cabi.check(cabi.lib.xsup_pack_nhwc_bf16(xb.data_ptr(), packed.data_ptr(), B, C, H * W, cabi.stream_ptr(xb.device)),
                       "xsup_pack_nhwc_bf16")
        return packed
    return xb.to(dtype=torch.bfloat16, memory_format=torch.channels_last)


def _round_tf32(t: torch.Tensor) -> torch.Tensor:
    """fp32 -> nearest tf32 value (10 mantissa bits, ties away from zero like cvt.rna.tf32.f32), still stored as fp32; a new tensor
    with the same strides.  Bit arithmetic on the sign-magnitude pattern: add half an ulp of the kept mantissa, clear the rest."""
    return ((t.view(torch.int32) + 0x1000) & -0x2000).view(torch.float32)


def conv_integral_head(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], num_kp: int, num_hypo: int,
                       neighbor_size: int, return_logits: bool = False, precision: str = "bf16"):
    """The head's final `Conv2d(C, K*D, 1)` (deconv_head.py:33-35) fused with the integral multi-hypothesis head
    (…_multi.py:69-88), forward only (the eval path, eval.py:120): tcgen05 tensor-core GEMM with the softmax statistics
    taken from the TMEM accumulators, so the `[B, K*D, H, W]` logits are never written to HBM.

    x `[B, C, H, W]`: used in place when it is bf16 with channels-last storage (what `autocast` + `channels_last`
    backbones produce); anything else is converted once (one extra pass over the activations).
    weight `[K*D, C]` or `[K*D, C, 1, 1]`, bias `[K*D]` or None.  Operands are rounded to bf16, accumulation and all
    statistics are fp32 - the reference's own conv runs in TF32 on the same hardware.
    `precision="tf32"`: fp32 operands through `tcgen05.mma.kind::tf32` instead (x is taken as fp32 channels-last, converted once if
    it is not) - what the reference's cuDNN conv computes on this GPU; about twice the tensor-core time, forward / eval only.
    Returns (kps [B,NH,K,3], depth_prob_map [K,D], peak_idx [B,K,NH]) and, with `return_logits`, the fp32 logits."""
    cabi.require_cuda(x, "x")
    if precision not in ("bf16", "tf32"):
        raise ValueError("precision must be 'bf16' or 'tf32', got %r" % (precision,))
    B, C, H, W = x.shape
    w2 = weight.detach().reshape(weight.shape[0], -1)
    if w2.shape[1] != C:
        raise ValueError("weight is %s, expected [K*D, %d]" % (tuple(weight.shape), C))
    D = w2.shape[0] // num_kp
    if D * num_kp != w2.shape[0]:
        raise ValueError("weight rows %d are not a multiple of num_kp %d" % (w2.shape[0], num_kp))
    dev = x.device
    tf32 = precision == "tf32"
    if tf32:
        # the tensor core drops the low 13 mantissa bits of a tf32 operand; round to nearest first (what cvt.rna.tf32.f32 and
        # the library convolutions do), otherwise the truncation bias costs a bit of accuracy
        xb = _round_tf32(x.detach().to(dtype=torch.float32).contiguous(memory_format=torch.channels_last))
        wb = _round_tf32(w2.to(device=dev, dtype=torch.float32).contiguous())
    else:
        xb = _nhwc_bf16(x)
        wb = w2.to(device=dev, dtype=torch.bfloat16).contiguous()
    bf = bias.detach().to(device=dev, dtype=torch.float32).contiguous() if bias is not None else None
    shape = cabi.make_shape(B, num_kp, D, H, W, num_hypo, neighbor_size, torch.bfloat16, cabi.HEAD_MULTI)
    kps = torch.empty(B, num_hypo, num_kp, 3, dtype=torch.float32, device=dev)
    dmap = torch.empty(num_kp, D, dtype=torch.float32, device=dev)
    idx = torch.empty(B, num_kp, num_hypo, dtype=torch.int64, device=dev)
    stats = torch.empty(cabi.lib.xsup_stats_floats(shape), dtype=torch.float32, device=dev)
    logits = torch.empty(B, num_kp * D, H, W, dtype=torch.float32, device=dev) if return_logits else None
    fn, who = (cabi.lib.xsup_conv_head_fwd_tf32, "xsup_conv_head_fwd_tf32") if tf32 else (cabi.lib.xsup_conv_head_fwd, "xsup_conv_head_fwd")
    with torch.cuda.device(dev):
        cabi.check(fn(xb.data_ptr(), wb.data_ptr(), bf.data_ptr() if bf is not None else None,
                      kps.data_ptr(), dmap.data_ptr(), idx.data_ptr(), stats.data_ptr(),
                      logits.data_ptr() if return_logits else None, shape, C, cabi.stream_ptr(dev)), who)
    return (kps, dmap, idx, logits) if return_logits else (kps, dmap, idx)


class ConvIntegralHead(torch.autograd.Function):
    """Differentiable conv-fused head: `Conv2d(C, K*D, 1)` + integral multi-hypothesis head with NO logits tensor and NO
    d loss / d logits tensor in either direction.

    forward : `xsup_conv_head_fwd` (tcgen05 GEMM, softmax statistics from the TMEM accumulators).
    backward: `xsup_integral_coef` (g_kps + saved statistics -> per-unit coefficients), then `xsup_conv_head_bwd`: two
              tensor-core launches that each recompute the logit tiles, form d loss / d logits in the epilogue as a bf16
              shared-memory MMA operand and contract it on the spot - weight-stationary for d W (+ d bias), activation-
              stationary for d x.  No library GEMM, no intermediate in HBM.
    Gradients carry one bf16 rounding of d loss / d logits (relative 2^-9 per element), the same class of error as a
    bf16 autocast backward of the reference's conv; d W / d bias are accumulated with fp32 atomics over the samples."""

    @staticmethod
    def forward(ctx, x, weight, bias, num_kp, num_hypo, neighbor_size):
        cabi.require_cuda(x, "x")
        B, C, H, W = x.shape
        dev = x.device
        w2 = weight.detach().reshape(weight.shape[0], -1)
        D = w2.shape[0] // num_kp
        if w2.shape[1] != C or D * num_kp != w2.shape[0]:
            raise ValueError("weight is %s, expected [num_kp*D, %d(,1,1)]" % (tuple(weight.shape), C))
        xb = _nhwc_bf16(x)
        wb = w2.to(device=dev, dtype=torch.bfloat16).contiguous()
        bf = bias.detach().to(device=dev, dtype=torch.float32).contiguous() if bias is not None else None
        shape = cabi.make_shape(B, num_kp, D, H, W, num_hypo, neighbor_size, torch.bfloat16, cabi.HEAD_MULTI)
        kps = torch.empty(B, num_hypo, num_kp, 3, dtype=torch.float32, device=dev)
        dmap = torch.empty(num_kp, D, dtype=torch.float32, device=dev)
        idx = torch.empty(B, num_kp, num_hypo, dtype=torch.int64, device=dev)
        stats = torch.empty(cabi.lib.xsup_stats_floats(shape), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            cabi.check(cabi.lib.xsup_conv_head_fwd(xb.data_ptr(), wb.data_ptr(), bf.data_ptr() if bf is not None else None,
                                                   kps.data_ptr(), dmap.data_ptr(), idx.data_ptr(), stats.data_ptr(), None, shape, C,
                                                   cabi.stream_ptr(dev)), "xsup_conv_head_fwd")
        ctx.save_for_backward(xb, wb, bf, stats)
        ctx.shape, ctx.C = shape, C
        ctx.meta = (x.dtype, weight.dtype, tuple(weight.shape), bias.dtype if bias is not None else None,
                    x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous())
        ctx.mark_non_differentiable(dmap, idx)
        return kps, dmap, idx

    @staticmethod
    def backward(ctx, g_kps, _g_dmap, _g_idx):
        xb, wb, bf, stats = ctx.saved_tensors
        shape, C = ctx.shape, ctx.C
        x_dtype, w_dtype, w_shape, b_dtype, x_was_cl = ctx.meta
        dev = xb.device
        B, K, D, H, W = shape.B, shape.K, shape.D, shape.H, shape.W
        KD, HW = K * D, H * W
        g_kps = g_kps.to(torch.float32).contiguous()
        coef = torch.empty(cabi.lib.xsup_coef_floats(shape), dtype=torch.float32, device=dev)
        rowcoef = torch.empty(cabi.lib.xsup_conv_bwd_ws_floats(shape), dtype=torch.float32, device=dev)
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1] or (bf is not None and ctx.needs_input_grad[2])
        dx_f32 = x_dtype != torch.bfloat16                                     # fp32 input: do not round the result a second time
        # channels-last storage [B, H, W, C], returned as the logical NCHW view of it
        dx = torch.empty((B, C, H, W), dtype=torch.float32 if dx_f32 else torch.bfloat16, device=dev,
                         memory_format=torch.channels_last) if need_x else None
        dw = torch.empty(KD, C, dtype=torch.float32, device=dev) if need_w else None
        db = torch.empty(KD, dtype=torch.float32, device=dev) if (need_w and bf is not None) else None
        st = cabi.stream_ptr(dev)
        with torch.cuda.device(dev):
            cabi.check(cabi.lib.xsup_integral_coef(stats.data_ptr(), g_kps.data_ptr(), coef.data_ptr(), shape, st), "xsup_integral_coef")
            cabi.check(cabi.lib.xsup_conv_head_bwd(xb.data_ptr(), wb.data_ptr(), bf.data_ptr() if bf is not None else None, coef.data_ptr(),
                                                   rowcoef.data_ptr(), dx.data_ptr() if need_x else None, 1 if dx_f32 else 0,
                                                   dw.data_ptr() if need_w else None, db.data_ptr() if db is not None else None,
                                                   shape, C, st), "xsup_conv_head_bwd")
        g_x = g_w = g_b = None
        if need_x:
            g_x = dx.to(x_dtype)
            if not x_was_cl:
                g_x = g_x.contiguous()
        if ctx.needs_input_grad[1]:
            g_w = dw.to(w_dtype).reshape(w_shape)
        if bf is not None and ctx.needs_input_grad[2]:
            g_b = db.to(b_dtype)
        return g_x, g_w, g_b, None, None, None


def conv_integral_head_train(x, weight, bias, num_kp, num_hypo, neighbor_size):
    """Differentiable form of `conv_integral_head` (see `ConvIntegralHead`): -> (kps, depth_prob_map, peak_idx)."""
    return ConvIntegralHead.apply(x, weight, bias, num_kp, num_hypo, neighbor_size)


class ConvIntegralReprojMinLoss(torch.autograd.Function):
    """`Conv2d(C, K*D, 1)` -> integral multi-hypothesis head -> per-hypothesis world lift + loss terms + min over slots, forward
    and backward, with neither the logits nor d loss / d logits ever in HBM: `IntegralReprojMinLoss` with the head's final
    layer (deconv_head.py:33-35) pulled in.  Four launches forward+backward apart from the two tiny set-up kernels:
    `xsup_conv_head_fwd` (tcgen05), `xsup_reproj_fused_fwd`, `xsup_reproj_fused_bwd` (-> coefficient blocks),
    `xsup_conv_head_bwd` (tcgen05: d W + d bias launch, d x launch).

    Inputs as `ConvIntegralHead` (x `[B,C,H,W]`, weight `[K*D,C(,1,1)]`, bias) followed by the loss arguments of
    `IntegralReprojMinLoss`; outputs (loss_pseudo, loss_sym, sel_idx, kps, kps_world, depth_prob_map, peak_idx).
    Gradients flow to x, weight and bias from both losses and from anything downstream of `kps` / `kps_world`."""

    @staticmethod
    def forward(ctx, x, weight, bias, target, trans_image, pelvis, k_mat, trans_world, rot_world, num_kp, num_hypo, neighbor_size,
                img_hw, rect_width, w_mse, w_bone, w_kp, w_kp2d, reduction, group):
        cabi.require_cuda(x, "x")
        ctx.set_materialize_grads(False)
        B, C, H, W = x.shape
        dev = x.device
        if B == 0:
            raise ValueError("ConvIntegralReprojMinLoss needs a non-empty batch")
        w2 = weight.detach().reshape(weight.shape[0], -1)
        D = w2.shape[0] // num_kp
        if w2.shape[1] != C or D * num_kp != w2.shape[0]:
            raise ValueError("weight is %s, expected [num_kp*D, %d(,1,1)]" % (tuple(weight.shape), C))
        K, NH = num_kp, num_hypo
        xb = _nhwc_bf16(x)
        wb = w2.to(device=dev, dtype=torch.bfloat16).contiguous()
        bf = bias.detach().to(device=dev, dtype=torch.float32).contiguous() if bias is not None else None
        shape = cabi.make_shape(B, K, D, H, W, NH, neighbor_size, torch.bfloat16, cabi.HEAD_MULTI)
        kps = torch.empty(B, NH, K, 3, dtype=torch.float32, device=dev)
        dmap = torch.empty(K, D, dtype=torch.float32, device=dev)
        idx = torch.empty(B, K, NH, dtype=torch.int64, device=dev)
        stats = torch.empty(cabi.lib.xsup_stats_floats(shape), dtype=torch.float32, device=dev)
        target = target.to(device=dev, dtype=torch.float32).contiguous()
        if tuple(target.shape) != (B, K, 3):
            raise ValueError("target must be [B,K,3] = %s, got %s" % ((B, K, 3), tuple(target.shape)))
        keep, cam = _cam_args(dict(trans_image=trans_image, pelvis=pelvis, k_mat=k_mat, trans_world=trans_world,
                                   rot_world=rot_world), B)
        if xdist.is_active(group) and not isinstance(group, xdist.PeerExchange):
            raise ValueError("ConvIntegralReprojMinLoss takes group=None or a dist.PeerExchange (the in-kernel exchange)")
        use_sym = any(w is not None for w in (w_bone, w_kp, w_kp2d))
        cfg = cabi.LossCfg(B, K, NH, int(img_hw[0]), int(img_hw[1]), float(rect_width), float(w_mse),
                           float(w_bone or 0.0), float(w_kp or 0.0), float(w_kp2d or 0.0), int(use_sym),
                           cabi.REDUCE[reduction], xdist.global_batch(B, group))
        world = torch.empty_like(kps)
        sample_terms = torch.empty(B, cabi.LOSS_TERMS, NH, dtype=torch.float32, device=dev)
        partial = torch.empty(cabi.LOSS_TERMS, NH, dtype=torch.float32, device=dev)
        loss = torch.empty(2, dtype=torch.float32, device=dev)
        sel = torch.empty({"batch": (2,), "sample": (2, B), "joint": (B, K)}[reduction], dtype=torch.int64, device=dev)
        st = cabi.stream_ptr(dev)
        with torch.cuda.device(dev):
            with _range("xsup.K7.conv_head_fwd"):
                cabi.check(cabi.lib.xsup_conv_head_fwd(xb.data_ptr(), wb.data_ptr(), bf.data_ptr() if bf is not None else None,
                                                       kps.data_ptr(), dmap.data_ptr(), idx.data_ptr(), stats.data_ptr(), None, shape, C, st),
                           "xsup_conv_head_fwd")
            with _range("xsup.K2.loss_select_fwd"):
                xchg = group.descriptor() if xdist.is_active(group) else None
                ticket = stats.data_ptr() + 4 * (B * K * int(cabi.lib.xsup_stats_stride(shape)) + 1)
                cabi.check(cabi.lib.xsup_reproj_fused_fwd(kps.data_ptr(), target.data_ptr(), cam, world.data_ptr(), sample_terms.data_ptr(),
                                                          partial.data_ptr(), loss.data_ptr(), sel.data_ptr(), cfg, xchg, ticket, st),
                           "xsup_reproj_fused_fwd")
        ctx.save_for_backward(xb, wb, bf, stats, kps, target, sel, *keep)
        ctx.shape, ctx.cfg, ctx.C = shape, cfg, C
        ctx.meta = (x.dtype, weight.dtype, tuple(weight.shape), bias.dtype if bias is not None else None,
                    x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous())
        ctx.mark_non_differentiable(sel, dmap, idx)
        return loss[0], loss[1], sel, kps, world, dmap, idx

    @staticmethod
    def backward(ctx, g_lp, g_ls, _g_sel, g_kps_out, g_world, _g_dmap, _g_idx):
        xb, wb, bf, stats, kps, target, sel, *keep = ctx.saved_tensors
        shape, cfg, C = ctx.shape, ctx.cfg, ctx.C
        x_dtype, w_dtype, w_shape, b_dtype, x_was_cl = ctx.meta
        if g_lp is None and g_ls is None and g_kps_out is None and g_world is None:
            return (None,) * 20
        dev = xb.device
        B, K, D, H, W = shape.B, shape.K, shape.D, shape.H, shape.W
        cam = cabi.make_cam(*keep, B)

        def ptr(t):
            if t is None:
                return None, None
            t = t.to(torch.float32).contiguous()
            return t, t.data_ptr()
        g_lp, p_lp = ptr(g_lp)
        g_ls, p_ls = ptr(g_ls)
        g_kps_out, p_gk = ptr(g_kps_out)
        g_world, p_gw = ptr(g_world)
        coef = torch.empty(cabi.lib.xsup_coef_floats(shape), dtype=torch.float32, device=dev)
        rowcoef = torch.empty(cabi.lib.xsup_conv_bwd_ws_floats(shape), dtype=torch.float32, device=dev)
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1] or (bf is not None and ctx.needs_input_grad[2])
        dx_f32 = x_dtype != torch.bfloat16
        dx = torch.empty((B, C, H, W), dtype=torch.float32 if dx_f32 else torch.bfloat16, device=dev,
                         memory_format=torch.channels_last) if need_x else None
        dw = torch.empty(K * D, C, dtype=torch.float32, device=dev) if need_w else None
        db = torch.empty(K * D, dtype=torch.float32, device=dev) if (need_w and bf is not None) else None
        st = cabi.stream_ptr(dev)
        with torch.cuda.device(dev):
            with _range("xsup.K2.loss_bwd_coef"):
                cabi.check(cabi.lib.xsup_reproj_fused_bwd(kps.data_ptr(), target.data_ptr(), cam, sel.data_ptr(), p_lp, p_ls, p_gk, p_gw,
                                                          stats.data_ptr(), coef.data_ptr(), None, cfg, shape, st), "xsup_reproj_fused_bwd")
            with _range("xsup.K8.conv_head_bwd"):
                cabi.check(cabi.lib.xsup_conv_head_bwd(xb.data_ptr(), wb.data_ptr(), bf.data_ptr() if bf is not None else None, coef.data_ptr(),
                                                       rowcoef.data_ptr(), dx.data_ptr() if need_x else None, 1 if dx_f32 else 0,
                                                       dw.data_ptr() if need_w else None, db.data_ptr() if db is not None else None,
                                                       shape, C, st), "xsup_conv_head_bwd")
        g_x = g_w = g_b = None
        if need_x:
            g_x = dx.to(x_dtype)
            if not x_was_cl:
                g_x = g_x.contiguous()
        if ctx.needs_input_grad[1]:
            g_w = dw.to(w_dtype).reshape(w_shape)
        if bf is not None and ctx.needs_input_grad[2]:
            g_b = db.to(b_dtype)
        return (g_x, g_w, g_b) + (None,) * 17


def conv_integral_reproj_min_loss(x, weight, bias, target, cams: Dict[str, torch.Tensor], num_kp, num_hypo, neighbor_size,
                                  img_hw: Sequence[int] = (256, 256), rect_width: float = 2000.0, w_mse: float = 1.0,
                                  w_bone: Optional[float] = None, w_kp: Optional[float] = None, w_kp2d: Optional[float] = None,
                                  reduction: str = "batch", group=None):
    """Functional form of `ConvIntegralReprojMinLoss` (camera tensors as a dict, like `integral_reproj_min_loss`)."""
    return ConvIntegralReprojMinLoss.apply(x, weight, bias, target, cams["trans_image"], cams["pelvis"], cams["k_mat"],
                                           cams["trans_world"], cams["rot_world"], num_kp, num_hypo, neighbor_size, tuple(img_hw),
                                           rect_width, w_mse, w_bone, w_kp, w_kp2d, reduction, group)
