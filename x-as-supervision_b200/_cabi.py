"""ctypes binding of libxsup_b200.so (the C ABI declared in include/xsup_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing the
import of this module raises, and every compute call raises `RuntimeError`
with the library's own message when a launch is rejected or fails.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libxsup_b200.so")

F32, BF16 = 0, 1
HEAD_MULTI, HEAD_SINGLE = 0, 1
REDUCE = {"batch": 0, "sample": 1, "joint": 2}
LOSS_TERMS = 4
FLAG_NORM, FLAG_MONO, FLAG_PATCH = 1, 2, 4

# every symbol include/xsup_b200.h declares (tests check the library exports all of them)
SYMBOLS = (
    "xsup_abi_version", "xsup_last_error", "xsup_launch_count", "xsup_stats_stride", "xsup_coef_stride",
    "xsup_stats_floats", "xsup_coef_floats",
    "xsup_integral_fwd", "xsup_integral_bwd", "xsup_find_peak",
    "xsup_patch_to_world_fwd", "xsup_patch_to_world_bwd", "xsup_world_to_patch_fwd",
    "xsup_reproj_loss_fwd", "xsup_reproj_select", "xsup_reproj_loss_bwd",
    "xsup_xchg_floats", "xsup_partial_allreduce",
    "xsup_skel_ws_floats", "xsup_draw_lines_ws_floats", "xsup_mask_loss_ws_floats",
    "xsup_draw_lines_fwd", "xsup_draw_lines_bwd", "xsup_skeleton_mask_fwd", "xsup_skeleton_mask_bwd",
    "xsup_mask_loss_fwd", "xsup_mask_loss_bwd",
    "xsup_eval_select", "xsup_triangulate", "xsup_root_centre_fwd", "xsup_root_centre_bwd",
    "xsup_disc_min_loss_fwd", "xsup_disc_min_loss_bwd",
    "xsup_conv_head_fwd", "xsup_pack_nhwc_bf16", "xsup_integral_coef", "xsup_conv_head_bwd_g",
    "xsup_pose_term_fwd", "xsup_pose_term_bwd",
    # ABI v10
    "xsup_geom_patch_to_world", "xsup_geom_patch_to_world_vjp", "xsup_geom_world_to_patch", "xsup_geom_world_to_patch_vjp",
    "xsup_reproj_fused_fwd", "xsup_reproj_fused_bwd", "xsup_integral_bwd_apply", "xsup_pose_sqerr",
    # ABI v11
    "xsup_conv_bwd_ws_floats", "xsup_conv_head_bwd", "xsup_conv_head_fwd_tf32",
)
GEOM_NORM, GEOM_MONO, GEOM_PATCH_STAGE, GEOM_CAMERA_STAGE = 1, 2, 4, 8
SCHED_WORDS = 16
TERM_MSE, TERM_BONE, TERM_KP = 0, 1, 2
MAX_VIEWS = 8
MAX_LINES = 32
MASK_MSE, MASK_CLIP_MEAN, MASK_WEIGHTED = 0, 1, 2
MASK_SUMS = 4
XCHG_SLOT = 1024      # floats per (parity, source rank) mailbox slot: XSUP_XCHG_SLOT


class Shape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "K", "D", "H", "W", "NH", "NS", "dtype", "head")]


class Cam(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("trans_image", "pelvis", "k_mat", "trans_world", "rot_world")]


class LossCfg(C.Structure):
    _fields_ = [("B", C.c_int32), ("K", C.c_int32), ("NH", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32),
                ("rect_width", C.c_float), ("w_mse", C.c_float), ("w_bone", C.c_float), ("w_kp", C.c_float),
                ("w_kp2d", C.c_float), ("use_sym", C.c_int32), ("reduction", C.c_int32), ("batch_total", C.c_int32)]


class Skel(C.Structure):
    _fields_ = [("B", C.c_int32), ("K", C.c_int32), ("S", C.c_int32), ("L", C.c_int32), ("kp_batch_stride", C.c_int32),
                ("kp_joint_stride", C.c_int32), ("body_width", C.c_float), ("parent", C.c_int32 * MAX_LINES),
                ("child", C.c_int32 * MAX_LINES)]


class MaskLoss(C.Structure):
    _fields_ = [("n", C.c_int64), ("mode", C.c_int32), ("use_clip", C.c_int32)]


class Eval(C.Structure):
    _fields_ = [("B", C.c_int32), ("NH", C.c_int32), ("K", C.c_int32), ("img_size", C.c_float), ("best", C.c_int32),
                ("perm", C.c_int32 * 32)]


class Tri(C.Structure):
    _fields_ = [("V", C.c_int32), ("B", C.c_int32), ("K", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32),
                ("is_norm", C.c_int32), ("rect_width", C.c_float), ("kps", C.c_void_p * MAX_VIEWS), ("cam", Cam * MAX_VIEWS)]


class Xchg(C.Structure):
    _fields_ = [("peer_bufs", C.c_void_p), ("rank", C.c_int32), ("world", C.c_int32), ("step", C.c_uint32), ("seq", C.c_void_p),
                ("err", C.c_void_p)]


class Geom(C.Structure):
    _fields_ = [("B", C.c_int32), ("J", C.c_int32), ("img_d", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32),
                ("depth_scale", C.c_float), ("flags", C.c_int32), ("intr_stride", C.c_int32),
                ("trans_image", C.c_void_p), ("pelvis", C.c_void_p), ("fx", C.c_void_p), ("fy", C.c_void_p), ("cx", C.c_void_p),
                ("cy", C.c_void_p), ("trans_world", C.c_void_p), ("rot_world", C.c_void_p)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libxsup_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "from the repository root; there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
    lib.xsup_abi_version.restype = C.c_int
    lib.xsup_last_error.restype = C.c_char_p
    lib.xsup_launch_count.restype = C.c_uint64
    lib.xsup_stats_stride.restype = C.c_size_t
    lib.xsup_stats_stride.argtypes = [C.POINTER(Shape)]
    lib.xsup_coef_stride.restype = C.c_size_t
    lib.xsup_coef_stride.argtypes = [C.POINTER(Shape)]
    lib.xsup_stats_floats.restype = C.c_size_t
    lib.xsup_stats_floats.argtypes = [C.POINTER(Shape)]
    lib.xsup_coef_floats.restype = C.c_size_t
    lib.xsup_coef_floats.argtypes = [C.POINTER(Shape)]
    lib.xsup_integral_fwd.argtypes = [vp, vp, vp, vp, vp, C.POINTER(Shape), vp]
    lib.xsup_integral_bwd.argtypes = [vp, vp, vp, vp, vp, C.POINTER(Shape), vp]
    lib.xsup_find_peak.argtypes = [vp, vp, i32, i32, i32, vp]
    lib.xsup_patch_to_world_fwd.argtypes = [vp, C.POINTER(Cam), vp, i32, i32, i32, i32, f32, i32, vp]
    lib.xsup_patch_to_world_bwd.argtypes = [vp, vp, C.POINTER(Cam), vp, i32, i32, i32, i32, f32, i32, vp]
    lib.xsup_world_to_patch_fwd.argtypes = [vp, C.POINTER(Cam), vp, i32, i32, i32, i32, f32, i32, vp]
    lib.xsup_reproj_loss_fwd.argtypes = [vp, vp, C.POINTER(Cam), vp, vp, vp, C.POINTER(LossCfg), vp]
    lib.xsup_reproj_select.argtypes = [vp, vp, vp, vp, vp, vp, C.POINTER(LossCfg), vp]
    lib.xsup_reproj_loss_bwd.argtypes = [vp, vp, C.POINTER(Cam), vp, vp, vp, C.POINTER(LossCfg), vp]
    lib.xsup_xchg_floats.restype = C.c_size_t
    lib.xsup_xchg_floats.argtypes = [i32]
    lib.xsup_partial_allreduce.argtypes = [vp, i32, C.POINTER(Xchg), vp]
    lib.xsup_partial_allreduce.restype = C.c_int
    for name in ("xsup_skel_ws_floats", "xsup_draw_lines_ws_floats"):
        getattr(lib, name).restype = C.c_size_t
        getattr(lib, name).argtypes = [C.POINTER(Skel)]
    lib.xsup_mask_loss_ws_floats.restype = C.c_size_t
    lib.xsup_mask_loss_ws_floats.argtypes = [C.c_int64]
    sk, ml = C.POINTER(Skel), C.POINTER(MaskLoss)
    lib.xsup_draw_lines_fwd.argtypes = [vp, sk, vp, vp]
    lib.xsup_draw_lines_bwd.argtypes = [vp, sk, vp, vp, vp, vp, vp]
    lib.xsup_skeleton_mask_fwd.argtypes = [vp, sk, vp, vp, vp, vp, ml, vp, vp, vp]
    lib.xsup_skeleton_mask_bwd.argtypes = [vp, sk, vp, vp, vp, vp, vp, ml, vp, vp, vp, vp, vp]
    lib.xsup_mask_loss_fwd.argtypes = [vp, vp, vp, vp, ml, vp, vp, vp]
    lib.xsup_mask_loss_bwd.argtypes = [vp, vp, vp, ml, vp, vp, vp, vp]
    lib.xsup_conv_head_fwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(Shape), i32, vp]
    lib.xsup_conv_head_fwd.restype = C.c_int
    lib.xsup_conv_head_fwd_tf32.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(Shape), i32, vp]
    lib.xsup_conv_head_fwd_tf32.restype = C.c_int
    lib.xsup_pack_nhwc_bf16.argtypes = [vp, vp, i32, i32, i32, vp]
    lib.xsup_pack_nhwc_bf16.restype = C.c_int
    lib.xsup_pose_term_fwd.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.xsup_pose_term_fwd.restype = C.c_int
    lib.xsup_pose_term_bwd.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.xsup_pose_term_bwd.restype = C.c_int
    lib.xsup_integral_coef.argtypes = [vp, vp, vp, C.POINTER(Shape), vp]
    lib.xsup_integral_coef.restype = C.c_int
    lib.xsup_conv_head_bwd_g.argtypes = [vp, vp, vp, vp, vp, vp, C.POINTER(Shape), i32, vp]
    lib.xsup_conv_head_bwd_g.restype = C.c_int
    lib.xsup_conv_bwd_ws_floats.restype = C.c_size_t
    lib.xsup_conv_bwd_ws_floats.argtypes = [C.POINTER(Shape)]
    lib.xsup_conv_head_bwd.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, vp, C.POINTER(Shape), i32, vp]
    lib.xsup_conv_head_bwd.restype = C.c_int
    gp = C.POINTER(Geom)
    lib.xsup_geom_patch_to_world.argtypes = [vp, vp, gp, vp]
    lib.xsup_geom_patch_to_world_vjp.argtypes = [vp, vp, vp, gp, vp]
    lib.xsup_geom_world_to_patch.argtypes = [vp, vp, gp, vp]
    lib.xsup_geom_world_to_patch_vjp.argtypes = [vp, vp, vp, gp, vp]
    lib.xsup_reproj_fused_fwd.argtypes = [vp, vp, C.POINTER(Cam), vp, vp, vp, vp, vp, C.POINTER(LossCfg), C.POINTER(Xchg), vp, vp]
    lib.xsup_reproj_fused_bwd.argtypes = [vp, vp, C.POINTER(Cam), vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(LossCfg),
                                          C.POINTER(Shape), vp]
    lib.xsup_integral_bwd_apply.argtypes = [vp, vp, vp, C.POINTER(Shape), vp]
    lib.xsup_pose_sqerr.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp]
    for name in ("xsup_geom_patch_to_world", "xsup_geom_patch_to_world_vjp", "xsup_geom_world_to_patch",
                 "xsup_geom_world_to_patch_vjp", "xsup_reproj_fused_fwd", "xsup_reproj_fused_bwd", "xsup_integral_bwd_apply",
                 "xsup_pose_sqerr"):
        getattr(lib, name).restype = C.c_int
    lib.xsup_eval_select.argtypes = [vp, vp, C.POINTER(Eval), vp, vp, vp, vp, vp, vp, vp, vp]
    lib.xsup_triangulate.argtypes = [C.POINTER(Tri), vp, vp]
    lib.xsup_root_centre_fwd.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.xsup_root_centre_bwd.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.xsup_disc_min_loss_fwd.argtypes = [vp, i32, i32, i32, f32, vp, vp, vp]
    lib.xsup_disc_min_loss_bwd.argtypes = [vp, vp, vp, i32, i32, i32, f32, vp, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name.startswith(("xsup_eval", "xsup_triangulate", "xsup_root", "xsup_disc")):
            fn.restype = C.c_int
        if name.startswith(("xsup_integral", "xsup_find", "xsup_patch", "xsup_world", "xsup_reproj", "xsup_draw_lines_f",
                            "xsup_draw_lines_b", "xsup_skeleton", "xsup_mask_loss_f", "xsup_mask_loss_b")):
            fn.restype = C.c_int
    return lib


lib = _load()
ABI_VERSION = 11
if lib.xsup_abi_version() != ABI_VERSION:
    raise ImportError("libxsup_b200.so ABI version %d, expected %d: rebuild with __graft_entry__.build()"
                      % (lib.xsup_abi_version(), ABI_VERSION))


def check(rc: int, who: str) -> None:
    if rc != 0:
        raise RuntimeError("%s failed (code %d): %s" % (who, rc, lib.xsup_last_error().decode()))


def launch_count() -> int:
    return int(lib.xsup_launch_count())


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: xsup_b200 has no CPU path (got device %s)" % (name, t.device))


def make_shape(B, K, D, H, W, NH, NS, dtype, head=HEAD_MULTI) -> Shape:
    if dtype == torch.float32:
        dt = F32
    elif dtype == torch.bfloat16:
        dt = BF16
    else:
        raise TypeError("logits must be float32 or bfloat16, got %s" % dtype)
    return Shape(B, K, D, H, W, NH, NS, dt, head)


def make_cam(trans_image, pelvis, k_mat, trans_world, rot_world, B) -> Cam:
    shapes = ((trans_image, (B, 2, 3)), (pelvis, (B, 3)), (k_mat, (B, 3, 3)), (trans_world, (B, 3)), (rot_world, (B, 3, 3)))
    ptrs = []
    for t, shp in shapes:
        require_cuda(t, "camera tensor")
        if tuple(t.shape) != shp or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("camera tensor must be contiguous float32 of shape %s, got %s %s" % (shp, tuple(t.shape), t.dtype))
        ptrs.append(t.data_ptr())
    return Cam(*ptrs)
