"""Drop-in modules for the reference's detectors.

`KPDetector3DMulti` keeps the constructor of modules/keypoint_detector_integral_multi.py:10
(`name, num_kp, depth_dim, num_hypo, neighbor_size, num_layers=50` — the yaml's `detector_params`
are splatted into it, train.py:215), the attribute `self.net` (state_dict keys `net.backbone.*`,
`net.head.*`, so checkpoints load unchanged: train.py:127, eval.py:312) and the methods
`forward`, `find_peak`, `generate_3d_integral_preds_tensor`.  Only the parameter-free tail after
`self.net(x)` is replaced — by the CUDA kernels.

The backbone (ResNet + deconv head, modules/integral_base_modules/*) is out of scope: pass it as
`net=`; when omitted and the reference's `modules.integral_base_modules.network` is importable
(i.e. this package is used from inside the reference checkout) it is built exactly as the
reference does.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops

__all__ = ["KPDetector3DMulti", "KPDetector3D"]


def _reference_backbone(depth_dim, num_layers, num_kp):
    try:
        from modules.integral_base_modules.network import get_default_network_config, get_pose_net
    except Exception as e:  # pragma: no cover - depends on the caller's checkout
        raise RuntimeError(
            "no backbone given: pass `net=` (any module producing [B, num_kp*depth_dim, H, W] logits), or "
            "import this package from inside the X-as-Supervision checkout so that "
            "modules.integral_base_modules.network is importable (%s)" % (e,))
    cfg = get_default_network_config()
    cfg.depth_dim = depth_dim
    cfg.num_layers = num_layers
    return get_pose_net(cfg, num_joints=num_kp)


class KPDetector3DMulti(nn.Module):
    def __init__(self, name, num_kp, depth_dim, num_hypo, neighbor_size, num_layers=50, net=None):
        super().__init__()
        self.num_hypo = num_hypo
        self.neighbor_size = neighbor_size
        self.num_kp = num_kp
        self.depth_dim = depth_dim
        self.net = net if net is not None else _reference_backbone(depth_dim, num_layers, num_kp)
        self.name = name

    def find_peak(self, heatmap):
        """`[B,K,D]` depth marginal -> int64 `[B,K,NH]` (…_multi.py:24-34)."""
        return ops.find_peak(heatmap, self.num_hypo)

    def generate_3d_integral_preds_tensor(self, heatmaps, x_dim, y_dim, z_dim):
        """Probabilities `[B,K,D,H,W]` -> (x `[B,K,1]`, y `[B,K,1]`, z `[B,K,NH]`, depth_prob_map `[K,D]`)
        in bin units, as …_multi.py:36-64.  `softmax(log p) == p` for a normalised p, so this runs the
        same forward kernel on `log p` and undoes the [-1,1] normalisation of its output."""
        B, K, D, H, W = heatmaps.shape
        logits = torch.log(heatmaps).reshape(B, K * D, H, W)
        kps, dmap, _ = ops.integral_multi_head(logits, K, self.num_hypo, self.neighbor_size)
        x = (kps[:, 0, :, 0:1] + 1) / 2 * H          # forward() divides x by H and y by W (…:78-79)
        y = (kps[:, 0, :, 1:2] + 1) / 2 * W
        z = (kps[..., 2].permute(0, 2, 1) + 1) / 2 * D
        return x, y, z, dmap

    def forward(self, x):
        heatmap = self.net(x)
        kps, depth_prob_map, _ = ops.integral_multi_head(heatmap, self.num_kp, self.num_hypo, self.neighbor_size)
        return kps, depth_prob_map

    @torch.no_grad()
    def forward_fused(self, x, precision="bf16"):
        """Inference path (eval.py:120) with the final `Conv2d(C, K*D, 1)` of `self.net.head`
        (deconv_head.py:33-35) fused into the head tail on the tensor cores: the `[B, K*D, H, W]` logits are never
        materialised.  Needs the reference's network layout (`net.backbone`, `net.head.features[-1]` a 1x1 conv with
        bias); returns the same `(kps, depth_prob_map)` as `forward`, with the conv operands rounded to bf16, or - with
        `precision="tf32"` - to tf32, which is what the reference's own conv computes on this GPU."""
        feats, last = _split_final_conv(self.net)
        y = feats(x)
        kps, depth_prob_map, _ = ops.conv_integral_head(y, last.weight, last.bias, self.num_kp, self.num_hypo, self.neighbor_size,
                                                        precision=precision)
        return kps, depth_prob_map


def _split_final_conv(net):
    """(callable producing the input of the final 1x1 conv, that conv) for a ResPoseNet-shaped `net`
    (modules/integral_base_modules/network.py:10-19, deconv_head.py:22-35)."""
    head = getattr(net, "head", None)
    layers = getattr(head, "features", None)
    if layers is None or len(layers) == 0 or not isinstance(layers[-1], nn.Conv2d) or tuple(layers[-1].kernel_size) != (1, 1):
        raise RuntimeError("forward_fused needs net.head.features to end with a 1x1 nn.Conv2d (the reference's DeconvHead "
                           "with conv_kernel_size=1 and with_bias_end=True)")

    def feats(x):
        x = net.backbone(x)
        for layer in list(layers)[:-1]:
            x = layer(x)
        return x
    return feats, layers[-1]


class KPDetector3D(nn.Module):
    """modules/keypoint_detector_integral.py:6 — single hypothesis, output `[B,1,K,3]`."""

    def __init__(self, name, num_kp, depth_dim, num_layers=50, net=None):
        super().__init__()
        self.num_kp = num_kp
        self.depth_dim = depth_dim
        self.net = net if net is not None else _reference_backbone(depth_dim, num_layers, num_kp)
        self.name = name

    def forward(self, x):
        heatmap = self.net(x)
        return ops.integral_single_head(heatmap, self.num_kp)
