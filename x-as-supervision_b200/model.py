"""`Counter3DModel` / `Counter3DDisc` with the reference's constructor and `forward` signatures
(modules/model.py:24-192, 194-266), assembled from the CUDA-backed ops of this package.

This is the *caller* of the hot path, kept here as the integration proof: the per-camera Python loops over
hypotheses (model.py:71-79, 105-114, 126-129, 158-162) become one fused call per image
(`ops.integral_reproj_min_loss`), `draw_lines` + `torch.max` + `compute_mask_reconstruction_loss` become
`skeleton.skeleton_mask_loss`, and the NH discriminator calls become one batched call on `evalops.root_centre`.
`tests/test_gpu_model.py` checks `loss_values` and the heat-map gradients against the reference's own
`Counter3DModel.forward` (golden: `tests/golden/model_*.npz`).

Differences, all deliberate:
  * `regressor` must expose `.net`, `.num_kp`, `.num_hypo`, `.neighbor_size` (our `detector.KPDetector3DMulti`, or the
    reference's): the head tail is run fused with the loss, so `regressor.forward` itself is not called;
  * `use_aug` (random z-rotation of the discriminator input, model.py:133-141) is the caller's data augmentation and is
    not implemented: `NotImplementedError`;
  * the `output` dict carries the same keys with the same values, except that tensors the reference only produces
    for TensorBoard are detached.
"""
from __future__ import annotations

import torch

from . import evalops, ops, skeleton

__all__ = ["Counter3DModel", "Counter3DDisc", "GraphedLossGraph", "cal_links"]

cal_links = skeleton.cal_links
_CAM_FIELDS = ("trans_image", "pelvis", "k_mat", "trans_world", "rot_world")


def _cams(x, cam_key):
    return {k: x["{}_{}".format(cam_key, k)] for k in _CAM_FIELDS}


def _batch_dependent(module: torch.nn.Module) -> bool:
    """True when a forward pass of `module` depends on which samples share the batch or draws random numbers: batch
    normalisation (gcn.py:58-75 with cfg['use_bn']) or dropout, in training mode."""
    stateful = (torch.nn.modules.batchnorm._BatchNorm, torch.nn.modules.dropout._DropoutNd)
    return any(isinstance(m, stateful) and m.training for m in module.modules())


def score_hypotheses(discriminator: torch.nn.Module, joints: torch.Tensor) -> torch.Tensor:
    """Discriminator logits of all hypotheses, `[B,NH,K,dim] -> [B,NH,C]` (model.py:126-129, 251-254).

    A stateless discriminator scores the NH hypotheses in ONE call on `[B*NH,K,dim]`.  With batch normalisation or
    dropout active the batched call would pool the batch statistics over hypotheses (and update the running statistics
    once instead of NH times), so the hypotheses go through one by one exactly as the reference loops."""
    B, NH = joints.shape[:2]
    if _batch_dependent(discriminator):
        return torch.stack([discriminator(joints[:, i].contiguous()) for i in range(NH)], dim=1)
    return discriminator(joints.flatten(0, 1).contiguous()).reshape(B, NH, -1)


class Counter3DModel(torch.nn.Module):
    def __init__(self, cfg, regressor, smpl_layer, h36m_regressor, physique_network=None):
        super().__init__()
        self.regressor = regressor
        self.cam_id_list = cfg["cam_id_list"]
        self.body_width = float(cfg["body_width"] if "body_width" in cfg else 3.0) * 1e-3
        line_select_ids = cfg["line_select_ids"] if "line_select_ids" in cfg else None
        self.parent_ids, self.child_ids = cal_links(cfg["parent_ids"], line_select_ids=line_select_ids, use_root=False, extension=True)
        self.loss_config = cfg["loss_config"]
        self.use_learned_width = cfg["use_learned_width"] if "use_learned_width" in cfg else False
        self.smpl_layer = smpl_layer
        self.h36m_regressor = h36m_regressor
        self.physique_network = physique_network
        self.DISC_SUP_DIMENSION = cfg["smpl_disc_params"]["disc_sup_dim"] if "disc_sup_dim" in cfg["smpl_disc_params"] else 3
        self.use_aug = cfg["smpl_disc_params"]["use_aug"] if "use_aug" in cfg["smpl_disc_params"] else False

    # ------------------------------------------------------------------------------------------------------------
    def _head(self, img):
        r = self.regressor
        return r.net(img), r.num_kp, r.num_hypo, r.neighbor_size

    def forward(self, x, smpl_discriminator):
        if self.use_aug:
            raise NotImplementedError("use_aug (random_rotation_3D of the discriminator input) is the caller's augmentation")
        mono = "cam_mono_img" in x
        cam_id_list = ["mono"] if mono else self.cam_id_list
        cfg = self.loss_config
        loss_values, output = {}, {}
        sym_w = cfg["symmetry_loss"]["weight"] if "symmetry_loss" in cfg else None

        kps_ori, kps_world_ori = {}, {}
        loss_sym = 0
        for cam_id in cam_id_list:
            cam_key = "cam_{}".format(cam_id)
            logits, K, NH, NS = self._head(x["{}_img".format(cam_key)])
            img_hw = tuple(x["{}_img".format(cam_key)].shape[-2:])
            if mono:
                # single-view data: head alone, then the mono lift per hypothesis (model.py:73-75); no symmetry term (:101-102)
                kps, depth_map, _ = ops.integral_multi_head(logits, K, NH, NS)
                world = ops.convert_patch_to_world(kps.reshape(kps.shape[0], NH * K, 3), x, cam_key, is_norm=True, RECT_WIDTH=256,
                                                   mono=True, patch=False).reshape(kps.shape)
            else:
                # head + world lift of every hypothesis + symmetry terms + min over hypotheses: one fused op (model.py:64,71-79,105-114)
                zeros = torch.zeros(logits.shape[0], K, 3, device=logits.device)
                _, ls, _, kps, world, depth_map, _ = ops.integral_reproj_min_loss(
                    logits, zeros, _cams(x, cam_key), K, NH, NS, img_hw=img_hw, rect_width=2000.0, w_mse=0.0,
                    w_bone=sym_w["bone"] if sym_w is not None else None, w_kp=sym_w["kp"] if sym_w is not None else None,
                    w_kp2d=sym_w["kp_2d"] if sym_w is not None and "kp_2d" in sym_w else None, reduction="batch", group=None)
                loss_sym = loss_sym + ls
            kps_ori[cam_key], kps_world_ori[cam_key] = kps, world
            output["pose_2d_pred_{}_ori".format(cam_key)] = kps.detach()[0:1, 0, ...].clone()
            output["depth_map_{}".format(cam_key)] = depth_map
            output["pose_3d_depth_{}".format(cam_key)] = world.detach()[:, 0, ...].clone()

        if not mono:
            output["kp_gt_world"] = ops.convert_patch_to_world(x["cam_0_joints"], x, "cam_0", is_norm=False)[0:1, ...]

        # skeleton mask of hypothesis 0 (model.py:88-96), fused with the reconstruction loss when that is configured (:181-190)
        reconstructed, loss_rec = {}, 0
        for cam_id in cam_id_list:
            cam_key = "cam_{}".format(cam_id)
            size = x["{}_img".format(cam_key)].shape[-1]
            if "recons_loss" in cfg:
                weight = x["{}_geodesic_dis".format(cam_key)] if cfg["recons_loss"]["use_dis_map"] else None
                rec, lr = skeleton.skeleton_mask_loss(kps_ori[cam_key][:, 0, :, :2], x["{}_mask".format(cam_key)], weight, size,
                                                      self.parent_ids, self.child_ids, self.body_width, use_clip=True)
                loss_rec = loss_rec + lr
            else:
                rec = skeleton.skeleton_mask(kps_ori[cam_key][:, 0, :, :2], size, self.parent_ids, self.child_ids, self.body_width)
            reconstructed[cam_key] = rec
            output["mask_heatmap_line_{}".format(cam_key)] = rec.detach()

        if sym_w is not None:
            loss_values["symmetry"] = loss_sym

        if "smpl_gen_loss" in cfg:
            loss_gen = 0
            for cam_id in cam_id_list:
                cam_key = "cam_{}".format(cam_id)
                # (w - w[:, [0], :]) / 1000 exactly as model.py:124 writes it, then ONE discriminator call for all hypotheses
                joints = evalops.root_centre(kps_world_ori[cam_key], self.DISC_SUP_DIMENSION).detach()
                pred_logits = score_hypotheses(smpl_discriminator, joints)
                loss_gen = loss_gen + evalops.compute_disc_loss(pred_logits, None)
            loss_values["smpl_gen"] = loss_gen * cfg["smpl_gen_loss"]["weight"]

        if "smpl_pseudo_img_loss" in cfg:
            loss_pseudo = 0
            for cam_id in cam_id_list:
                cam_key = "cam_{}".format(cam_id)
                logits, K, NH, NS = self._head(x["{}_pseudo_img".format(cam_key)])
                gt = x["{}_pseudo_joints".format(cam_key)]
                img_hw = tuple(x["{}_img".format(cam_key)].shape[-2:])
                if mono:
                    kps, _, _ = ops.integral_multi_head(logits, K, NH, NS)
                    from . import losses
                    lp = torch.min(torch.stack([losses.compute_supervision(kps[:, i], gt) for i in range(NH)]))
                else:
                    # pseudo-GT MSE per hypothesis + min (model.py:150-162) in the fused op; the camera tensors only feed its world output
                    lp, _, _, kps, _, _, _ = ops.integral_reproj_min_loss(logits, gt, _cams(x, cam_key), K, NH, NS, img_hw=img_hw,
                                                                          rect_width=2000.0, w_mse=1.0, reduction="batch", group=None)
                loss_pseudo = loss_pseudo + lp
                output["pose_2d_pred_{}_pseudo".format(cam_key)] = kps.detach()[0:1, 0, ...].clone()
                output["pose_3d_pred_{}_pseudo".format(cam_key)] = ops.convert_patch_to_world(
                    kps.detach()[:, 0, ...], x, cam_key, is_norm=True, RECT_WIDTH=256, mono=True, patch=False)[0:1, ...]
                output["pose_3d_gt_{}_pseudo".format(cam_key)] = ops.convert_patch_to_world(
                    gt, x, cam_key, is_norm=True, RECT_WIDTH=256, mono=True, patch=False)[0:1, ...]
            loss_values["smpl_pseudo_img"] = loss_pseudo * cfg["smpl_pseudo_img_loss"]["weight"]

        if "physique_recons_loss" in cfg and self.physique_network is not None:
            loss_phy = 0
            use_dis_map = cfg["physique_recons_loss"]["use_dis_map"]
            for cam_id in cam_id_list:
                cam_key = "cam_{}".format(cam_id)
                phy = self.physique_network(reconstructed[cam_key])
                output["mask_physique_{}".format(cam_key)] = phy.detach()[0:1, ...]
                loss_phy = loss_phy + skeleton.compute_mask_reconstruction_loss(
                    phy, x["{}_mask".format(cam_key)], weight=x["{}_geodesic_dis".format(cam_key)] if use_dis_map else None)
            loss_values["physique_recons"] = loss_phy * cfg["physique_recons_loss"]["weight"]

        if "recons_loss" in cfg:
            loss_values["reconstruction"] = loss_rec * cfg["recons_loss"]["weight"]
        return loss_values, output


class Counter3DDisc(torch.nn.Module):
    """modules/model.py:194-266: discriminator update.  The regressor's NH hypotheses go through ONE batched discriminator call."""

    def __init__(self, cfg, smpl_discriminator, smpl_layer, h36m_regressor):
        super().__init__()
        self.smpl_discriminator = smpl_discriminator
        self.cam_id_list = cfg["cam_id_list"]
        line_select_ids = cfg["line_select_ids"] if "line_select_ids" in cfg else None
        self.parent_ids, self.child_ids = cal_links(cfg["parent_ids"], line_select_ids=line_select_ids, use_root=False, extension=False)
        self.loss_config = cfg["loss_config"]
        if "GCN" in getattr(self.smpl_discriminator, "name", ""):
            self.smpl_discriminator.parent_ids = self.parent_ids
            self.smpl_discriminator.child_ids = self.child_ids
        self.smpl_layer = smpl_layer
        self.h36m_regressor = h36m_regressor
        self.DISC_SUP_DIMENSION = cfg["smpl_disc_params"]["disc_sup_dim"] if "disc_sup_dim" in cfg["smpl_disc_params"] else 3
        self.use_aug = cfg["smpl_disc_params"]["use_aug"] if "use_aug" in cfg["smpl_disc_params"] else False

    def forward(self, x, regressor):
        if self.use_aug:
            raise NotImplementedError("use_aug (random_rotation_3D of the discriminator input) is the caller's augmentation")
        loss_disc, output = 0, {}
        cam_id_list = ["mono"] if "cam_mono_img" in x else self.cam_id_list
        dim = self.DISC_SUP_DIMENSION
        for cam_id in cam_id_list:
            cam_key = "cam_{}".format(cam_id)
            pred_joints, _ = regressor(x["{}_img".format(cam_key)])
            smpl_joints = x["{}_pseudo_joints".format(cam_key)]
            smpl_joints_world = ops.convert_patch_to_world(smpl_joints, x, cam_key, is_norm=True, RECT_WIDTH=256, mono=True, patch=False)
            output["pose_smpl_2d_{}".format(cam_key)] = smpl_joints[0:1, ...]
            output["pose_smpl_3d_{}".format(cam_key)] = smpl_joints_world[0:1, ...].clone()
            pred_logits = score_hypotheses(self.smpl_discriminator, pred_joints.detach()[..., :dim])
            smpl_logits = self.smpl_discriminator(smpl_joints[..., :dim])
            output["smpl_logits_{}".format(cam_key)] = smpl_logits[0:1, ...]
            output["pred_logits_{}".format(cam_key)] = pred_logits[0:1, 0, ...]
            loss_disc = loss_disc + evalops.compute_disc_loss(pred_logits, smpl_logits)
        loss_disc = loss_disc * self.loss_config["smpl_disc_loss"]["weight"]
        return loss_disc, output


class GraphedLossGraph:
    """`Counter3DModel.forward` + backward of `sum(v.mean() for v in loss_values.values())` (train.py:182-184) captured once
    into a CUDA graph.  At the reference's batch size (32 per GPU, 4 cameras) the loss graph is a few dozen short launches
    and the step is bound by the host issuing them; replaying it as one graph removes that.

        g = GraphedLossGraph(model, x, smpl_discriminator)       # x: the data-loader dict (tensors are cloned into static buffers)
        g.x['cam_0_img'].detach().copy_(new_logits) ...          # refresh the static inputs in place
        loss_values = g()                                        # replay; g.grads[key] = d total / d x[key] for the '*_img' tensors

    Every tensor of `x` whose key ends in '_img' is treated as a differentiable input (it is what the backbone produces in
    the real pipeline); parameters of `smpl_discriminator` do not receive gradients here (the reference detaches its input
    in this pass, model.py:128)."""

    def __init__(self, model: "Counter3DModel", x, smpl_discriminator, warmup: int = 3):
        dev = next(v.device for v in x.values() if torch.is_tensor(v))
        self.x = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in x.items()}
        self.leaves = [k for k, v in self.x.items() if torch.is_tensor(v) and k.endswith("_img") and v.is_floating_point()]
        for k in self.leaves:
            self.x[k].requires_grad_(True)

        def run():
            for k in self.leaves:
                self.x[k].grad = None
            loss_values, output = model(self.x, smpl_discriminator)
            total = sum(v.mean() for v in loss_values.values())
            total.backward()
            return loss_values, output, total
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                run()
        torch.cuda.current_stream(dev).wait_stream(side)
        for k in self.leaves:
            self.x[k].grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            loss_values, output, total = run()
        self.loss_values = {k: v.detach() for k, v in loss_values.items()}
        self.output = output
        self.total = total.detach()
        self.grads = {k: self.x[k].grad for k in self.leaves}

    def __call__(self):
        self.graph.replay()
        return self.loss_values
