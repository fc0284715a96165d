"""Loss entries that upstream `Model.forward` also returns in eval mode (main/model.py:631-662; main/test.py:127-129
reduces and then ignores them).  They are harness-side bookkeeping, not the hot path: plain torch ops on the
batch-major head outputs.  The part of upstream's JointvoteLoss that PRODUCES the predicted joints
(common/nets/loss.py:31-36,54-57) is the hoisdf_vote_joints_fwd kernel, not this file.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ..config import cfg


def joint_vote_losses(points, hand_off, hand_cls, hand_joints, joint_gt):
    """points (B,P,3), hand_off (L,B,P,60), hand_cls (L,B,P,20), hand_joints (L,B,20,3), joint_gt (B,20,3) [mm].
    Mirrors common/nets/loss.py:37-61 for batch-major tensors."""
    l, b, p, j = hand_cls.shape
    vote = points[None, :, :, None, :] + hand_off.view(l, b, p, j, 3)
    cls_gt = (torch.norm(points[:, :, None, :] - joint_gt[:, None] / 1000, dim=-1) < cfg.hand_cls_dist).float()
    gt = joint_gt[None, :, None].expand(l, b, p, j, 3)
    reg = F.smooth_l1_loss(vote * 1000, gt, reduction="none") * cls_gt[None, ..., None]
    loss_joint_3d = (reg.sum((1, 2, 3)) / cls_gt.sum()).mean()
    loss_joint_cls = F.binary_cross_entropy_with_logits(hand_cls, cls_gt[None].expand(l, b, p, j))
    loss_all = F.smooth_l1_loss(hand_joints * 1000, joint_gt[None].expand(l, b, j, 3))
    return loss_joint_3d, loss_joint_cls, loss_all


def render_gaussian_heatmap(joint_coord):
    """upstream model.py:128-143: sum of 21 isotropic Gaussians on the 128x128 output grid, x255."""
    h, w = cfg.output_hm_shape[1], cfg.output_hm_shape[2]
    yy, xx = torch.meshgrid(torch.arange(h, device=joint_coord.device), torch.arange(w, device=joint_coord.device),
                            indexing="ij")
    xx, yy = xx[None, None].float(), yy[None, None].float()
    x, y = joint_coord[:, :, 0, None, None], joint_coord[:, :, 1, None, None]
    hm = torch.exp(-(((xx - x) / cfg.sigma) ** 2) / 2 - (((yy - y) / cfg.sigma) ** 2) / 2)
    return hm.sum(1) * 255


def dexycb_losses(out, taps, targets, decoder_out, hand_sdf_sample, obj_sdf_sample, pred_mano, gt_mano):
    """The extra entries of the dexycb eval branch (upstream model.py:393-422,640-654)."""
    c = cfg.ClampingDistance
    loss = {
        "sdfhand_loss": F.l1_loss(hand_sdf_sample, targets["hand_sdf"].clamp(-c, c).unsqueeze(-1)),
        "sdfobj_loss": F.l1_loss(obj_sdf_sample, targets["obj_sdf"].clamp(-c, c).unsqueeze(-1)),
        "joint_heatmap": (decoder_out[:, 0] - render_gaussian_heatmap(targets["joint_coord"])) ** 2,
        "obj_seg": F.binary_cross_entropy(decoder_out[:, 2], targets["obj_seg"], reduction="none"),
        "hand_seg": F.binary_cross_entropy(decoder_out[:, 1], targets["hand_seg"], reduction="none"),
    }
    exp = lambda k: gt_mano[k].unsqueeze(0).expand(pred_mano[k].shape)  # noqa: E731
    loss["mano_mesh_loss"] = cfg.lambda_verts3d * F.mse_loss(pred_mano["verts3d"], exp("verts3d"))
    loss["mano_joint_loss"] = cfg.lambda_joints3d * F.mse_loss(pred_mano["joints3d"], exp("joints3d"))
    loss["pose_param_loss"] = cfg.lambda_manopose * F.mse_loss(pred_mano["mano_pose"], exp("mano_pose"))
    loss["shape_param_loss"] = cfg.lambda_manoshape * F.mse_loss(pred_mano["mano_shape"], exp("mano_shape"))
    return loss


def eval_losses(taps, targets, meta_info, joint_gt=None):
    b = taps["hand_points_notrans"].shape[0]
    dev = taps["hand_points_notrans"].device
    if joint_gt is None:
        joint_gt = torch.zeros(b, 20, 3, device=dev)      # upstream model.py:629 (ho3d eval has no joint GT)
    l3d, lcls, lall = joint_vote_losses(taps["hand_points_notrans"], taps["hand_off"], taps["hand_cls"],
                                        taps["hand_joints"], joint_gt)
    obj_rot, obj_trans = taps["obj_rot"], taps["obj_trans"]
    rot_gt = targets["obj_rot"].to(dev)[None, :, None, :].expand_as(obj_rot)
    trans_gt = targets["rel_obj_trans"].to(dev)[None, :, None, :].expand_as(obj_trans)
    return {
        "loss_joint_3d": l3d, "loss_joint_cls": lcls, "loss_all_joint_3d": lall,
        "obj_rot": F.smooth_l1_loss(obj_rot, rot_gt), "obj_trans": F.smooth_l1_loss(obj_trans, trans_gt),
    }
