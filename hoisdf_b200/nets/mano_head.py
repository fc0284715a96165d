"""`ManoHead` / `ManoLayer` with the upstream buffer names (common/nets/mano_head.py:220-278,
manopth/manopth/manolayer.py:74-106), running on the fused hoisdf_b200 MANO kernel
(rot6d -> axis-angle -> Rodrigues -> blend shapes -> kinematic chain -> LBS, one CTA per hand).

The licensed MANO_RIGHT.pkl is not redistributable; `ManoLayer.from_buffers` builds the layer from tensors
(a released checkpoint's state-dict carries them as `mano_head.mano_layer.th_*`), and `ManoLayer(mano_root=...)`
loads the pkl the way upstream does when `chumpy` and the file are available.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import _capi, ops

_BUFFER_SHAPES = {
    "th_betas": (1, 10), "th_shapedirs": (778, 3, 10), "th_posedirs": (778, 3, 135), "th_v_template": (1, 778, 3),
    "th_J_regressor": (16, 778), "th_weights": (778, 16), "th_hands_mean": (1, 45),
    "th_selected_comps": (45, 45),
}


class ManoLayer(nn.Module):
    """Right hand, axis-angle input, flat hand mean, centred on joint 0 (what upstream model.py:735-742 builds)."""

    def __init__(self, center_idx=0, flat_hand_mean=True, ncomps=45, side="right", mano_root="tool/mano_models",
                 use_pca=False, root_rot_mode="axisang", joint_rot_mode="axisang", robust_rot=False, buffers=None):
        super().__init__()
        if center_idx != 0 or not flat_hand_mean or side != "right" or use_pca or root_rot_mode != "axisang" \
                or joint_rot_mode != "axisang":
            raise NotImplementedError("hoisdf_b200.ManoLayer supports the upstream model.py:735-742 configuration only")
        self.center_idx, self.side, self.use_pca, self.ncomps, self.rot = 0, side, False, 45, 3
        if buffers is None:
            buffers = _load_mano_pkl(os.path.join(mano_root, "MANO_RIGHT.pkl"))
        for name, shape in _BUFFER_SHAPES.items():
            t = buffers.get(name)
            if t is None:
                t = torch.eye(45) if name == "th_selected_comps" else torch.zeros(shape)
            self.register_buffer(name, torch.as_tensor(t, dtype=torch.float32).reshape(shape).clone())
        faces = buffers.get("th_faces", torch.zeros(1538, 3, dtype=torch.long))
        self.register_buffer("th_faces", torch.as_tensor(faces).long().reshape(1538, 3).clone())

    @classmethod
    def from_buffers(cls, buffers):
        return cls(buffers=dict(buffers))

    def _struct(self):
        bufs = [self.th_shapedirs, self.th_posedirs, self.th_v_template, self.th_J_regressor, self.th_weights,
                self.th_hands_mean]
        for t in bufs:
            assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float32
        return _capi.ManoModel(*[t.data_ptr() for t in bufs])

    def forward_6d(self, pose6d: torch.Tensor, betas: torch.Tensor):
        """pose6d (N,16,6), betas (N,10) -> verts (N,778,3), joints (N,21,3) in metres."""
        return ops.mano(self._struct(), pose6d.contiguous(), betas.contiguous())

    def forward_aa(self, pose_aa: torch.Tensor, betas: torch.Tensor):
        """axis-angle pose (N,48), betas (N,10) -> verts, joints in metres (ground-truth branch)."""
        return ops.mano_aa(self._struct(), pose_aa.contiguous().float(), betas.contiguous().float())


def _load_mano_pkl(path):
    try:
        import pickle

        import numpy as np
        with open(path, "rb") as fh:
            d = pickle.load(fh, encoding="latin1")  # needs chumpy importable, like upstream manolayer.py:66
    except Exception as e:  # noqa
        raise RuntimeError(
            "cannot load %s (%s); pass `buffers=` or load a checkpoint whose state-dict carries "
            "mano_head.mano_layer.th_*" % (path, e))
    arr = lambda k: torch.as_tensor(np.array(d[k]), dtype=torch.float32)  # noqa
    return {
        "th_betas": torch.zeros(1, 10), "th_shapedirs": arr("shapedirs"), "th_posedirs": arr("posedirs"),
        "th_v_template": arr("v_template").unsqueeze(0),
        "th_J_regressor": torch.as_tensor(np.array(d["J_regressor"].toarray()), dtype=torch.float32),
        "th_weights": arr("weights"), "th_hands_mean": torch.zeros(1, 45),
        "th_faces": torch.as_tensor(np.array(d["f"]).astype("int64")),
        "th_selected_comps": arr("hands_components")[:45],
    }


class ManoHead(nn.Module):
    def __init__(self, mano_layer, coord_change_mat=None):
        super().__init__()
        self.mano_layer = mano_layer
        self.mano_pose_size = 16 * 3
        if coord_change_mat is not None:
            self.register_buffer("coord_change_mat", coord_change_mat)
        else:
            self.coord_change_mat = None

    def forward_bm(self, pose6d_bm: torch.Tensor, shape_bm: torch.Tensor):
        """pose6d (L,B,16,6), shape (L,B,10) batch-major -> verts (L,B,778,3), joints (L,B,21,3)."""
        l, b = pose6d_bm.shape[:2]
        v, j = self.mano_layer.forward_6d(pose6d_bm.reshape(l * b, 16, 6), shape_bm.reshape(l * b, 10))
        return v.view(l, b, 778, 3), j.view(l, b, 21, 3)

    def forward_gt(self, mano_params: torch.Tensor):
        """Ground-truth branch of upstream mano_head.py:258-276: mano_params (B,58) = axis-angle pose (48) | shape (10).
        Upstream copies the pose slice (`.contiguous()` of a column slice, :260) before subtracting th_hands_mean,
        so the caller's tensor is left untouched; same here."""
        gt_shape = mano_params[:, self.mano_pose_size:].to(torch.float32)
        gt_pose = mano_params[:, : self.mano_pose_size].to(torch.float32).clone()
        gt_pose[:, 3:] = gt_pose[:, 3:] - self.mano_layer.th_hands_mean
        verts, joints = self.mano_layer.forward_aa(gt_pose, gt_shape)
        return {"verts3d": verts, "joints3d": joints, "mano_shape": gt_shape,
                "mano_pose": batch_rodrigues(gt_pose.reshape(-1, 3)).view(-1, 16, 3, 3)}

    def forward(self, pose6d, shape, mano_params=None):
        """Upstream signature (mano_head.py:232): pose6d (L,16,B,6), shape (L,B,10)."""
        p6 = pose6d.permute(0, 2, 1, 3).contiguous()
        v, j = self.forward_bm(p6, shape.contiguous())
        pred = {"verts3d": v, "joints3d": j, "mano_shape": shape, "mano_pose": rot6d2mat(p6.reshape(-1, 6)).view(
            p6.shape[0], p6.shape[1], 16, 3, 3)}
        return pred, (None if mano_params is None else self.forward_gt(mano_params))


def rot6d2mat(x: torch.Tensor) -> torch.Tensor:
    """(N,6) -> (N,3,3), columns b1 b2 b3 (upstream mano_head.py:185-194); only feeds the eval-mode pose loss."""
    a1, a2 = x[:, 0:3], x[:, 3:6]
    b1 = torch.nn.functional.normalize(a1)
    b2 = torch.nn.functional.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1)
    return torch.stack((b1, b2, torch.cross(b1, b2, dim=1)), dim=-1)


def batch_rodrigues(theta: torch.Tensor) -> torch.Tensor:
    """(N,3) axis-angle -> (N,3,3) through a quaternion (upstream mano_head.py:12-51); loss bookkeeping only."""
    angle = torch.norm(theta + 1e-8, p=2, dim=1, keepdim=True)
    n = theta / angle
    q = torch.cat([torch.cos(angle * 0.5), torch.sin(angle * 0.5) * n], dim=1)
    q = q / q.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([w * w + x * x - y * y - z * z, 2 * x * y - 2 * w * z, 2 * w * y + 2 * x * z,
                        2 * w * z + 2 * x * y, w * w - x * x + y * y - z * z, 2 * y * z - 2 * w * x,
                        2 * x * z - 2 * w * y, 2 * w * x + 2 * y * z, w * w - x * x - y * y + z * z], dim=1).view(-1, 3, 3)
