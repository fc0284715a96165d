"""Differentiable MANO forward for the TRAINING step (upstream common/nets/mano_head.py:185-256 +
manopth/manopth/manolayer.py:111-276; right hand, axis-angle, flat hand mean, centred on the wrist).

Inference runs the fused `hoisdf_mano_fwd` kernel (csrc/heads.cu); it has no backward.  The training step needs
d(verts, joints) / d(pose6d, shape) for 4 decoder layers x B hands -- a few hundred KFLOP per hand -- so it is written
with plain torch ops and left to autograd (bookkeeping-sized work, like the loss entries in nets/loss.py).  The chain
rot6d -> R -> quaternion -> axis-angle -> quaternion -> R' follows upstream step by step (incl. its NaN -> 0 guard and the
1e-8 inside Rodrigues' norm) so that the gradients agree with upstream's autograd, not only the values.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

# native MANO joint order: 0 wrist, then (index, middle, little, ring, thumb) x (proximal, middle, distal)
_PARENT = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]
_TIPS = [745, 317, 444, 556, 673]
_ORDER21 = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]


def rot6d_to_matrix(x6: torch.Tensor) -> torch.Tensor:
    """(N, 6) -> (N, 3, 3): Gram-Schmidt of the two 3-vectors, result columns (b1, b2, b1 x b2)."""
    u, w = x6[:, :3], x6[:, 3:]
    b1 = F.normalize(u)
    b2 = F.normalize(w - (b1 * w).sum(-1, keepdim=True) * b1)
    return torch.stack((b1, b2, torch.cross(b1, b2, dim=1)), dim=2)


def matrix_to_quaternion(R: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """(N, 3, 3) -> (N, 4) (w, x, y, z): the four-branch conversion upstream inherits from torchgeometry (it works on the
    transposed matrix); the branch is picked from the diagonal, gradients flow through the picked branch only."""
    m = R.transpose(1, 2)
    d0, d1, d2 = m[:, 0, 0], m[:, 1, 1], m[:, 2, 2]
    neg = d2 < eps
    a, b = d0 > d1, d0 < -d1
    branch = torch.where(neg, torch.where(a, 0, 1), torch.where(b, 2, 3))
    t = torch.stack([1 + d0 - d1 - d2, 1 - d0 + d1 - d2, 1 - d0 - d1 + d2, 1 + d0 + d1 + d2], 1)       # (N, 4)
    cand = torch.stack([
        torch.stack([m[:, 1, 2] - m[:, 2, 1], t[:, 0], m[:, 0, 1] + m[:, 1, 0], m[:, 2, 0] + m[:, 0, 2]], 1),
        torch.stack([m[:, 2, 0] - m[:, 0, 2], m[:, 0, 1] + m[:, 1, 0], t[:, 1], m[:, 1, 2] + m[:, 2, 1]], 1),
        torch.stack([m[:, 0, 1] - m[:, 1, 0], m[:, 2, 0] + m[:, 0, 2], m[:, 1, 2] + m[:, 2, 1], t[:, 2]], 1),
        torch.stack([t[:, 3], m[:, 1, 2] - m[:, 2, 1], m[:, 2, 0] - m[:, 0, 2], m[:, 0, 1] - m[:, 1, 0]], 1)], 1)
    idx = branch.view(-1, 1, 1).expand(-1, 1, 4)
    q = cand.gather(1, idx).squeeze(1)
    tt = t.gather(1, branch.view(-1, 1))
    return 0.5 * q / torch.sqrt(tt)


def quaternion_to_axis_angle(q: torch.Tensor) -> torch.Tensor:
    v = q[:, 1:]
    s2 = (v * v).sum(1)
    s = torch.sqrt(s2)
    c = q[:, 0]
    two_theta = 2.0 * torch.where(c < 0.0, torch.atan2(-s, -c), torch.atan2(s, c))
    k = torch.where(s2 > 0.0, two_theta / s, torch.full_like(s, 2.0))
    return v * k[:, None]


def matrix_to_axis_angle(R: torch.Tensor) -> torch.Tensor:
    aa = quaternion_to_axis_angle(matrix_to_quaternion(R))
    return torch.where(torch.isnan(aa), torch.zeros_like(aa), aa)      # upstream: aa[isnan(aa)] = 0


def axis_angle_to_matrix(aa: torch.Tensor) -> torch.Tensor:
    """manopth's batch_rodrigues: through a unit quaternion, the angle taken as |aa + 1e-8|."""
    ang = torch.norm(aa + 1e-8, p=2, dim=1, keepdim=True)
    axis = aa / ang
    q = torch.cat([torch.cos(0.5 * ang), torch.sin(0.5 * ang) * axis], 1)
    q = q / q.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q.unbind(1)
    return torch.stack([
        w * w + x * x - y * y - z * z, 2 * x * y - 2 * w * z, 2 * w * y + 2 * x * z,
        2 * w * z + 2 * x * y, w * w - x * x + y * y - z * z, 2 * y * z - 2 * w * x,
        2 * x * z - 2 * w * y, 2 * w * x + 2 * y * z, w * w - x * x - y * y + z * z], 1).view(-1, 3, 3)


_index_cache = {}


def _index(device, which: str) -> torch.Tensor:
    key = (str(device), which)
    t = _index_cache.get(key)
    if t is None:
        t = torch.tensor(_TIPS if which == "tips" else _ORDER21, device=device, dtype=torch.int64)
        _index_cache[key] = t
    return t


def mano_forward(layer, rot: torch.Tensor, betas: torch.Tensor):
    """`layer`: the ManoLayer module (buffers th_*); rot (N, 16, 3, 3) joint rotations, betas (N, 10)
    -> verts (N, 778, 3), joints (N, 21, 3) in metres, wrist-centred."""
    n = rot.shape[0]
    v_shaped = layer.th_v_template + torch.einsum("vck,nk->nvc", layer.th_shapedirs, betas)
    J = torch.einsum("jv,nvc->njc", layer.th_J_regressor, v_shaped)
    eye = torch.eye(3, device=rot.device, dtype=rot.dtype)
    v_posed = v_shaped + torch.einsum("vck,nk->nvc", layer.th_posedirs, (rot[:, 1:] - eye).reshape(n, 135))
    Rg, tg = [rot[:, 0]], [J[:, 0]]
    for i in range(1, 16):
        p = _PARENT[i]
        Rg.append(Rg[p] @ rot[:, i])
        tg.append(tg[p] + (Rg[p] @ (J[:, i] - J[:, p]).unsqueeze(-1)).squeeze(-1))
    Rg, tg = torch.stack(Rg, 1), torch.stack(tg, 1)                       # (N,16,3,3), (N,16,3)
    off = tg - (Rg @ J.unsqueeze(-1)).squeeze(-1)
    Rv = torch.einsum("vj,njab->nvab", layer.th_weights, Rg)
    tv = torch.einsum("vj,nja->nva", layer.th_weights, off)
    verts = (Rv @ v_posed.unsqueeze(-1)).squeeze(-1) + tv
    # (index tensors cached per device: indexing with a Python list builds one with a host copy on every call, which also
    # cannot be captured into a CUDA graph)
    joints = torch.cat([tg, verts.index_select(1, _index(rot.device, "tips"))], 1).index_select(1, _index(rot.device, "order"))
    centre = joints[:, :1]
    # upstream scales to millimetres inside ManoLayer and back to metres in ManoHead
    return (verts - centre) * 1000 / 1000, (joints - centre) * 1000 / 1000


def mano_head_train(head, pose6d_bm: torch.Tensor, shape_bm: torch.Tensor):
    """pose6d (L, B, 16, 6), shape (L, B, 10) -> pred dict of upstream ManoHead.forward (mano_head.py:232-256)."""
    l, b = pose6d_bm.shape[:2]
    R = rot6d_to_matrix(pose6d_bm.reshape(l * b * 16, 6))
    aa = matrix_to_axis_angle(R).view(l * b, 48)
    layer = head.mano_layer
    full = torch.cat([aa[:, :3], layer.th_hands_mean + aa[:, 3:]], 1)
    rot = axis_angle_to_matrix(full.reshape(-1, 3)).view(l * b, 16, 3, 3)
    verts, joints = mano_forward(layer, rot, shape_bm.reshape(l * b, 10))
    return {"verts3d": verts.view(l, b, 778, 3), "joints3d": joints.view(l, b, 21, 3), "mano_shape": shape_bm,
            "mano_pose": R.view(l, b, 16, 3, 3)}
