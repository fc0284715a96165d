"""The differentiable torch MANO of the training step (hoisdf_b200/nets/mano_torch.py) against the oracle's restatement
of upstream mano_head.py / manolayer.py: values and gradients w.r.t. the 6-D pose and the shape, on the CPU."""
import torch

from hoisdf_b200 import synthetic as syn
from hoisdf_b200.nets.mano_head import ManoHead, ManoLayer
from hoisdf_b200.nets.mano_torch import mano_head_train
from oracle import hoisdf_oracle as O


def test_mano_torch_matches_oracle_values_and_gradients():
    seed, L, B = 5, 3, 4
    bufs = syn.mano_buffers(seed)
    head = ManoHead(ManoLayer.from_buffers(bufs))
    g = torch.Generator().manual_seed(seed)
    pose = torch.randn(L, B, 16, 6, generator=g)
    pose[0, 0, 3] = torch.tensor([1.0, 0.0, 0.0, 0.0, 1.0, 0.0])          # identity rotation: the NaN -> 0 guard
    pose[1, 1, 5] = torch.tensor([-1.0, 0.1, 0.0, 0.0, -1.0, 0.2])        # a ~180 degree rotation: another quaternion branch
    shape = torch.randn(L, B, 10, generator=g) * 0.5
    wv, wj = torch.randn(L, B, 778, 3, generator=g), torch.randn(L, B, 21, 3, generator=g)

    p1, s1 = pose.clone().requires_grad_(), shape.clone().requires_grad_()
    pred = mano_head_train(head, p1, s1)
    ((pred["verts3d"] * wv).sum() + (pred["joints3d"] * wj).sum() + pred["mano_pose"].sum()).backward()

    p2, s2 = pose.clone().requires_grad_(), shape.clone().requires_grad_()
    params = {"mano_head.mano_layer." + k: v for k, v in bufs.items()}
    verts, joints = O.mano_head(params, p2.permute(0, 2, 1, 3), s2)      # oracle layout: pose6d (L, 16, B, 6)
    rot = O.rot6d_to_mat(p2.reshape(-1, 6))
    ((verts * wv).sum() + (joints * wj).sum() + rot.sum()).backward()

    assert float((pred["verts3d"] - verts).abs().max()) < 2e-6
    assert float((pred["joints3d"] - joints).abs().max()) < 2e-6
    assert float((pred["mano_pose"].reshape(-1, 3, 3) - rot).abs().max()) < 1e-6
    # an EXACT identity rotation has a singular quaternion -> axis-angle Jacobian (sqrt at 0): upstream's autograd returns
    # non-finite gradients for that joint's 6 inputs, and so must this implementation -- same entries, nowhere else
    for a, b in ((p1.grad, p2.grad), (s1.grad, s2.grad)):
        fa, fb = torch.isfinite(a), torch.isfinite(b)
        assert torch.equal(fa, fb)
        assert float((a[fa] - b[fb]).abs().max()) <= 2e-4 * float(b[fb].abs().max())
    assert not torch.isfinite(p2.grad[0, 0, 3]).all() and torch.isfinite(p2.grad[1]).all() and torch.isfinite(s2.grad).all()
