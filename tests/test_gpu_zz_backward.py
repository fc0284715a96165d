"""Backward kernels of the SDF branch (csrc/backward.cu) on the GPU, through the C ABI, against PyTorch autograd -- the same
checks tests/test_kernel_emulation.py runs on the CPU emulator.  (File name: sorted last, these kernels are groundwork
for the training step and are not called by the model yet.)"""
import ctypes as C

import pytest
import torch

from hoisdf_b200 import synthetic as syn
from oracle import hoisdf_oracle as O

pytestmark = pytest.mark.gpu


def _stream():
    return torch.cuda.current_stream().cuda_stream


def gemm(lib, a, ta, b, tb, out=None):
    m, k = (a.shape[1], a.shape[0]) if ta else a.shape
    n = b.shape[0] if tb else b.shape[1]
    c = torch.zeros(m, n, device=a.device) if out is None else out
    st = lib.hoisdf_gemm_f32(a.data_ptr(), a.stride(0), int(ta), b.data_ptr(), b.stride(0), int(tb), c.data_ptr(), n, m, n, k,
                             int(out is not None), _stream())
    assert st == 0, st
    return c


def test_backward_kernels_match_autograd(cuda):
    from hoisdf_b200 import _capi
    lib = _capi.lib
    g = torch.Generator().manual_seed(1)
    rnd = lambda *s: (torch.rand(*s, generator=g) * 2 - 1)        # noqa: E731
    # GEMM, all transposes, ragged sizes, accumulation
    m, n, k = 300, 45, 137
    for ta in (False, True):
        for tb in (False, True):
            a, b = rnd(*((k, m) if ta else (m, k))), rnd(*((n, k) if tb else (k, n)))
            ref = (a.T if ta else a).double() @ (b.T if tb else b).double()
            c = gemm(lib, a.to(cuda), ta, b.to(cuda), tb)
            assert float((c.cpu() - ref).abs().max()) < 3e-6 * float(ref.abs().max()), (ta, tb)
            c2 = gemm(lib, a.to(cuda), ta, b.to(cuda), tb, out=c.clone())
            assert float((c2.cpu() - 2 * ref).abs().max()) < 6e-6 * float(ref.abs().max())
    # ReLU mask + bias sums
    y, dy = torch.relu(rnd(530, 70)), rnd(530, 70)
    dz, db = dy.to(cuda).clone(), torch.zeros(70, device=cuda)
    assert lib.hoisdf_act_bias_bwd(dz.data_ptr(), 70, y.to(cuda).data_ptr(), 70, 530, 70, 1, db.data_ptr(), 0, _stream()) == 0
    want = dy * (y > 0)
    assert torch.equal(dz.cpu(), want) and float((db.cpu() - want.sum(0)).abs().max()) < 1e-4
    # weight norm
    rows, cols = 223, 512
    gg, v, dw = (rnd(rows, 1) * 0.5 + 1).requires_grad_(), rnd(rows, cols).requires_grad_(), rnd(rows, cols)
    (O.fold_weight_norm(gg, v) * dw).sum().backward()
    dg, dv = torch.zeros(rows, device=cuda), torch.zeros(rows, cols, device=cuda)
    assert lib.hoisdf_weight_norm_bwd(gg.detach().reshape(-1).to(cuda).data_ptr(), v.detach().to(cuda).data_ptr(),
                                      dw.to(cuda).data_ptr(), cols, rows, cols, dg.data_ptr(), dv.data_ptr(), 0, _stream()) == 0
    assert float((dg.cpu() - gg.grad.reshape(-1)).abs().max()) < 2e-5 and float((dv.cpu() - v.grad).abs().max()) < 2e-6
    # SDF loss head
    nrow, clamp = 2000, 0.05
    z, gt = (rnd(nrow) * 0.2).requires_grad_(), rnd(nrow) * 0.1
    (3.0 * torch.nn.functional.l1_loss(torch.clamp(torch.tanh(z), -clamp, clamp), torch.clamp(gt, -clamp, clamp))).backward()
    dzl = torch.zeros(nrow, device=cuda)
    assert lib.hoisdf_sdf_loss_bwd(z.detach().to(cuda).data_ptr(), gt.to(cuda).data_ptr(), nrow, clamp, 3.0, dzl.data_ptr(),
                                   _stream()) == 0
    # tanhf on the device and on the host differ by an ulp: only the clamp boundary cases may flip
    diff = (dzl.cpu() - z.grad).abs()
    assert float(diff.max()) < 1e-6 or int((diff > 1e-6).sum()) <= 2


def test_gather_backward_matches_grid_sample_autograd(cuda):
    from hoisdf_b200 import _capi
    lib = _capi.lib
    B, P = 2, 301
    gen = torch.Generator().manual_seed(4)
    maps = [torch.randn(B, c, h, h, generator=gen).requires_grad_() for c, h in ((32, 16), (64, 8), (128, 4))]
    uv = torch.rand(B, P, 2, generator=gen) * 295 - 20
    grid = O.grid_from_uv(uv, O.default_cfg()).unsqueeze(1)
    feats = torch.cat([torch.nn.functional.grid_sample(m, grid, padding_mode="border", align_corners=True) for m in maps], 1)
    feats = feats.squeeze(2).permute(0, 2, 1)
    dout = torch.rand(B, P, feats.shape[2], generator=gen) * 2 - 1
    (feats * dout).sum().backward()
    grads = [torch.zeros(B, m.shape[2], m.shape[3], m.shape[1], device=cuda) for m in maps]
    pyr = _capi.Pyramid()
    for i, t in enumerate(grads):
        pyr.map[i], pyr.c[i], pyr.h[i], pyr.w[i] = t.data_ptr(), t.shape[3], t.shape[1], t.shape[2]
    pyr.levels, pyr.img_h, pyr.img_w = len(grads), 256, 256
    uvd, dod = uv.reshape(-1, 2).contiguous().to(cuda), dout.reshape(B * P, -1).contiguous().to(cuda)
    assert lib.hoisdf_gather_bwd(C.byref(pyr), uvd.data_ptr(), B * P, None, B, P, dod.data_ptr(), dod.shape[1], _stream()) == 0
    for got, m in zip(grads, maps):
        want = m.grad.permute(0, 2, 3, 1)
        assert float((got.cpu() - want).abs().max()) < 1e-5 * max(1.0, float(want.abs().max()))


def _rnd(seed, *shape, lo=-1.0, hi=1.0):
    import numpy as np
    g = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy((g.random(size=shape, dtype=np.float32) * (hi - lo) + lo).astype(np.float32))


def test_layernorm_and_softmax_backward_on_gpu(cuda):
    """hoisdf_layernorm_bwd / hoisdf_softmax_rows_fwd / _bwd (the transformer layers' backward pieces) on the B200 against
    PyTorch autograd -- the checks tests/test_kernel_emulation.py runs on the CPU emulator."""
    from hoisdf_b200 import _capi
    lib = _capi.lib
    rows, d = 1045, 256
    h, gamma, beta = _rnd(1, rows, d, lo=-2, hi=2).requires_grad_(), _rnd(2, d, lo=0.5, hi=1.5).requires_grad_(), \
        _rnd(3, d).requires_grad_()
    dy = _rnd(4, rows, d)
    (torch.nn.functional.layer_norm(h, (d,), gamma, beta, 1e-5) * dy).sum().backward()
    dh, dg, dbt = torch.zeros(rows, d, device=cuda), torch.zeros(d, device=cuda), torch.zeros(d, device=cuda)
    stats = torch.zeros(rows * 2, device=cuda)
    hd, gd, dyd = h.detach().to(cuda), gamma.detach().to(cuda), dy.to(cuda)
    assert lib.hoisdf_layernorm_bwd(hd.data_ptr(), gd.data_ptr(), dyd.data_ptr(), rows, d, dh.data_ptr(), dg.data_ptr(),
                                    dbt.data_ptr(), stats.data_ptr(), 0, _stream()) == 0
    assert float((dh.cpu() - h.grad).abs().max()) < 5e-6
    assert float((dg.cpu() - gamma.grad).abs().max()) < 2e-5 * max(1.0, float(gamma.grad.abs().max()))
    assert float((dbt.cpu() - beta.grad).abs().max()) < 2e-5 * max(1.0, float(beta.grad.abs().max()))
    # row softmax with a key-validity limit, and its backward (in place)
    r, c, valid = 519, 800, 600
    s = _rnd(5, r, c, lo=-3, hi=3).requires_grad_()
    dp = _rnd(6, r, c)
    p_ref = torch.softmax(s[:, :valid], -1)
    (p_ref * dp[:, :valid]).sum().backward()
    p = torch.full((r, c), 7.0, device=cuda)
    sdv = s.detach().to(cuda)
    assert lib.hoisdf_softmax_rows_fwd(sdv.data_ptr(), c, r, c, valid, None, 0, p.data_ptr(), c, _stream()) == 0
    assert float((p[:, :valid].cpu() - p_ref.detach()).abs().max()) < 1e-6 and not bool(p[:, valid:].any())
    ds = dp.to(cuda).clone()
    assert lib.hoisdf_softmax_rows_bwd(p.data_ptr(), c, ds.data_ptr(), c, r, c, ds.data_ptr(), c, _stream()) == 0
    assert float((ds[:, :valid].cpu() - s.grad[:, :valid]).abs().max()) < 1e-6 and not bool(ds[:, valid:].any())


def test_adamw_step_on_gpu(cuda):
    """hoisdf_adamw_step against torch.optim.AdamW (upstream common/base.py:68) over three updates."""
    from hoisdf_b200 import _capi
    lib = _capi.lib
    n = 100003
    w = _rnd(1, n).requires_grad_()
    opt = torch.optim.AdamW([w], lr=1e-4)
    p, m, v = w.detach().clone().to(cuda), torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    for step in range(1, 4):
        g = _rnd(10 + step, n, lo=-3, hi=3)
        w.grad = g.clone()
        opt.step()
        gd = g.to(cuda)
        assert lib.hoisdf_adamw_step(p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-4, 0.9, 0.999, 1e-8, 0.01,
                                     step, _stream()) == 0
        assert float((p.cpu() - w.detach()).abs().max()) < 2e-7
    st = opt.state[w]
    assert float((m.cpu() - st["exp_avg"]).abs().max()) < 1e-6 and float((v.cpu() - st["exp_avg_sq"]).abs().max()) < 1e-6


def test_vote_loss_backward_on_gpu(cuda):
    """hoisdf_vote_loss_bwd (JointvoteLoss, upstream common/nets/loss.py:22-61) against autograd of the oracle."""
    from hoisdf_b200 import _capi
    lib = _capi.lib
    L, B, Pn = 3, 4, 600
    pts = _rnd(1, B, Pn, 3, lo=-0.1, hi=0.1)
    gt = _rnd(2, B, 20, 3, lo=-100, hi=100)                        # millimetres
    off, cls = _rnd(3, L, B, Pn, 60, lo=-0.05, hi=0.05), _rnd(4, L, B, Pn, 20, lo=-3, hi=3)
    cfg = O.default_cfg(hand_cls_dist=0.06)
    off_t, cls_t = off.clone().requires_grad_(), cls.clone().requires_grad_()
    off_u, cls_u = off_t.permute(0, 2, 1, 3), cls_t.permute(0, 2, 1, 3)          # upstream layout (L, P, B, .)
    joints = O.vote_joints(pts, off_u, cls_u)
    l1, l2, l3 = O.joint_vote_losses(pts, off_u, cls_u, joints, gt, cfg)
    gw = (1.0, 0.5, 2.0)
    (gw[0] * l1 + gw[1] * l2 + gw[2] * l3).backward()
    d_off, d_cls, npos = torch.zeros_like(off, device=cuda), torch.zeros_like(cls, device=cuda), torch.zeros(1, device=cuda)
    args = [t.to(cuda).contiguous() for t in (pts, off, cls, gt)]
    assert lib.hoisdf_vote_loss_bwd(args[0].data_ptr(), args[1].data_ptr(), args[2].data_ptr(), args[3].data_ptr(), L, B, Pn,
                                    cfg.hand_cls_dist, gw[0], gw[1], gw[2], d_off.data_ptr(), d_cls.data_ptr(), npos.data_ptr(),
                                    _stream()) == 0
    mask = (torch.norm(pts.unsqueeze(2) - gt.unsqueeze(1) / 1000, dim=-1) < cfg.hand_cls_dist)
    assert 0 < int(mask.sum()) < mask.numel() and float(npos[0]) == float(mask.sum())
    assert float((d_off.cpu() - off_t.grad).abs().max()) < 2e-5 * float(off_t.grad.abs().max())
    assert float((d_cls.cpu() - cls_t.grad).abs().max()) < 2e-5 * float(cls_t.grad.abs().max())


def test_tokens_backward_on_gpu(cuda):
    """hoisdf_tokens_bwd (token assembly + SDF activation, upstream main/model.py:123-126,520-531) against autograd."""
    from hoisdf_b200 import _capi
    lib = _capi.lib
    B, Pn, S, t0 = 4, 600, 800, 200
    fea, sdf = _rnd(1, B, Pn, 223).requires_grad_(), _rnd(2, B, Pn, lo=-0.15, hi=0.15).requires_grad_()
    beta = torch.tensor([0.1], requires_grad=True)
    d_tok = _rnd(3, B, S, 256)
    tok_fea = fea * (torch.sigmoid(sdf[..., None] / beta) / beta)
    (tok_fea * d_tok[:, t0:t0 + Pn, 33:]).sum().backward()
    d_fea, d_sdf, d_beta = torch.zeros(B * Pn, 223, device=cuda), torch.zeros(B * Pn, device=cuda), torch.zeros(1, device=cuda)
    nbytes = lib.hoisdf_tokens_bwd_workspace_bytes(B, Pn)
    ws = torch.zeros(nbytes // 4 + 1, device=cuda)
    dt, fd, sdv, bd = d_tok.to(cuda), fea.detach().reshape(B * Pn, 223).to(cuda), sdf.detach().reshape(-1).to(cuda), \
        beta.detach().to(cuda)
    assert lib.hoisdf_tokens_bwd(dt.data_ptr(), S, t0, fd.data_ptr(), 223, sdv.data_ptr(), bd.data_ptr(), B, Pn, d_fea.data_ptr(),
                                 223, d_sdf.data_ptr(), d_beta.data_ptr(), 0, ws.data_ptr(), nbytes, _stream()) == 0
    assert float((d_fea.cpu() - fea.grad.reshape(B * Pn, 223)).abs().max()) < 1e-5 * float(fea.grad.abs().max())
    assert float((d_sdf.cpu() - sdf.grad.reshape(-1)).abs().max()) < 1e-5 * float(sdf.grad.abs().max())
    assert abs(float(d_beta[0]) - float(beta.grad)) < 1e-4 * abs(float(beta.grad))
