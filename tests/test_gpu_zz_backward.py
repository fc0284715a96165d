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
