"""Photometric augmentation of the training feed (csrc/augment.cu; upstream data/ho3d.py:355-364, data/dataset_util.py:144-201)
executed unchanged on the CPU emulator against the libraries upstream calls -- Pillow's ImageFilter.GaussianBlur and
torchvision's PIL colour adjustments -- bit for bit."""
import ctypes as C

import numpy as np
import pytest
from PIL import Image, ImageFilter

from test_kernel_emulation import build_emulated


@pytest.fixture(scope="module")
def emu():
    lib = build_emulated("augment")
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    lib.hoisdf_gaussian_blur_params.argtypes = [C.c_float, i32, vp]
    lib.hoisdf_gaussian_blur_u8.argtypes = [vp, vp, vp, i64, i64, i64, i64, vp, i32, vp]
    return lib


def blur_params(lib, radii, passes=3):
    out = np.zeros((len(radii), 3), np.uint32)
    for i, r in enumerate(radii):
        assert lib.hoisdf_gaussian_blur_params(float(r), passes, out[i].ctypes.data) == 0
    return out


def emu_blur(lib, imgs, radii):
    imgs = np.ascontiguousarray(imgs)
    b, h, w, ch = imgs.shape
    params = blur_params(lib, radii)
    dst = np.full_like(imgs, 9)
    scratch = np.zeros_like(imgs)
    rc = lib.hoisdf_gaussian_blur_u8(imgs.ctypes.data, dst.ctypes.data, scratch.ctypes.data, b, h, w, ch, params.ctypes.data, 3,
                                     None)
    assert rc == 0
    return dst


def pil_blur(img, radius):
    pil = Image.fromarray(img if img.shape[2] == 3 else img[:, :, 0])
    return np.asarray(pil.filter(ImageFilter.GaussianBlur(radius))).reshape(img.shape)


@pytest.mark.parametrize("h,w,ch", [(64, 64, 3), (37, 53, 3), (5, 3, 3), (1, 9, 3), (9, 1, 1), (40, 70, 1)])
def test_gaussian_blur_matches_pillow_for_upstream_radii(emu, h, w, ch):
    """radius = random.random() * 0.5 (ho3d.py:356): box radius < 1, a 3-tap filter per pass."""
    rng = np.random.default_rng(h * 100 + w)
    radii = [0.0, 0.5, 0.25] + list(rng.uniform(0, 0.5, 5))
    imgs = rng.integers(0, 256, (len(radii), h, w, ch), dtype=np.uint8)
    got = emu_blur(emu, imgs, radii)
    for i, r in enumerate(radii):
        assert np.array_equal(got[i], pil_blur(imgs[i], r)), (i, r)
    assert np.array_equal(got[0], imgs[0])                       # radius 0 is the identity


def test_gaussian_blur_matches_pillow_for_large_radii(emu):
    """Box radii >= 1 (2 n + 3 taps), lines shorter than the box, per-sample radii in one batch."""
    rng = np.random.default_rng(3)
    radii = [0.9, 1.5, 2.0, 3.7, 6.0, 11.0]
    for h, w in ((31, 45), (4, 6)):
        imgs = rng.integers(0, 256, (len(radii), h, w, 3), dtype=np.uint8)
        got = emu_blur(emu, imgs, radii)
        for i, r in enumerate(radii):
            assert np.array_equal(got[i], pil_blur(imgs[i], r)), (h, w, r)
    assert blur_params(emu, [6.0])[0, 0] >= 2                                   # the loop over taps was exercised


def test_blur_params_follow_pillow_over_a_radius_sweep(emu):
    """The host helper's (n, ww, fw) are what Pillow uses: a 2 x 2 checker blurred by Pillow equals the 3-tap formula with
    them for 400 radii (any difference of one unit in ww shows at this contrast)."""
    radii = np.linspace(0.0, 1.4, 400)
    params = blur_params(emu, radii)
    assert (params[:, 0] == 0).all()
    line = np.array([0, 255, 0, 0, 255, 255, 0, 255], np.int64)
    img = np.repeat(line[None, :, None], 3, axis=2).astype(np.uint8)                 # (1, 8, 3): only the horizontal passes act
    for r, (n, ww, fw) in zip(radii, params.astype(np.int64)):
        cur = line
        for _ in range(3):
            left, right = np.concatenate([cur[:1], cur[:-1]]), np.concatenate([cur[1:], cur[-1:]])
            cur = (cur * ww + (left + right) * fw + (1 << 23)) >> 24
        want = np.asarray(Image.fromarray(img).filter(ImageFilter.GaussianBlur(float(r))))[0, :, 0]
        assert np.array_equal(cur, want), r


def test_blur_argument_checks(emu):
    buf = np.zeros(64, np.uint8)
    p = np.zeros(3, np.uint32)
    call = emu.hoisdf_gaussian_blur_u8
    assert call(None, buf.ctypes.data, buf.ctypes.data, 1, 2, 2, 3, p.ctypes.data, 3, None) == -1
    assert call(buf.ctypes.data, buf.ctypes.data, buf.ctypes.data, 1, 2, 2, 2, p.ctypes.data, 3, None) == -2
    assert call(buf.ctypes.data, buf.ctypes.data, buf.ctypes.data, 1, 2, 9000, 3, p.ctypes.data, 3, None) == -2
    assert emu.hoisdf_gaussian_blur_params(-1.0, 3, p.ctypes.data) == -2
