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


# ---------------------------------------------------------------------------------------------- colour jitter
BRIGHTNESS, SATURATION, HUE, CONTRAST = 1, 2, 3, 4


@pytest.fixture(scope="module")
def emu_jitter(emu):
    vp, i64 = C.c_void_p, C.c_int64
    emu.hoisdf_color_jitter_u8.argtypes = [vp, vp, i64, i64, i64, vp, vp, vp, vp]
    return emu


def emu_jitter_run(lib, imgs, steps):
    """steps: per sample a list of up to four (op, factor) pairs (hue factor = torchvision's hue_factor)."""
    imgs = np.ascontiguousarray(imgs)
    b, h, w, _ = imgs.shape
    ops = np.zeros((b, 4), np.int32)
    fac = np.zeros((b, 4), np.float32)
    for i, seq in enumerate(steps):
        for j, (op, f) in enumerate(seq):
            ops[i, j] = op
            fac[i, j] = float(np.int32(f * 255).astype(np.uint8)) if op == HUE else f
    dst = np.full_like(imgs, 3)
    sums = np.full(4 * b, 77, np.uint64)
    assert lib.hoisdf_color_jitter_u8(imgs.ctypes.data, dst.ctypes.data, b, h, w, ops.ctypes.data, fac.ctypes.data,
                                      sums.ctypes.data, None) == 0
    return dst


def tv_jitter(img, seq):
    import torchvision.transforms.functional as TF
    fn = {BRIGHTNESS: TF.adjust_brightness, SATURATION: TF.adjust_saturation, HUE: TF.adjust_hue, CONTRAST: TF.adjust_contrast}
    pil = Image.fromarray(img)
    for op, f in seq:
        pil = fn[op](pil, f)
    return np.asarray(pil)


def all_colours(step=1):
    v = np.arange(0, 256, step, dtype=np.uint8)
    r, g, b = np.meshgrid(v, v, v, indexing="ij")
    n = len(v)
    return np.stack([r, g, b], -1).reshape(n * n, n, 3)


def test_every_adjustment_matches_torchvision_over_the_colour_cube(emu_jitter):
    """Each of the four adjustments alone, on an image holding a 52^3 lattice of the colour cube + random colours, for factors
    inside and outside [0, 1] (ImagingBlend's two branches) and hue shifts of both signs."""
    rng = np.random.default_rng(0)
    img = np.concatenate([all_colours(5), rng.integers(0, 256, (300, 52, 3), dtype=np.uint8)])
    cases = [[(BRIGHTNESS, 0.5)], [(BRIGHTNESS, 1.4999)], [(BRIGHTNESS, 1.0)], [(SATURATION, 0.61)], [(SATURATION, 1.37)],
             [(CONTRAST, 0.55)], [(CONTRAST, 1.45)], [(HUE, 0.15)], [(HUE, -0.15)], [(HUE, 0.003)], [(HUE, -0.5)], [(HUE, 0.0)]]
    got = emu_jitter_run(emu_jitter, np.stack([img] * len(cases)), cases)
    for i, seq in enumerate(cases):
        assert np.array_equal(got[i], tv_jitter(img, seq)), seq


def test_hue_matches_torchvision_over_the_colour_cube(emu_jitter):
    """Pillow's RGB -> HSV -> RGB round trip with a shift over a 128^3 lattice of the colour cube (2 M colours, ~10 s on the
    emulator); HOISDF_EXHAUSTIVE=1 runs all 2^24 colours (90 s; passed when this test was written)."""
    import os
    img = all_colours(1 if os.environ.get("HOISDF_EXHAUSTIVE") == "1" else 2)
    img = img.reshape(-1, 4096 if img.shape[0] * img.shape[1] % 4096 == 0 else img.shape[1], 3)
    got = emu_jitter_run(emu_jitter, img[None], [[(HUE, 0.1)]])
    assert np.array_equal(got[0], tv_jitter(img, [(HUE, 0.1)]))


def test_shuffled_sequences_match_upstream_color_jitter(emu_jitter):
    """Whole `color_jitter` sequences (dataset_util.py:167-201): factors drawn as `get_color_params` draws them with upstream's
    ranges (ho3d.py:37-40: hue 0.15, saturation / contrast / brightness 0.5), all 24 orders over the batch; the contrast
    step's mean is that of the image as it is at that step."""
    import itertools
    rng = np.random.default_rng(7)
    orders = list(itertools.permutations([BRIGHTNESS, SATURATION, HUE, CONTRAST]))
    imgs = rng.integers(0, 256, (len(orders), 24, 31, 3), dtype=np.uint8)
    imgs[1] = (imgs[1] // 4)                                       # a dark image: contrast mean far from 128
    steps = []
    for order in orders:
        f = {BRIGHTNESS: rng.uniform(0.5, 1.5), SATURATION: rng.uniform(0.5, 1.5), CONTRAST: rng.uniform(0.5, 1.5),
             HUE: rng.uniform(-0.15, 0.15)}
        steps.append([(op, float(f[op])) for op in order])
    steps[5] = steps[5][:2]                                        # fewer than four adjustments (a range of 0 drops one)
    steps[6] = []
    got = emu_jitter_run(emu_jitter, imgs, steps)
    for i, seq in enumerate(steps):
        assert np.array_equal(got[i], tv_jitter(imgs[i], seq)), seq
    assert np.array_equal(got[6], imgs[6])


def test_jitter_argument_checks(emu_jitter):
    buf = np.zeros(64, np.uint8)
    z = np.zeros(8, np.int64)
    call = emu_jitter.hoisdf_color_jitter_u8
    assert call(None, buf.ctypes.data, 1, 2, 2, z.ctypes.data, z.ctypes.data, z.ctypes.data, None) == -1
    assert call(buf.ctypes.data, buf.ctypes.data, 1, 0, 2, z.ctypes.data, z.ctypes.data, z.ctypes.data, None) == -2


# ---------------------------------------------------------------------------------------------- the whole training image
def emu_training_image(lib, frame, center, scale, rot, radius, steps):
    """warp -> GaussianBlur -> colour jitter -> ToTensor / 255 through the emulated kernels: the four feed calls of a training
    frame in upstream's order (ho3d.py:351-364, :550)."""
    from hoisdf_b200 import feed
    from test_feed import emu_warp
    vp, i64 = C.c_void_p, C.c_int64
    lib.hoisdf_image_crop_fwd.argtypes = [vp, i64, i64, i64, i64, i64, i64, vp, vp, i64, C.c_float, vp, vp, vp, vp]
    coef = feed.pil_coefficients(feed.crop_affine(center, scale, 256, rot))[None]
    _, warped = emu_warp(lib, frame[None], coef, 256)
    blurred = emu_blur(lib, warped, [radius])
    named = {"brightness": BRIGHTNESS, "saturation": SATURATION, "hue": HUE, "contrast": CONTRAST}
    jittered = emu_jitter_run(lib, blurred, [[(named[n], f) for n, f in steps]])
    as_float, _ = emu_warp(lib, jittered, np.array([[1.0, 0, 0, 0, 1.0, 0]]), 256)
    return as_float[0]


def test_training_image_matches_the_unmodified_upstream_item_live(emu_jitter):
    """`ho3d.Dataset.__getitem__` (mode "train") with its blur and colour jitter ON (constructor defaults) against the emulated
    kernels fed with the product's own draws after the same `random.seed`: the network input image, bit for bit."""
    import random
    from hoisdf_b200 import feed
    from oracle import reference_shim as rs
    if not rs.available():
        pytest.skip("upstream reference not mounted")
    for seed in (4, 9):
        inputs, _, _, taps = rs.ho3d_train_item(seed, filters=True)
        a = taps["affine"][0]
        random.seed(seed)
        radius = random.random() * 0.5                                              # ho3d.py:356
        steps = feed.draw_color_jitter(brightness=0.5, contrast=0.5, saturation=0.5, hue=0.15)
        assert len(steps) == 4
        got = emu_training_image(emu_jitter, taps["frame"], a["center"], a["scale"], a["rot"], radius, steps)
        assert np.array_equal(got, inputs["img"].numpy()), seed


def test_training_image_matches_the_upstream_fixture(emu_jitter):
    """The same against the committed fixture of one upstream item (tests/golden/feed_seed31.npz, every 8th row of the image)."""
    import os
    import random
    from hoisdf_b200 import feed
    from oracle import feed_oracle as FO
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "feed_seed31.npz"))
    seed = int(g["seed"])
    state = random.getstate()
    random.seed(seed)                                       # Python's Mersenne Twister stream is frozen across versions
    radius = random.random() * 0.5
    steps = feed.draw_color_jitter(brightness=0.5, contrast=0.5, saturation=0.5, hue=0.15)
    random.setstate(state)
    assert [n for n, _ in steps] == [str(n) for n in g["filt_order"]]
    assert np.array_equal(np.array([f for _, f in steps]), g["filt_factors"]) and radius == float(g["filt_radius"])
    got = emu_training_image(emu_jitter, FO.synthetic_aug(seed)[0], g["filt_center"], float(g["filt_scale"]),
                             float(g["filt_rot"]), radius, steps)
    assert np.array_equal(got[:, ::8], g["filt_img_rows"])


def test_randomized_fused_training_images_equal_the_step_route(emu_jitter):
    """`hoisdf_train_image_fwd` at 60 seeded draws of frame size, crop size (row pitches with and without padding), warp
    coefficients, mirror flag, blur radius (every box radius < 1) and jitter sequence against the emulated step-by-step route
    (crop -> blur -> jitter -> float), bytes and floats."""
    import itertools
    from test_feed import emu_warp, _random_coefficients
    from hoisdf_b200 import feed
    lib = emu_jitter
    vp, i64 = C.c_void_p, C.c_int64
    lib.hoisdf_image_crop_fwd.argtypes = [vp, i64, i64, i64, i64, i64, i64, vp, vp, i64, C.c_float, vp, vp, vp, vp]
    lib.hoisdf_train_image_fwd.argtypes = [vp, i64, i64, i64, i64, i64, vp, vp, vp, vp, vp, i64, vp, vp, vp]
    rng = np.random.default_rng(77)
    orders = list(itertools.permutations([BRIGHTNESS, SATURATION, HUE, CONTRAST]))
    n = 0
    for _ in range(60):
        h, w, res = int(rng.integers(2, 40)), int(rng.integers(2, 40)), int(rng.integers(1, 36))
        img = rng.integers(0, 256, (1, h, w, 3), dtype=np.uint8)
        coef = _random_coefficients(rng, h, w)
        if (coef[1] != 0 or coef[3] != 0) and not feed._fixed_point_ok(coef, res):
            continue
        n += 1
        mirror = np.array([int(rng.integers(0, 2))], np.int32)
        radius = float(rng.choice([0.0, rng.uniform(0, 0.5), rng.uniform(0.5, 1.3)]))
        blur = blur_params(lib, [radius])
        assert blur[0, 0] == 0
        seq = [(op, float(rng.uniform(-0.5, 0.5) if op == HUE else rng.uniform(0.3, 1.7)))
               for op in orders[int(rng.integers(0, 24))]][:int(rng.integers(0, 5))]
        ops = np.zeros((1, 4), np.int32)
        fac = np.zeros((1, 4), np.float32)
        for j, (op, f) in enumerate(seq):
            ops[0, j], fac[0, j] = op, (float(np.int32(f * 255).astype(np.uint8)) if op == HUE else f)
        of = np.full((1, 3, res, res), np.nan, np.float32)
        ou = np.full((1, res, res, 3), 9, np.uint8)
        c = np.ascontiguousarray(coef[None], dtype=np.float64)
        rc = lib.hoisdf_train_image_fwd(img.ctypes.data, 1, h, w, w * 3, h * w * 3, c.ctypes.data, mirror.ctypes.data,
                                        blur.ctypes.data, ops.ctypes.data, fac.ctypes.data, res, of.ctypes.data, ou.ctypes.data, None)
        assert rc == 0
        _, warped = emu_warp(lib, img, c, res, mirror=mirror)
        want = emu_jitter_run(lib, emu_blur(lib, warped, [radius]), [seq])
        assert np.array_equal(ou, want), (h, w, res, coef, radius, seq)
        assert np.array_equal(of[0], want[0].astype(np.float32).transpose(2, 0, 1) / np.float32(255.0))
    assert n > 45
