"""Data feed, GPU-less half:
  * oracle/feed_oracle.py against the UNMODIFIED upstream functions where /root/reference is mounted, and against the committed
    fixture tests/golden/feed_seed31.npz (generated from the upstream functions by oracle/make_golden.py:feed_case) everywhere;
  * the host geometry of hoisdf_b200/feed.py against the oracle (bit-equal float32 / float64 results);
  * csrc/feed.cu executed unchanged on the CPU emulator against Pillow itself: evaluation crops (scale-only path), rotated
    training warps (16.16 fixed-point path), mode-"L" masks with the NEAREST shrink, ragged sizes and windows that leave the
    frame (zero fill) -- every comparison bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest
from PIL import Image

from hoisdf_b200 import feed
from oracle import feed_oracle as FO
from oracle import reference_shim as rs
from test_kernel_emulation import build_emulated

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "feed_seed31.npz")


@pytest.fixture(scope="module")
def emu():
    lib = build_emulated("feed")
    vp, i64 = C.c_void_p, C.c_int64
    lib.hoisdf_image_crop_fwd.argtypes = [vp, i64, i64, i64, i64, i64, i64, vp, vp, i64, C.c_float, vp, vp, vp, vp]
    lib.hoisdf_image_crop_fwd.restype = C.c_int
    return lib


def emu_warp(lib, frames, coef, size, divisor=255.0, mirror=None):
    """frames (B, H, W, C) uint8 -> (out_f32 (B, C, size, size), out_u8 (B, size, size, C)) from ONE call with both outputs."""
    frames = np.ascontiguousarray(frames)
    b, h, w, ch = frames.shape
    coef = np.ascontiguousarray(coef, dtype=np.float64).reshape(b, 6)
    of = np.full((b, ch, size, size), np.nan, np.float32)
    ou = np.full((b, size, size, ch), 7, np.uint8)
    tables = np.zeros((b, 2, size), np.int32)
    mirror = None if mirror is None else np.ascontiguousarray(mirror, dtype=np.int32)
    rc = lib.hoisdf_image_crop_fwd(frames.ctypes.data, b, h, w, ch, w * ch, h * w * ch, coef.ctypes.data,
                                   None if mirror is None else mirror.ctypes.data, size, divisor,
                                   of.ctypes.data, ou.ctypes.data, tables.ctypes.data, None)
    assert rc == 0
    return of, ou


def pil_warp(frame, coef, size):
    img = Image.fromarray(frame if frame.shape[2] == 3 else frame[:, :, 0])
    out = np.asarray(img.transform((size, size), Image.AFFINE, tuple(float(c) for c in coef)))
    return out.reshape(size, size, -1)


# ---------------------------------------------------------------------------------------------- oracle pinning
@pytest.mark.skipif(not rs.available(), reason="upstream reference not mounted")
def test_feed_oracle_matches_upstream_live():
    import torchvision.transforms as T
    mods = rs.load_data_modules()
    DU, H = mods["dataset_util"], mods["ho3d"]

    class Self:
        inp_res = 256

    for seed in range(12):
        img, K, bh, p2d = FO.synthetic_frame(seed)
        im, K_ref, hand_ref, obj_ref = H.Dataset.data_crop(Self(), Image.fromarray(img), K, bh, p2d)
        ref = (T.ToTensor()(np.asarray(im).astype(np.float32)) / 255.0).numpy()           # ho3d.py:624
        got, K_got, hand_got, obj_got = FO.data_crop(img, K, bh, p2d)
        assert np.array_equal(ref, got) and got.dtype == np.float32
        assert np.array_equal(K_ref, K_got) and K_ref.dtype == K_got.dtype
        assert np.array_equal(hand_ref.astype(np.float32), hand_got) and np.array_equal(obj_ref.astype(np.float32), obj_got)
    for seed in range(6):
        img, hs, os_, center, scale, rot = FO.synthetic_aug(seed)
        affine, _ = DU.get_affine_transform(center, scale, [256, 256], rot=rot)
        ref_img = DU.transform_img(Image.fromarray(img), affine, [256, 256]).crop((0, 0, 256, 256))
        ref_seg = DU.transform_img(Image.fromarray(hs), affine, [256, 256]).crop((0, 0, 256, 256))
        ref_seg = np.asarray(ref_seg.resize((128, 128), Image.NEAREST)).astype(np.float32)
        pil_bytes, tensor, hand_seg, _, aff = FO.aug_warp(img, hs, os_, center, scale, rot)
        assert np.array_equal(affine, aff)
        assert np.array_equal(np.asarray(ref_img), pil_bytes) and np.array_equal(ref_seg, hand_seg)


def test_feed_oracle_matches_golden():
    g = np.load(GOLDEN)
    n_eval, n_aug = int(g["n_eval"]), int(g["n_aug"])
    for i in range(n_eval):
        img, K, bh, p2d = FO.synthetic_frame(int(g["seed"]) + i)
        got = FO.data_crop(img, K, bh, p2d)
        assert np.array_equal(g["u8_to_f32"][g["eval_bytes"][i]].transpose(2, 0, 1), got[0])
        for name, val in zip(("eval_K", "eval_bbox_hand", "eval_bbox_obj"), got[1:]):
            assert np.array_equal(g[name][i], val), (name, i)
    for i in range(n_aug):
        img, hs, os_, center, scale, rot = FO.synthetic_aug(int(g["seed"]) + i)
        pil_bytes, _, hand_seg, obj_seg, aff = FO.aug_warp(img, hs, os_, center, scale, rot)
        assert np.array_equal(g["aug_bytes"][i], pil_bytes) and np.array_equal(g["aug_affine"][i], aff)
        assert np.array_equal(g["aug_hand_seg"][i], hand_seg) and np.array_equal(g["aug_obj_seg"][i], obj_seg)


# ---------------------------------------------------------------------------------------------- host geometry
def test_host_geometry_is_bit_equal_to_the_oracle():
    frames = [FO.synthetic_frame(s) for s in range(40, 56)]
    K = np.stack([f[1] for f in frames])
    bh = np.stack([f[2] for f in frames])
    p2d = np.stack([f[3] for f in frames])
    coef, meta = feed.crop_geometry(K, bh, p2d, (640, 480), 256)
    for i, (img, k, b, p) in enumerate(frames):
        _, K_ref, hand_ref, obj_ref = FO.data_crop(img, k, b, p)
        assert np.array_equal(meta["cam_intr"][i], K_ref)
        assert np.array_equal(meta["bbox_hand"][i], hand_ref) and np.array_equal(meta["bbox_obj"][i], obj_ref)
        centre, scale = FO.fuse_bbox(FO.get_bbox_joints(b.reshape(2, 2), 1.5), FO.get_bbox_joints(p, 1.5), (640, 480))
        aff, _ = FO.get_affine_transform(centre, scale, [256, 256])
        inv = np.linalg.inv(aff)
        assert np.array_equal(coef[i], np.array([inv[0, 0], inv[0, 1], inv[0, 2], inv[1, 0], inv[1, 1], inv[1, 2]], np.float64))
        assert coef[i, 1] == 0.0 and coef[i, 3] == 0.0                 # the evaluation crop takes Pillow's scale-only path
    for s in range(6):
        _, _, _, center, scale, rot = FO.synthetic_aug(s)
        want, _ = FO.get_affine_transform(center, scale, [256, 256], rot=rot)
        assert np.array_equal(feed.crop_affine(center, scale, 256, rot), want)
    assert np.array_equal(feed.resize_coefficients(256, 128), [2.0, 0, 0, 0, 2.0, 0])


def test_crop_refuses_host_tensors_and_bad_layouts():
    import torch
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        feed.crop_images(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), np.array([[1.0, 0, 0, 0, 1, 0]]), 4)
    with pytest.raises(ValueError):
        feed.crop_images(torch.zeros(1, 8, 8, dtype=torch.uint8), np.array([[1.0, 0, 0, 0, 1, 0]]), 4)
    assert not feed._fixed_point_ok(np.array([300.0, 0.1, 0, 0, 1, 0]), 256)
    assert feed._fixed_point_ok(np.array([2.0, 0.1, -50, 0.1, 2.0, 30]), 256)


# ---------------------------------------------------------------------------------------------- kernel on the emulator
def test_evaluation_crop_kernel_is_bit_exact_with_pillow(emu):
    frames = [FO.synthetic_frame(s) for s in range(60, 64)]
    K = np.stack([f[1] for f in frames])
    bh = np.stack([f[2] for f in frames])
    p2d = np.stack([f[3] for f in frames])
    coef, _ = feed.crop_geometry(K, bh, p2d, (640, 480), 256)
    imgs = np.stack([f[0] for f in frames])
    of, ou = emu_warp(emu, imgs, coef, 256)
    for i, (img, k, b, p) in enumerate(frames):
        want = FO.data_crop(img, k, b, p)[0]
        assert np.array_equal(of[i], want)
        assert np.array_equal(ou[i], pil_warp(img, coef[i], 256))
    assert (ou == 0).all(axis=3).any()                      # some window leaves the frame: the zero fill was exercised


def test_rotated_warp_and_masks_are_bit_exact_with_pillow(emu):
    for s in range(4):
        img, hs, os_, center, scale, rot = FO.synthetic_aug(s)
        pil_bytes, tensor, hand_seg, obj_seg, aff = FO.aug_warp(img, hs, os_, center, scale, rot)
        coef = feed.pil_coefficients(feed.crop_affine(center, scale, 256, rot))[None]
        assert coef[0, 1] != 0.0 and feed._fixed_point_ok(coef[0], 256)
        of, ou = emu_warp(emu, img[None], coef, 256)
        assert np.array_equal(ou[0], pil_bytes) and np.array_equal(of[0], tensor)
        masks = np.stack([hs, os_])[:, :, :, None]
        _, warped = emu_warp(emu, masks, np.tile(coef, (2, 1)), 256, divisor=1.0)
        small, _ = emu_warp(emu, warped, np.tile(feed.resize_coefficients(256, 128), (2, 1)), 128, divisor=1.0)
        assert np.array_equal(small[0, 0], hand_seg) and np.array_equal(small[1, 0], obj_seg)


@pytest.mark.parametrize("h,w,size", [(37, 53, 19), (5, 7, 33), (480, 640, 1), (64, 64, 64)])
def test_ragged_sizes_and_out_of_frame_windows(emu, h, w, size):
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(1, 256, (3, h, w, 3), dtype=np.uint8)
    coef = np.array([[w / size * 1.37, 0, -0.31 * w, 0, h / size * 0.83, 0.4 * h],              # scale-only, partly outside
                     [0.91, -0.43, 0.2 * w, 0.43, 0.91, -0.3 * h],                              # rotation, partly outside
                     [1e-3, 0, w + 5.0, 0, 1e-3, 2.0]])                                         # entirely outside
    of, ou = emu_warp(emu, img, coef, size)
    for i in range(3):
        want = pil_warp(img[i], coef[i], size)
        assert np.array_equal(ou[i], want)
        assert np.array_equal(of[i], want.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0))
    assert not ou[2].any()


def test_argument_checks(emu):
    buf = np.zeros(64, np.uint8)
    coef = np.array([1.0, 0, 0, 0, 1, 0])
    t = np.zeros(8, np.int32)
    call = emu.hoisdf_image_crop_fwd
    assert call(None, 1, 2, 2, 3, 6, 12, coef.ctypes.data, None, 2, 255.0, buf.ctypes.data, None, t.ctypes.data, None) == -1
    assert call(buf.ctypes.data, 1, 2, 2, 3, 6, 12, coef.ctypes.data, None, 2, 255.0, None, None, t.ctypes.data, None) == -1
    assert call(buf.ctypes.data, 1, 2, 2, 2, 6, 12, coef.ctypes.data, None, 2, 255.0, buf.ctypes.data, None, t.ctypes.data, None) == -2
    assert call(buf.ctypes.data, 1, 2, 2, 3, 5, 12, coef.ctypes.data, None, 2, 255.0, buf.ctypes.data, None, t.ctypes.data, None) == -2
    assert call(buf.ctypes.data, 1, 2, 2, 3, 6, 12, coef.ctypes.data, None, 2, 0.0, buf.ctypes.data, None, t.ctypes.data, None) == -2


# ---------------------------------------------------------------------------------------------- SDF point sets
N_HAND, N_OBJ = 24, 8


def _item_from_fixture(g):
    """The oracle's restatement of the fixture's training item (frame / masks / SDF file regenerated from the seed)."""
    seed = int(g["seed"])
    sdf = FO.synthetic_sdf_frame(seed, N_HAND, N_OBJ)[0]
    frame, hand_mask, obj_mask = FO.synthetic_aug(seed)[:3]
    return sdf, frame, hand_mask, obj_mask


@pytest.mark.skipif(not rs.available(), reason="upstream reference not mounted")
def test_feed_oracle_matches_upstream_training_item_live():
    """The UNMODIFIED upstream `Dataset.__getitem__` (mode "train") on synthetic files against the oracle: image, both
    masks, the four point sets and the two SDF targets, bit for bit."""
    for seed in (1, 2, 6):
        inputs, targets, meta, taps = rs.ho3d_train_item(seed, N_HAND, N_OBJ)
        a = taps["affine"][0]
        np.random.seed(seed)                       # the product's draws after the same seed are upstream's draws
        assert np.array_equal(feed.draw_sdf_indices(taps["sdf"], taps["n_hand_rows"], N_HAND, N_OBJ, 0.02),
                              np.concatenate(taps["draws"]))
        got_in, got_t = FO.sdf_point_sets(taps["sdf"], np.concatenate(taps["draws"]), N_HAND, N_OBJ, meta["mano_root"],
                                          meta["obj_center_cam"], taps["hand_sdf_scale"], taps["obj_sdf_scale"],
                                          rot_mat=a["rot_mat"])
        for k, v in got_in.items():
            assert np.array_equal(v, inputs[k]), k
        for k, v in got_t.items():
            assert np.array_equal(v, targets[k]), k
        _, tensor, hs, os_, aff = FO.aug_warp(taps["frame"], taps["hand_mask"], taps["obj_mask"], a["center"], a["scale"],
                                              a["rot"])
        assert np.array_equal(aff, a["affinetrans"]) and np.array_equal(tensor, inputs["img"].numpy())
        assert np.array_equal(hs, targets["hand_seg"].numpy()) and np.array_equal(os_, targets["obj_seg"].numpy())


def test_feed_oracle_matches_golden_training_item():
    g = np.load(GOLDEN)
    sdf, frame, hand_mask, obj_mask = _item_from_fixture(g)
    state = np.random.get_state()
    np.random.seed(int(g["seed"]))                 # numpy's legacy generator is a frozen stream: upstream's draws, from the fixture
    assert np.array_equal(feed.draw_sdf_indices(sdf, FO.synthetic_sdf_frame(int(g["seed"]), N_HAND, N_OBJ)[1], N_HAND, N_OBJ, 0.02),
                          g["item_draws"])
    np.random.set_state(state)
    got_in, got_t = FO.sdf_point_sets(sdf, g["item_draws"], N_HAND, N_OBJ, g["item_mano_root"], g["item_obj_center_cam"],
                                      float(g["item_hand_sdf_scale"]), float(g["item_obj_sdf_scale"]),
                                      rot_mat=g["item_rot_mat"])
    for k, v in got_in.items():
        assert np.array_equal(v, g["item_" + k]), k
    assert np.array_equal(got_t["hand_sdf"], g["item_hand_sdf"]) and np.array_equal(got_t["obj_sdf"], g["item_obj_sdf"])
    pil_bytes, _, hs, os_, _ = FO.aug_warp(frame, hand_mask, obj_mask, g["item_center"], float(g["item_scale"]),
                                           float(g["item_rot"]))
    assert np.array_equal(pil_bytes, g["item_img_bytes"])
    assert np.array_equal(hs, g["item_hand_seg"]) and np.array_equal(os_, g["item_obj_seg"])


@pytest.fixture(scope="module")
def emu_rows(emu):
    vp, i64, f = C.c_void_p, C.c_int64, C.c_float
    emu.hoisdf_sdf_rows_fwd.argtypes = [vp, vp, vp, i64, i64, i64, i64, vp, vp, vp, vp, f, f, vp, vp, vp, vp, vp, vp, vp, vp]
    emu.hoisdf_sdf_rows_fwd.restype = C.c_int
    return emu


def emu_point_sets(lib, frames, n_hand, n_obj, hand_scale, obj_scale, use_rot=True, flip=None):
    """frames: list of (rows, all_idx, rot, root, centre) -> dict of outputs + status, one call for the batch."""
    b = len(frames)
    rows = np.ascontiguousarray(np.concatenate([f[0] for f in frames]))
    offsets = np.cumsum([0] + [len(f[0]) for f in frames]).astype(np.int64)
    index = np.ascontiguousarray(np.stack([f[1] for f in frames]).astype(np.int64))
    rot = np.ascontiguousarray(np.stack([f[2] for f in frames]).astype(np.float32)) if use_rot else None
    root = np.ascontiguousarray(np.stack([f[3] for f in frames]).astype(np.float32))
    centre = np.ascontiguousarray(np.stack([f[4] for f in frames]).astype(np.float32))
    n_sel = index.shape[1]
    pre = n_sel == 2 * (n_hand + n_obj)
    out = {"hand_sdf_points": np.full((b, n_hand, 3), np.nan, np.float32), "obj_sdf_points": np.full((b, n_obj, 3), np.nan, np.float32),
           "hand_pre_points": np.full((b, n_hand, 3), np.nan, np.float32) if pre else None,
           "obj_pre_points": np.full((b, n_obj, 3), np.nan, np.float32) if pre else None,
           "hand_sdf": np.full((b, n_hand), np.nan, np.float32), "obj_sdf": np.full((b, n_obj), np.nan, np.float32)}
    status = np.zeros(1, np.int32)
    p = lambda a: None if a is None else a.ctypes.data                   # noqa: E731
    fl = None if flip is None else np.ascontiguousarray(flip, dtype=np.int32)
    rc = lib.hoisdf_sdf_rows_fwd(p(rows), p(offsets), p(index), b, n_sel, n_hand, n_obj, p(rot), p(fl), p(root), p(centre),
                                 hand_scale, obj_scale, p(out["hand_sdf_points"]), p(out["obj_sdf_points"]),
                                 p(out["hand_pre_points"]), p(out["obj_pre_points"]), p(out["hand_sdf"]), p(out["obj_sdf"]),
                                 p(status), None)
    return rc, out, int(status[0])


def close32(got, want):
    """float32 results of at most three fused operations: equal to numpy's up to the last bit of the largest operand
    (bit-equal on the build container's OpenBLAS; another sgemm kernel may order the three products differently)."""
    return float(np.abs(got - want).max()) <= 4 * np.finfo(np.float32).eps * max(1.0, float(np.abs(want).max()))


@pytest.mark.parametrize("train,use_rot,flip", [(True, True, False), (False, False, False), (False, False, True), (True, True, True)])
def test_sdf_rows_kernel_vs_oracle(emu_rows, train, use_rot, flip):
    n_hand, n_obj, hs, os_ = 37, 11, 6.2, 5.8
    frames = []
    for s in range(5):
        rows, _, idx, rot, root, centre = FO.synthetic_sdf_frame(s, n_hand, n_obj, train=train)
        frames.append((rows, idx, rot, root, centre))
    flips = np.array([1, 0, 1, 1, 0], np.int32) if flip else None
    rc, out, status = emu_point_sets(emu_rows, frames, n_hand, n_obj, hs, os_, use_rot, flips)
    assert rc == 0 and status == 0
    exact = True
    for i, (rows, idx, rot, root, centre) in enumerate(frames):
        want_in, want_t = FO.sdf_point_sets(rows, idx, n_hand, n_obj, root, centre, hs, os_, rot_mat=rot if use_rot else None,
                                            do_flip=bool(flip and flips[i]))
        for k, v in {**want_in, **want_t}.items():
            assert close32(out[k][i], v), (k, i)
            exact &= np.array_equal(out[k][i], v)
    if not train:
        assert out["hand_pre_points"] is None
    if not use_rot:
        assert exact                           # without the sgemm the arithmetic is two separately rounded operations


def test_sdf_rows_kernel_reproduces_upstream_item(emu_rows):
    g = np.load(GOLDEN)
    sdf = _item_from_fixture(g)[0]
    frame = (sdf, g["item_draws"], g["item_rot_mat"], g["item_mano_root"], g["item_obj_center_cam"])
    rc, out, status = emu_point_sets(emu_rows, [frame], N_HAND, N_OBJ, float(g["item_hand_sdf_scale"]),
                                     float(g["item_obj_sdf_scale"]))
    assert rc == 0 and status == 0
    for k in ("hand_sdf_points", "obj_sdf_points", "hand_pre_points", "obj_pre_points", "hand_sdf", "obj_sdf"):
        assert close32(out[k][0], g["item_" + k]), k


def test_sdf_rows_kernel_flags_bad_indices_and_arguments(emu_rows):
    rows, _, idx, rot, root, centre = FO.synthetic_sdf_frame(0, 8, 4, train=False)
    bad = idx.copy()
    bad[3] = len(rows)
    rc, out, status = emu_point_sets(emu_rows, [(rows, bad, rot, root, centre)], 8, 4, 1.0, 1.0, False)
    assert rc == 0 and status == 1 and np.isnan(out["hand_sdf_points"][0, 3]).all()
    assert not np.isnan(out["hand_sdf_points"][0, :3]).any()
    rc, _, _ = emu_point_sets(emu_rows, [(rows, idx[:-1], rot, root, centre)], 8, 4, 1.0, 1.0, False)
    assert rc == -2


def test_mirrored_warp_equals_warping_the_mirrored_frame(emu):
    """data/dexycb.py:427-430,479-481: left-hand frames and masks are flipped left-right before the warp."""
    for s in range(3):
        img, hs, _, center, scale, rot = FO.synthetic_aug(20 + s)
        rot = rot if s else 0.0                                                # s = 0: the scale-only path
        coef = feed.pil_coefficients(feed.crop_affine(center, scale, 256, rot))[None]
        _, got = emu_warp(emu, img[None], coef, 256, mirror=[1])
        assert np.array_equal(got[0], pil_warp(np.ascontiguousarray(img[:, ::-1, :]), coef[0], 256))
        _, plain = emu_warp(emu, img[None], coef, 256, mirror=[0])
        assert np.array_equal(plain[0], pil_warp(img, coef[0], 256))
        _, got = emu_warp(emu, hs[None, :, :, None], coef, 256, divisor=1.0, mirror=[1])
        assert np.array_equal(got[0], pil_warp(np.ascontiguousarray(hs[:, ::-1])[:, :, None], coef[0], 256))


@pytest.mark.skipif(not rs.available(), reason="upstream reference not mounted")
def test_dexycb_crop_geometry_and_warp_match_upstream_live(emu):
    """The UNMODIFIED `data.dexycb.Dataset.data_crop` (dexycb.py:355-404) against the host geometry + the emulated kernel:
    every returned array bit-equal (image, boxes, intrinsics, heat-map joints, normalised corners, both masks)."""
    rs.load_data_modules()
    import data.dexycb as D

    class Self:
        inp_res, heatmap_res = 256, 128

    for seed in range(70, 76):
        img, K32, _, p2d = FO.synthetic_frame(seed)
        _, hs, os_, _, _, _ = FO.synthetic_aug(seed)
        rng = np.random.default_rng(seed)
        K = K32.astype(np.float64)                                            # dexycb.py:421-426 builds K in float64
        uv = (p2d.mean(0) + rng.uniform(-60, 60, (21, 2))).astype(np.float32)
        ref = D.Dataset.data_crop(Self(), Image.fromarray(img), K, uv, p2d, Image.fromarray(hs), Image.fromarray(os_))
        r_img, r_hand, r_obj, r_K, r_uv, r_p2d, r_hs, r_os = ref
        coef, meta = feed.crop_geometry_dexycb(K[None], uv[None], p2d[None], (640, 480), 256, 128)
        assert np.array_equal(meta["bbox_hand"][0], r_hand) and np.array_equal(meta["bbox_obj"][0], r_obj)
        assert np.array_equal(meta["cam_intr"][0], r_K) and meta["cam_intr"].dtype == r_K.dtype
        assert np.array_equal(meta["joints_uv"][0], r_uv) and np.array_equal(meta["p2d"][0], r_p2d)
        _, got = emu_warp(emu, img[None], coef, 256)
        assert np.array_equal(got[0], np.asarray(r_img))
        _, warped = emu_warp(emu, np.stack([hs, os_])[:, :, :, None], np.tile(coef, (2, 1)), 256, divisor=1.0)
        small, _ = emu_warp(emu, warped, np.tile(feed.resize_coefficients(256, 128), (2, 1)), 128, divisor=1.0)
        assert np.array_equal(small[0, 0], r_hs.astype(np.float32)) and np.array_equal(small[1, 0], r_os.astype(np.float32))


def test_dexycb_crop_matches_golden(emu):
    g = np.load(GOLDEN)
    seed = int(g["seed"])
    img, K32, _, p2d = FO.synthetic_frame(seed)
    _, hs, os_, _, _, _ = FO.synthetic_aug(seed)
    coef, meta = feed.crop_geometry_dexycb(K32.astype(np.float64)[None], g["dex_uv_in"][None], p2d[None], (640, 480), 256, 128)
    for k, name in (("bbox_hand", "dex_bbox_hand"), ("bbox_obj", "dex_bbox_obj"), ("cam_intr", "dex_K"),
                    ("joints_uv", "dex_joints_uv"), ("p2d", "dex_p2d")):
        assert np.array_equal(meta[k][0], g[name]), k
    _, got = emu_warp(emu, img[None], coef, 256)
    assert np.array_equal(got[0][::32], g["dex_img_rows"])
    _, warped = emu_warp(emu, np.stack([hs, os_])[:, :, :, None], np.tile(coef, (2, 1)), 256, divisor=1.0)
    small, _ = emu_warp(emu, warped, np.tile(feed.resize_coefficients(256, 128), (2, 1)), 128, divisor=1.0)
    assert np.array_equal(small[0, 0], g["dex_hand_seg"]) and np.array_equal(small[1, 0], g["dex_obj_seg"])


# ---------------------------------------------------------------------------------------------- training sample, host side
def product_sample(seed, n_hand=N_HAND, n_obj=N_OBJ):
    """The host half of one training sample as the product computes it: seeds both generators like the upstream run, then
    draws in upstream's order (point indices, geometry, blur radius, jitter)."""
    import random
    ann = FO.synthetic_annotation(seed)
    sdf, nh = FO.synthetic_sdf_frame(seed, n_hand, n_obj)[:2]
    np.random.seed(seed)
    random.seed(seed)
    index = feed.draw_sdf_indices(sdf, nh, n_hand, n_obj, 0.02)
    centre, scale = feed.fuse_boxes(feed.bbox_from_points(ann["joints_uv"], 1.5), feed.bbox_from_points(ann["obj_p2d"], 1.5),
                                    (640, 480))
    centre, scale, rot = feed.draw_train_geometry(centre, scale)
    sample = feed.train_geometry(ann["cam_intr"], ann["joints_uv"], ann["joints_3d"], ann["mano_param"], ann["obj_p2d"],
                                 ann["obj_p3d"], ann["obj_rot"], ann["obj_trans"], centre, scale, rot,
                                 ann["obj_depth_mean_value"], obj_name="003_cracker_box")
    sample.update(index=index, blur_radius=random.random() * 0.5,
                  jitter=feed.draw_color_jitter(brightness=0.5, contrast=0.5, saturation=0.5, hue=0.15))
    return sample, sdf


TARGET_KEYS = ("joint_coord", "joint_cam_no_trans", "obj_rot", "rel_obj_trans", "mano_param")
META_KEYS = ("cam_intr", "mano_root", "obj_center_cam", "bbox_hand", "bbox_obj")


@pytest.mark.skipif(not rs.available(), reason="upstream reference not mounted")
def test_training_sample_host_geometry_matches_upstream_live():
    """`train_geometry` + the product's draws against the unmodified `Dataset.__getitem__`: every non-pixel entry of the item
    (targets and meta_info) equal in value AND dtype; the draws equal upstream's."""
    state = np.random.get_state()
    for seed in (0, 5, 13):
        inputs, targets, meta, taps = rs.ho3d_train_item(seed, filters=True)
        sample, _ = product_sample(seed)
        a = taps["affine"][0]
        assert np.array_equal(sample["index"], np.concatenate(taps["draws"])) and np.array_equal(sample["rot_mat"], a["rot_mat"])
        for k in TARGET_KEYS:
            assert np.array_equal(sample[k], targets[k]) and sample[k].dtype == targets[k].dtype, k
        for k in META_KEYS:
            assert np.array_equal(sample[k], meta[k]) and sample[k].dtype == meta[k].dtype, k
        assert sample["obj_mask"] == meta["obj_mask"]
    np.random.set_state(state)


def test_training_sample_host_geometry_matches_golden():
    g = np.load(GOLDEN)
    state = np.random.get_state()
    sample, _ = product_sample(int(g["seed"]))
    np.random.set_state(state)
    for k in TARGET_KEYS:
        assert np.array_equal(sample[k], g["filt_t_" + k]), k
    for k in META_KEYS:
        assert np.array_equal(sample[k], g["filt_m_" + k]), k
    assert sample["blur_radius"] == float(g["filt_radius"]) and [n for n, _ in sample["jitter"]] == [str(n) for n in g["filt_order"]]


# ---------------------------------------------------------------------------------------------- evaluation sample
EVAL_META = ("cam_intr", "mano_root", "obj_center_cam", "bbox_hand", "bbox_obj")


@pytest.mark.skipif(not rs.available(), reason="upstream reference not mounted")
def test_evaluation_sample_matches_upstream_live(emu):
    """The unmodified `Dataset.__getitem__` in evaluation mode (what main/test.py's loader yields) on a synthetic sequence
    directory against `eval_geometry` + the emulated crop kernel: image, targets and meta_info, values and dtypes."""
    for seed in range(4):
        inputs, targets, meta = rs.ho3d_eval_item(seed)
        img, ann, corners = FO.synthetic_eval_annotation(seed)
        g = feed.eval_geometry(ann, corners, (640, 480), 0.7)
        for k in ("obj_rot", "rel_obj_trans"):
            assert np.array_equal(g[k], targets[k]) and g[k].dtype == targets[k].dtype, k
        for k in EVAL_META:
            assert np.array_equal(g[k], meta[k]) and g[k].dtype == meta[k].dtype, k
        assert g["obj_mask"] == meta["obj_mask"] and g["obj_cls"] == meta["obj_cls"]
        got, _ = emu_warp(emu, img[None], g["coef"][None], 256)
        assert np.array_equal(got[0], inputs["img"].numpy())


def test_evaluation_sample_matches_golden(emu):
    g = np.load(GOLDEN)
    img, ann, corners = FO.synthetic_eval_annotation(int(g["seed"]))
    s = feed.eval_geometry(ann, corners, (640, 480), 0.7)
    for k in ("obj_rot", "rel_obj_trans") + EVAL_META:
        assert np.array_equal(s[k], g["evi_" + k]), k
    assert s["obj_mask"] == bool(g["evi_obj_mask"]) and s["obj_cls"] == str(g["evi_obj_cls"])
    got, _ = emu_warp(emu, img[None], s["coef"][None], 256)
    assert np.array_equal(got[0][:, ::8], g["evi_img_rows"])


# ---------------------------------------------------------------------------------------------- DexYCB test sample
DEX_TARGETS = ("joint_coord", "joint_cam_no_trans", "obj_rot", "rel_obj_trans", "mano_param")


def dexycb_product_sample(seed, left=None, n_hand=N_HAND, n_obj=N_OBJ):
    img, hm, om, info, hold = FO.synthetic_dexycb_sample(seed, left)
    sdf, nh = FO.synthetic_sdf_frame(seed, n_hand, n_obj)[:2]
    s = feed.dexycb_eval_geometry(info, hold["components_right"], hold["components_left"], hold["handmean"],
                                  hold["obj_bbox3d"][info["ycb_ids"][1]], (640, 480))
    state = np.random.get_state()
    np.random.seed(seed)
    s["index"] = feed.draw_sdf_indices(sdf, nh, n_hand, n_obj)
    np.random.set_state(state)
    return s, img, hm, om, sdf


def emu_dexycb_pixels(lib, s, img, hm, om):
    mirror = [int(s["flip"])]
    as_float, _ = emu_warp(lib, img[None], s["coef"][None], 256, mirror=mirror)
    _, warped = emu_warp(lib, np.stack([hm, om])[:, :, :, None], np.tile(s["coef"], (2, 1)), 256, divisor=1.0, mirror=mirror * 2)
    small, _ = emu_warp(lib, warped, np.tile(feed.resize_coefficients(256, 128), (2, 1)), 128, divisor=1.0)
    return as_float[0], small[0, 0], small[1, 0]


@pytest.mark.skipif(not rs.available(), reason="upstream reference not mounted")
def test_dexycb_test_sample_matches_upstream_live(emu):
    """The unmodified `dexycb.Dataset.__getitem__` in test mode (BASELINE configs[2]'s feed), right AND left hands (the mirror
    path), against `dexycb_eval_geometry`, the product's draws and the emulated kernels: every entry, values and dtypes."""
    for seed in range(4):
        inputs, targets, meta, taps = rs.dexycb_test_item(seed)
        s, img, hm, om, sdf = dexycb_product_sample(seed)
        assert s["flip"] == bool(seed % 2) and np.array_equal(s["index"], np.concatenate(taps["draws"]))
        for k in DEX_TARGETS:
            assert np.array_equal(s[k], targets[k]) and s[k].dtype == targets[k].dtype, k
        for k in EVAL_META:
            assert np.array_equal(s[k], meta[k]) and s[k].dtype == meta[k].dtype, k
        assert s["obj_cls"] == meta["obj_cls"] and inputs["hand_pre_points"] is False
        got_img, got_hs, got_os = emu_dexycb_pixels(emu, s, img, hm, om)
        assert np.array_equal(got_img, inputs["img"].numpy())
        assert np.array_equal(got_hs, targets["hand_seg"].numpy()) and np.array_equal(got_os, targets["obj_seg"].numpy())
        want_in, want_t = FO.sdf_point_sets(sdf, s["index"], N_HAND, N_OBJ, s["mano_root"], s["obj_center_cam"], 6.2, 5.8,
                                            do_flip=s["flip"])
        for k in ("hand_sdf_points", "obj_sdf_points"):
            assert np.array_equal(want_in[k], inputs[k]), k
        assert np.array_equal(want_t["hand_sdf"], targets["hand_sdf"]) and np.array_equal(want_t["obj_sdf"], targets["obj_sdf"])


def test_dexycb_test_sample_matches_golden(emu):
    g = np.load(GOLDEN)
    s, img, hm, om, sdf = dexycb_product_sample(int(g["seed"]), left=True)
    assert np.array_equal(s["index"], g["dxi_draws"]) and s["flip"] and s["obj_cls"] == int(g["dxi_obj_cls"])
    for k in DEX_TARGETS + EVAL_META:
        assert np.array_equal(s[k], g["dxi_" + k]), k
    got_img, got_hs, got_os = emu_dexycb_pixels(emu, s, img, hm, om)
    assert np.array_equal(got_img[:, ::8], g["dxi_img_rows"])
    assert np.array_equal(got_hs, g["dxi_hand_seg"]) and np.array_equal(got_os, g["dxi_obj_seg"])


# ---------------------------------------------------------------------------------------------- DexYCB training sample
def dexycb_train_product_sample(seed, left=None, n_hand=N_HAND, n_obj=N_OBJ):
    """Host half of one DexYCB training sample as the product computes it, draws in upstream's order after the same seeds."""
    import random
    img, hm, om, info, hold = FO.synthetic_dexycb_sample(seed, left)
    sdf, nh = FO.synthetic_sdf_frame(seed, n_hand, n_obj)[:2]
    state = np.random.get_state()
    np.random.seed(seed)
    random.seed(seed)
    index = feed.draw_sdf_indices(sdf, nh, n_hand, n_obj, 0.02)
    s = feed.dexycb_train_geometry(info, hold["components_right"], hold["components_left"], hold["handmean"],
                                   hold["obj_bbox3d"][info["ycb_ids"][1]], (640, 480))
    s.update(index=index, blur_radius=random.random() * 0.5,
             jitter=feed.draw_color_jitter(brightness=0.5, contrast=0.5, saturation=0.5, hue=0.15))
    np.random.set_state(state)
    return s, img, hm, om, sdf


@pytest.mark.skipif(not rs.available(), reason="upstream reference not mounted")
def test_dexycb_training_sample_host_geometry_matches_upstream_live():
    for seed in range(6):
        inputs, targets, meta, taps = rs.dexycb_test_item(seed, mode="train", filters=True)
        s = dexycb_train_product_sample(seed)[0]
        assert np.array_equal(s["index"], np.concatenate(taps["draws"]))
        for k in DEX_TARGETS:
            assert np.array_equal(s[k], targets[k]) and s[k].dtype == targets[k].dtype, k
        for k in EVAL_META:
            assert np.array_equal(s[k], meta[k]) and s[k].dtype == meta[k].dtype, k


def test_dexycb_training_sample_host_geometry_matches_golden():
    g = np.load(GOLDEN)
    s = dexycb_train_product_sample(int(g["seed"]), left=True)[0]
    assert np.array_equal(s["index"], g["dxt_draws"]) and s["flip"]
    for k in DEX_TARGETS + EVAL_META:
        assert np.array_equal(s[k], g["dxt_" + k]), k


# ---------------------------------------------------------------------------------------------- randomized sweeps against Pillow
def _random_coefficients(rng, h, w):
    kind = int(rng.integers(0, 4))
    if kind == 0:           # scale-only, any sign / magnitude of the scales
        return np.array([rng.uniform(-3, 3), 0, rng.uniform(-w, 2 * w), 0, rng.uniform(-3, 3), rng.uniform(-h, 2 * h)])
    if kind == 1:           # degenerate scales and integer shifts
        pick = [0.0, 1.0, -1.0, 0.5, 2.0]
        return np.array([rng.choice(pick), 0, float(rng.integers(-5, w + 5)), 0, rng.choice(pick), float(rng.integers(-5, h + 5))])
    if kind == 2:           # general affine
        return rng.uniform(-2, 2, 6) * np.array([1, 1, w, 1, 1, h])
    return np.array([rng.uniform(-1e3, 1e3), rng.uniform(-1e-3, 1e-3), rng.uniform(-1e3, 1e3), rng.uniform(-1e-3, 1e-3),
                     rng.uniform(-1e3, 1e3), rng.uniform(-1e3, 1e3)])            # almost everything outside the frame


def test_randomized_warps_match_pillow(emu):
    """300 seeded draws of frame size, output size, channel count and coefficients (negative, zero and huge scales, shears,
    integer shifts): the emulated kernel against `Image.transform(..., AFFINE)`, every byte."""
    rng = np.random.default_rng(123)
    n = 0
    for _ in range(300):
        h, w, size, ch = int(rng.integers(1, 40)), int(rng.integers(1, 40)), int(rng.integers(1, 40)), int(rng.choice([1, 3]))
        img = rng.integers(1, 256, (1, h, w, ch), dtype=np.uint8)
        coef = _random_coefficients(rng, h, w)
        if (coef[1] != 0 or coef[3] != 0) and not feed._fixed_point_ok(coef, size):
            continue
        n += 1
        _, got = emu_warp(emu, img, coef[None], size)
        assert np.array_equal(got[0], pil_warp(img[0], coef, size)), (h, w, size, ch, coef)
    assert n > 250


def test_randomized_mask_crops_match_pillow(emu):
    """The one-launch mask kernel against `transform` + `resize(NEAREST)` for 120 seeded draws of sizes (incl. non-integer
    shrink ratios and out_res = res) and coefficients."""
    vp, i64 = C.c_void_p, C.c_int64
    emu.hoisdf_mask_crop_fwd.argtypes = [vp, i64, i64, i64, i64, i64, vp, vp, i64, i64, vp, vp]
    rng = np.random.default_rng(321)
    n = 0
    for _ in range(120):
        h, w = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        res = int(rng.integers(1, 48))
        out_res = int(rng.integers(1, res + 1))
        mask = np.ascontiguousarray(rng.integers(1, 256, (h, w), dtype=np.uint8))
        coef = _random_coefficients(rng, h, w)
        if (coef[1] != 0 or coef[3] != 0) and not feed._fixed_point_ok(coef, res):
            continue
        n += 1
        flip = np.array([int(rng.integers(0, 2))], np.int32)
        out = np.full((out_res, out_res), np.nan, np.float32)
        c = np.ascontiguousarray(coef, dtype=np.float64)
        assert emu.hoisdf_mask_crop_fwd(mask.ctypes.data, 1, h, w, w, h * w, c.ctypes.data, flip.ctypes.data, res, out_res,
                                        out.ctypes.data, None) == 0
        src = np.ascontiguousarray(mask[:, ::-1]) if flip[0] else mask
        pil = Image.fromarray(src).transform((res, res), Image.AFFINE, tuple(float(v) for v in coef))
        want = np.asarray(pil.resize((out_res, out_res), Image.NEAREST)).astype(np.float32)
        assert np.array_equal(out, want), (h, w, res, out_res, coef, flip)
    assert n > 90
