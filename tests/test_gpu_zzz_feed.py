"""Data feed on the GPU (csrc/feed.cu through the C ABI via hoisdf_b200/feed.py): bit-exact against
Pillow (the library upstream's dataset warps frames with, data/dataset_util.py:44-51) through the oracle's restatement of
`data_crop` / `data_aug` (oracle/feed_oracle.py) and against the committed upstream fixture tests/golden/feed_seed31.npz.
The same kernel source passes the same comparisons on the CPU emulator (tests/test_feed.py); this file sorts last so that it is
the final thing the -x run reaches."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import feed_oracle as FO

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "feed_seed31.npz")


def _batch(seeds):
    frames = [FO.synthetic_frame(s) for s in seeds]
    return frames, np.stack([f[0] for f in frames]), np.stack([f[1] for f in frames]), np.stack([f[2] for f in frames]), \
        np.stack([f[3] for f in frames])


def test_evaluation_crop_batch_vs_oracle_and_fixture(cuda):
    from hoisdf_b200 import feed
    g = np.load(GOLDEN)
    seeds = list(range(int(g["seed"]), int(g["seed"]) + 32))            # BASELINE configs[1] batch
    frames, imgs, K, bh, p2d = _batch(seeds)
    img, meta = feed.data_crop(torch.from_numpy(imgs).to(cuda), K, bh, p2d, 256)
    assert img.shape == (32, 3, 256, 256) and img.dtype == torch.float32
    got = img.cpu().numpy()
    for i, (f, k, b, p) in enumerate(frames):
        want, K_ref, hand_ref, obj_ref = FO.data_crop(f, k, b, p)
        assert np.array_equal(got[i], want), i
        assert np.array_equal(meta["cam_intr"][i], K_ref) and np.array_equal(meta["bbox_hand"][i], hand_ref)
        assert np.array_equal(meta["bbox_obj"][i], obj_ref)
    for i in range(int(g["n_eval"])):
        assert np.array_equal(got[i], g["u8_to_f32"][g["eval_bytes"][i]].transpose(2, 0, 1))
        # (host float geometry against a fixture made on another CPU: BLAS kernels may round the 3 x 3 products differently)
        assert np.allclose(meta["cam_intr"][i], g["eval_K"][i], rtol=1e-6, atol=1e-6)


def test_rotated_warp_and_masks_vs_oracle_and_fixture(cuda):
    from hoisdf_b200 import feed
    g = np.load(GOLDEN)
    seeds = list(range(int(g["seed"]), int(g["seed"]) + 8))
    draws = [FO.synthetic_aug(s) for s in seeds]
    coef = np.stack([feed.pil_coefficients(feed.crop_affine(c, sc, 256, r)) for _, _, _, c, sc, r in draws])
    frames = torch.from_numpy(np.stack([d[0] for d in draws])).to(cuda)
    as_bytes = feed.crop_images(frames, coef, 256, as_bytes=True).cpu().numpy()
    as_float = feed.crop_images(frames, coef, 256).cpu().numpy()
    hand = feed.crop_masks(torch.from_numpy(np.stack([d[1] for d in draws])).to(cuda), coef, 256, 128).cpu().numpy()
    obj = feed.crop_masks(torch.from_numpy(np.stack([d[2] for d in draws])).to(cuda), coef, 256, 128).cpu().numpy()
    for i, (img, hs, os_, c, sc, r) in enumerate(draws):
        pil_bytes, tensor, hand_seg, obj_seg, _ = FO.aug_warp(img, hs, os_, c, sc, r)
        assert np.array_equal(as_bytes[i], pil_bytes) and np.array_equal(as_float[i], tensor), i
        assert np.array_equal(hand[i], hand_seg) and np.array_equal(obj[i], obj_seg), i
    assert np.array_equal(as_bytes[0], g["aug_bytes"][0]) and np.array_equal(hand[0], g["aug_hand_seg"][0])
    assert np.array_equal(obj[0], g["aug_obj_seg"][0])


@pytest.mark.parametrize("h,w,size", [(37, 53, 19), (5, 7, 33), (480, 640, 1), (1080, 1920, 512)])
def test_ragged_sizes_and_out_of_frame_windows(cuda, h, w, size):
    from hoisdf_b200 import feed
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(1, 256, (3, h, w, 3), dtype=np.uint8)
    coef = np.array([[w / size * 1.37, 0, -0.31 * w, 0, h / size * 0.83, 0.4 * h],
                     [0.91, -0.43, 0.2 * w, 0.43, 0.91, -0.3 * h],
                     [1e-3, 0, w + 5.0, 0, 1e-3, 2.0]])
    got = feed.crop_images(torch.from_numpy(img).to(cuda), coef, size, as_bytes=True).cpu().numpy()
    for i in range(3):
        want = np.asarray(Image.fromarray(img[i]).transform((size, size), Image.AFFINE, tuple(float(c) for c in coef[i])))
        assert np.array_equal(got[i], want), i
    assert not got[2].any()


def test_whole_feed_properties_at_full_batch(cuda):
    """Size-independent properties at BASELINE configs[2]'s batch (128 frames): the identity transform returns the frame, a
    pure integer shift moves it, and the byte and float outputs agree."""
    from hoisdf_b200 import feed
    gen = torch.Generator().manual_seed(5)
    frames = torch.randint(0, 256, (128, 256, 256, 3), dtype=torch.uint8, generator=gen).to(cuda)
    ident = np.tile(np.array([1.0, 0, 0, 0, 1.0, 0]), (128, 1))
    assert torch.equal(feed.crop_images(frames, ident, 256, as_bytes=True), frames)
    shift = np.tile(np.array([1.0, 0, 3.0, 0, 1.0, -2.0]), (128, 1))
    moved = feed.crop_images(frames, shift, 256, as_bytes=True)
    assert torch.equal(moved[:, 2:, :253], frames[:, :254, 3:]) and not moved[:, :2].any() and not moved[:, :, 253:].any()
    as_float = feed.crop_images(frames, shift, 256)
    # (on the host: torch's CUDA division by a scalar multiplies by the reciprocal, which is not the IEEE quotient)
    assert torch.equal(as_float.cpu(), moved.cpu().permute(0, 3, 1, 2).float() / 255.0)


def _close32(got, want):
    return float(np.abs(got - want).max()) <= 4 * np.finfo(np.float32).eps * max(1.0, float(np.abs(want).max()))


@pytest.mark.parametrize("train,use_rot,flip", [(True, True, False), (False, False, True)])
def test_sdf_point_sets_vs_oracle(cuda, train, use_rot, flip):
    """`hoisdf_sdf_rows_fwd` at the training configuration's set sizes (cfg.num_samp_hand / num_samp_obj = 600 / 200, batch 64)."""
    from hoisdf_b200 import feed
    n_hand, n_obj, hs, os_, B = 600, 200, 6.2, 5.8, 64
    frames = [FO.synthetic_sdf_frame(s, n_hand, n_obj, train=train) for s in range(B)]
    rows = torch.from_numpy(np.concatenate([f[0] for f in frames])).to(cuda)
    offsets = torch.from_numpy(np.cumsum([0] + [len(f[0]) for f in frames]).astype(np.int64))
    index = torch.from_numpy(np.stack([f[2] for f in frames]))
    rot = torch.from_numpy(np.stack([f[3] for f in frames])) if use_rot else None
    flips = torch.from_numpy((np.arange(B) % 3 == 0).astype(np.int32)) if flip else None
    root = torch.from_numpy(np.stack([f[4] for f in frames]))
    centre = torch.from_numpy(np.stack([f[5] for f in frames]))
    inputs, targets = feed.sdf_point_sets(rows, offsets, index, n_hand, n_obj, root, centre, hs, os_, rot=rot, flip=flips)
    got = {k: v.cpu().numpy() for k, v in {**inputs, **targets}.items()}
    assert ("hand_pre_points" in got) == train
    for i, (data, _, idx, r, ro, ce) in enumerate(frames):
        want_in, want_t = FO.sdf_point_sets(data, idx, n_hand, n_obj, ro, ce, hs, os_, rot_mat=r if use_rot else None,
                                            do_flip=bool(flip and i % 3 == 0))
        for k, v in {**want_in, **want_t}.items():
            assert _close32(got[k][i], v), (k, i)
            if not use_rot:
                assert np.array_equal(got[k][i], v), (k, i)
    bad = index.clone()
    bad[5, 7] = 10 ** 9
    with pytest.raises(IndexError):
        feed.sdf_point_sets(rows, offsets, bad, n_hand, n_obj, root, centre, hs, os_, rot=rot, flip=flips)


def test_training_item_reproduces_the_upstream_fixture(cuda):
    """The three feed calls together against ONE sample of the unmodified upstream `Dataset.__getitem__` (mode "train";
    fixture made by oracle/make_golden.py:feed_case through oracle/reference_shim.py:ho3d_train_item)."""
    from hoisdf_b200 import feed
    g = np.load(GOLDEN)
    seed = int(g["seed"])
    sdf = FO.synthetic_sdf_frame(seed, 24, 8)[0]
    frame, hand_mask, obj_mask = FO.synthetic_aug(seed)[:3]
    coef = feed.pil_coefficients(feed.crop_affine(g["item_center"], float(g["item_scale"]), 256, float(g["item_rot"])))[None]
    img = feed.crop_images(torch.from_numpy(frame[None]).to(cuda), coef, 256)
    assert np.array_equal(img[0].cpu().numpy(), g["u8_to_f32"][g["item_img_bytes"]].transpose(2, 0, 1))
    masks = feed.crop_masks(torch.from_numpy(np.stack([hand_mask, obj_mask])).to(cuda), np.tile(coef, (2, 1)), 256, 128)
    assert np.array_equal(masks[0].cpu().numpy(), g["item_hand_seg"]) and np.array_equal(masks[1].cpu().numpy(), g["item_obj_seg"])
    inputs, targets = feed.sdf_point_sets(torch.from_numpy(sdf).to(cuda), torch.tensor([0, len(sdf)]),
                                          torch.from_numpy(g["item_draws"])[None], 24, 8,
                                          torch.from_numpy(g["item_mano_root"])[None],
                                          torch.from_numpy(g["item_obj_center_cam"])[None], float(g["item_hand_sdf_scale"]),
                                          float(g["item_obj_sdf_scale"]), rot=torch.from_numpy(g["item_rot_mat"])[None])
    for k, v in {**inputs, **targets}.items():
        assert _close32(v[0].cpu().numpy(), g["item_" + k]), k


def test_mirrored_warp_equals_warping_the_mirrored_frame(cuda):
    """data/dexycb.py:427-430,479-481 (left hands): `mirror` reads the frame flipped left-right, in both of Pillow's paths."""
    from hoisdf_b200 import feed
    draws = [FO.synthetic_aug(20 + s) for s in range(4)]
    coef = np.stack([feed.pil_coefficients(feed.crop_affine(c, sc, 256, r if i else 0.0))
                     for i, (_, _, _, c, sc, r) in enumerate(draws)])
    frames = torch.from_numpy(np.stack([d[0] for d in draws])).to(cuda)
    mirror = np.array([1, 1, 0, 1])
    got = feed.crop_images(frames, coef, 256, as_bytes=True, mirror=mirror)
    want = feed.crop_images(torch.where(torch.from_numpy(mirror).to(cuda).bool()[:, None, None, None], frames.flip(2), frames),
                            coef, 256, as_bytes=True)
    assert torch.equal(got, want)
    for i, d in enumerate(draws):
        src = np.ascontiguousarray(d[0][:, ::-1, :]) if mirror[i] else d[0]
        pil = np.asarray(Image.fromarray(src).transform((256, 256), Image.AFFINE, tuple(float(c) for c in coef[i])))
        assert np.array_equal(got[i].cpu().numpy(), pil), i
    masks = torch.from_numpy(np.stack([d[1] for d in draws])).to(cuda)
    assert torch.equal(feed.crop_masks(masks, coef, 256, 128, mirror=np.ones(4)), feed.crop_masks(masks.flip(2).contiguous(), coef, 256, 128))



def test_fused_mask_crop_equals_the_two_step_route(cuda):
    """`hoisdf_mask_crop_fwd` (warp + NEAREST shrink + float in one launch) against two `hoisdf_image_crop_fwd` calls and against
    Pillow: rotated and un-rotated warps, mirrored masks, windows leaving the frame, shrink ratios 2, 4 and a non-integer one."""
    from hoisdf_b200 import feed
    draws = [FO.synthetic_aug(50 + s) for s in range(6)]
    coef = np.stack([feed.pil_coefficients(feed.crop_affine(c, sc, 256, r if i % 2 else 0.0))
                     for i, (_, _, _, c, sc, r) in enumerate(draws)])
    coef[4] = [3.1, 0, -200.0, 0, 2.4, 150.0]
    masks_np = np.stack([d[1] for d in draws])
    masks = torch.from_numpy(masks_np).to(cuda)
    mirror = np.array([0, 1, 1, 0, 1, 0])
    for out_res in (128, 64, 100):
        got = feed.crop_masks(masks, coef, 256, out_res, mirror=mirror)
        warped = feed._warp(masks.unsqueeze(3), coef, 256, 1.0, True, mirror)
        two_step = feed._warp(warped, np.tile(feed.resize_coefficients(256, out_res), (6, 1)), out_res, 1.0, False).squeeze(1)
        assert got.shape == (6, out_res, out_res) and torch.equal(got, two_step), out_res
        for i in range(6):
            src = np.ascontiguousarray(masks_np[i][:, ::-1]) if mirror[i] else masks_np[i]
            pil = Image.fromarray(src).transform((256, 256), Image.AFFINE, tuple(float(c) for c in coef[i]))
            want = np.asarray(pil.resize((out_res, out_res), Image.NEAREST)).astype(np.float32)
            assert np.array_equal(got[i].cpu().numpy(), want), (out_res, i)


# ---------------------------------------------------------------------------------------------- photometric augmentation
def test_gaussian_blur_vs_pillow(cuda):
    from PIL import ImageFilter
    from hoisdf_b200 import feed
    rng = np.random.default_rng(11)
    for (h, w, ch), radii in (((256, 256, 3), list(rng.uniform(0, 0.5, 14)) + [0.0, 0.5]), ((37, 53, 3), [0.3, 0.9, 2.0, 6.0]),
                              ((64, 48, 1), [0.45, 3.7])):
        imgs = rng.integers(0, 256, (len(radii), h, w, ch), dtype=np.uint8)
        got = feed.gaussian_blur(torch.from_numpy(imgs).to(cuda), radii).cpu().numpy()
        for i, r in enumerate(radii):
            pil = Image.fromarray(imgs[i] if ch == 3 else imgs[i, :, :, 0])
            want = np.asarray(pil.filter(ImageFilter.GaussianBlur(float(r)))).reshape(h, w, ch)
            assert np.array_equal(got[i], want), (h, w, r)


def test_color_jitter_vs_torchvision(cuda):
    import itertools
    import torchvision.transforms.functional as TF
    from hoisdf_b200 import feed
    fn = {"brightness": TF.adjust_brightness, "saturation": TF.adjust_saturation, "hue": TF.adjust_hue,
          "contrast": TF.adjust_contrast}
    rng = np.random.default_rng(12)
    orders = list(itertools.permutations(fn))
    imgs = rng.integers(0, 256, (len(orders), 96, 80, 3), dtype=np.uint8)
    imgs[1] //= 4
    steps = []
    for order in orders:
        f = {"brightness": rng.uniform(0.5, 1.5), "saturation": rng.uniform(0.5, 1.5), "contrast": rng.uniform(0.5, 1.5),
             "hue": rng.uniform(-0.15, 0.15)}
        steps.append([(n, float(f[n])) for n in order])
    steps[3] = steps[3][:1]
    steps[4] = []
    got = feed.color_jitter(torch.from_numpy(imgs).to(cuda), steps).cpu().numpy()
    for i, seq in enumerate(steps):
        pil = Image.fromarray(imgs[i])
        for n, f in seq:
            pil = fn[n](pil, f)
        assert np.array_equal(got[i], np.asarray(pil)), seq
    # the colour cube (a 64^3 lattice) through the hue round trip
    v = np.arange(0, 256, 4, dtype=np.uint8)
    cube = np.stack(np.meshgrid(v, v, v, indexing="ij"), -1).reshape(512, 512, 3)
    got = feed.color_jitter(torch.from_numpy(cube[None]).to(cuda), [[("hue", -0.12)]]).cpu().numpy()
    assert np.array_equal(got[0], np.asarray(TF.adjust_hue(Image.fromarray(cube), -0.12)))


def test_training_image_reproduces_the_upstream_fixture(cuda):
    """warp -> blur -> jitter -> tensor of ONE sample of the unmodified upstream `Dataset.__getitem__` with its filters on."""
    import random
    from hoisdf_b200 import feed
    g = np.load(GOLDEN)
    seed = int(g["seed"])
    state = random.getstate()
    random.seed(seed)
    radius = random.random() * 0.5
    steps = feed.draw_color_jitter(brightness=0.5, contrast=0.5, saturation=0.5, hue=0.15)
    random.setstate(state)
    coef = feed.pil_coefficients(feed.crop_affine(g["filt_center"], float(g["filt_scale"]), 256, float(g["filt_rot"])))[None]
    frame = torch.from_numpy(FO.synthetic_aug(seed)[0][None]).to(cuda)
    img = feed.to_tensor(feed.color_jitter(feed.gaussian_blur(feed.crop_images(frame, coef, 256, as_bytes=True), [radius]),
                                           [steps]))
    assert np.array_equal(img[0].cpu().numpy()[:, ::8], g["filt_img_rows"])


def test_eval_batch_reproduces_the_upstream_item(cuda):
    """`feed.eval_batch` at BASELINE configs[1]'s batch (32 frames; the fixture's frame first and last) against one sample of
    the unmodified upstream `Dataset.__getitem__` in evaluation mode -- the feed of main/test.py's loop."""
    from hoisdf_b200 import feed
    g = np.load(GOLDEN)
    seed = int(g["seed"])
    picks = [seed] + list(range(100, 130)) + [seed]
    raw = [FO.synthetic_eval_annotation(s) for s in picks]
    samples = [feed.eval_geometry(ann, corners, (640, 480), 0.7) for _, ann, corners in raw]
    inputs, targets, meta = feed.eval_batch(torch.from_numpy(np.stack([r[0] for r in raw])).to(cuda), samples)
    assert inputs["img"].shape == (32, 3, 256, 256) and meta["obj_mask"].dtype == torch.bool and len(meta["obj_cls"]) == 32
    for i in (0, 31):
        assert np.array_equal(inputs["img"][i].cpu().numpy()[:, ::8], g["evi_img_rows"])
        for k in ("obj_rot", "rel_obj_trans"):
            assert np.allclose(targets[k][i].cpu().numpy(), g["evi_" + k], rtol=1e-5, atol=1e-5), k
        for k in ("cam_intr", "mano_root", "obj_center_cam", "bbox_hand", "bbox_obj"):
            assert np.allclose(meta[k][i].cpu().numpy(), g["evi_" + k], rtol=1e-5, atol=1e-4), k
        assert bool(meta["obj_mask"][i]) == bool(g["evi_obj_mask"]) and meta["obj_cls"][i] == str(g["evi_obj_cls"])
    for i, (img, ann, corners) in enumerate(raw):                     # every frame against the oracle's data_crop
        K = np.array(ann["camMat"], dtype=np.float32)
        want = FO.data_crop(img, K, np.array(ann["handBoundingBox"], dtype=np.float32), _project(ann, corners, K))[0]
        assert np.array_equal(inputs["img"][i].cpu().numpy(), want), i


def _project(ann, corners, K):
    """data/ho3d_util.py:44-63 (test-side restatement for the oracle call above)."""
    import cv2
    pose = np.zeros((4, 4))
    pose[:3, 3] = ann["objTrans"]
    pose[:3, :3] = cv2.Rodrigues(ann["objRot"].reshape((3,)))[0]
    pose[1, :] = -pose[1, :]
    pose[2, :] = -pose[2, :]
    uv = np.matmul(np.array(K), np.matmul(pose[:3, :3], np.array(corners).T) + pose[:3, 3].reshape(-1, 1)).T
    return uv[:, :2] / uv[:, -1:]


def test_dexycb_eval_batch_reproduces_the_upstream_item(cuda):
    """`feed.dexycb_eval_batch` (BASELINE configs[2]'s feed) on a batch mixing right and left hands, the fixture's LEFT-hand
    sample first and last: image (mirrored warp), masks, point sets (x-flipped), SDF targets and the host geometry of one
    sample of the unmodified upstream `dexycb.Dataset.__getitem__` in test mode."""
    from hoisdf_b200 import feed
    from test_feed import dexycb_product_sample, DEX_TARGETS, EVAL_META
    g = np.load(GOLDEN)
    seed = int(g["seed"])
    made = [dexycb_product_sample(seed, left=True)] + [dexycb_product_sample(s) for s in range(200, 206)] + \
           [dexycb_product_sample(seed, left=True)]
    frames = torch.from_numpy(np.stack([m[1] for m in made])).to(cuda)
    hand_masks = torch.from_numpy(np.stack([m[2] for m in made])).to(cuda)
    obj_masks = torch.from_numpy(np.stack([m[3] for m in made])).to(cuda)
    rows = torch.from_numpy(np.concatenate([m[4] for m in made])).to(cuda)
    offsets = torch.from_numpy(np.cumsum([0] + [len(m[4]) for m in made]).astype(np.int64))
    inputs, targets, meta = feed.dexycb_eval_batch(frames, hand_masks, obj_masks, rows, offsets, [m[0] for m in made], 24, 8,
                                                   6.2, 5.8)
    assert inputs["img"].shape == (8, 3, 256, 256) and inputs["hand_pre_points"] is False and meta["bbox_hand"].dtype == torch.float64
    for i in (0, 7):
        assert np.array_equal(inputs["img"][i].cpu().numpy()[:, ::8], g["dxi_img_rows"])
        assert np.array_equal(targets["hand_seg"][i].cpu().numpy(), g["dxi_hand_seg"])
        assert np.array_equal(targets["obj_seg"][i].cpu().numpy(), g["dxi_obj_seg"])
        for k in ("hand_sdf_points", "obj_sdf_points"):
            assert _close32(inputs[k][i].cpu().numpy(), g["dxi_" + k]), k
        for k in ("hand_sdf", "obj_sdf"):
            assert _close32(targets[k][i].cpu().numpy(), g["dxi_" + k]), k
        for k in DEX_TARGETS:
            assert np.allclose(targets[k][i].cpu().numpy(), g["dxi_" + k], rtol=1e-5, atol=1e-4), k
        for k in EVAL_META:
            assert np.allclose(meta[k][i].cpu().numpy(), g["dxi_" + k], rtol=1e-5, atol=1e-4), k
        assert int(meta["obj_cls"][i]) == int(g["dxi_obj_cls"])
    for i, (s, img, hm, om, sdf) in enumerate(made):                  # every frame against Pillow on the mirrored source
        src = np.ascontiguousarray(img[:, ::-1, :]) if s["flip"] else img
        pil = np.asarray(Image.fromarray(src).transform((256, 256), Image.AFFINE, tuple(float(c) for c in s["coef"])))
        want = np.ascontiguousarray(pil.astype(np.float32).transpose(2, 0, 1)) / np.float32(255.0)
        assert np.array_equal(inputs["img"][i].cpu().numpy(), want), i


# ---------------------------------------------------------------------------------------------- feed -> model (last: they build whole models)

def test_eval_batch_feeds_the_model(cuda):
    """Raw frames -> `feed.eval_batch` -> `Model.forward(..., "eval")`: the collated dicts are consumable as they are (upstream's
    keys and dtypes, incl. the float64 `obj_rot` target and the string entries of meta_info), and the outputs equal those of
    the same forward on an image batch assembled from the ORACLE's `data_crop` -- main/test.py:120-126 end to end."""
    from hoisdf_b200 import feed, synthetic as syn
    from hoisdf_b200.config import cfg
    from hoisdf_b200.model import get_model
    old = (cfg.setting, cfg.dataset, cfg.num_samp_hand, cfg.num_samp_obj)
    try:
        cfg.set_setting("ho3d")
        type(cfg).dataset = "ho3d"
        type(cfg).num_samp_hand, type(cfg).num_samp_obj = 96, 40
        model = get_model("test", mano_buffers=syn.mano_buffers(5))
        model.load_state_dict(syn.full_state_dict(5, "ho3d"), strict=True)
        model = model.to(cuda).eval()
        raw = [FO.synthetic_eval_annotation(s) for s in (100, 101, 102, 103)]
        samples = [feed.eval_geometry(ann, corners, (640, 480), 0.7) for _, ann, corners in raw]
        inputs, targets, meta = feed.eval_batch(torch.from_numpy(np.stack([r[0] for r in raw])).to(cuda), samples)
        out = model(inputs, targets, meta, "eval")
        crops = []
        for img, ann, corners in raw:
            K = np.array(ann["camMat"], dtype=np.float32)
            crops.append(FO.data_crop(img, K, np.array(ann["handBoundingBox"], dtype=np.float32), _project(ann, corners, K))[0])
        want = model({"img": torch.from_numpy(np.stack(crops)).to(cuda)}, targets, meta, "eval")
        for k in ("hand_joints_out", "mano_joints_out", "mano_mesh_out", "obj_rot_out", "obj_trans_out"):
            assert out[k].shape[0] == 4 and torch.isfinite(out[k]).all(), k
            err = float((out[k] - want[k]).abs().max()) / max(float(want[k].abs().max()), 1e-12)
            assert err < 1e-5, (k, err)
    finally:
        cfg.set_setting(old[0])
        type(cfg).dataset = old[1]
        type(cfg).num_samp_hand, type(cfg).num_samp_obj = old[2], old[3]


def test_dexycb_eval_batch_feeds_the_model(cuda):
    """Raw DexYCB material -> `feed.dexycb_eval_batch` -> `Model.forward(..., "eval")` with cfg.dataset = "dexycb" (BASELINE
    configs[2]'s path from the frames on): the collated dicts are consumable as they are -- every key the dataset branch reads
    (upstream model.py:370-422,606-654) -- and the ground-truth pass-through entries come back unchanged."""
    from hoisdf_b200 import feed, synthetic as syn
    from hoisdf_b200.config import cfg
    from hoisdf_b200.model import get_model
    from test_feed import dexycb_product_sample
    old = (cfg.setting, cfg.dataset, cfg.num_samp_hand, cfg.num_samp_obj)
    try:
        cfg.set_setting("dexycb")
        type(cfg).dataset, type(cfg).num_samp_hand, type(cfg).num_samp_obj = "dexycb", 96, 40
        model = get_model("test", mano_buffers=syn.mano_buffers(14))
        model.load_state_dict(syn.full_state_dict(14, "dexycb"), strict=True)
        model = model.to(cuda).eval()
        made = [dexycb_product_sample(s, n_hand=96, n_obj=40) for s in (200, 201, 202, 203)]
        rows = torch.from_numpy(np.concatenate([m[4] for m in made])).to(cuda)
        offsets = torch.from_numpy(np.cumsum([0] + [len(m[4]) for m in made]).astype(np.int64))
        inputs, targets, meta = feed.dexycb_eval_batch(
            torch.from_numpy(np.stack([m[1] for m in made])).to(cuda), torch.from_numpy(np.stack([m[2] for m in made])).to(cuda),
            torch.from_numpy(np.stack([m[3] for m in made])).to(cuda), rows, offsets, [m[0] for m in made], 96, 40,
            cfg.hand_sdf_scale, cfg.obj_sdf_scale)
        out = model(inputs, targets, meta, "eval")
        for k in ("hand_joints_out", "mano_joints_out", "mano_mesh_out", "obj_rot_out", "obj_trans_out", "joint_heatmap_out",
                  "hand_seg_pred_out", "obj_seg_pred_out", "mano_joints_gt_out", "mano_mesh_gt_out"):
            assert out[k].shape[0] == 4 and torch.isfinite(out[k]).all(), k
        assert torch.equal(out["hand_seg_gt_out"], targets["hand_seg"]) and torch.equal(out["obj_seg_gt_out"], targets["obj_seg"])
        for k in ("sdfhand_loss", "sdfobj_loss", "joint_heatmap", "obj_seg", "hand_seg"):
            assert torch.isfinite(out[k]).all(), k
    finally:
        cfg.set_setting(old[0])
        type(cfg).dataset = old[1]
        type(cfg).num_samp_hand, type(cfg).num_samp_obj = old[2], old[3]


# ---------------------------------------------------------------------------------------------- the fused training-image kernel and its users (last)


def test_fused_training_image_equals_the_step_by_step_calls(cuda):
    """`hoisdf_train_image_fwd` (one CTA per frame, the image resident in shared memory) against warp -> blur -> jitter ->
    tensor through the separate entry points: rotated and scale-only warps, mirrored frames, windows leaving the frame, radius
    0, fewer than four adjustments, every adjustment first -- bit-identical; a box radius >= 1 takes the fallback."""
    import itertools
    from hoisdf_b200 import feed
    rng = np.random.default_rng(21)
    draws = [FO.synthetic_aug(40 + s) for s in range(8)]
    coef = np.stack([feed.pil_coefficients(feed.crop_affine(c, sc, 256, r if i % 4 else 0.0))
                     for i, (_, _, _, c, sc, r) in enumerate(draws)])
    coef[5] = [3.1, 0, -200.0, 0, 2.4, 150.0]                            # mostly outside the frame
    frames = torch.from_numpy(np.stack([d[0] for d in draws])).to(cuda)
    mirror = np.array([0, 1, 0, 1, 1, 0, 0, 1])
    radii = [0.0, 0.49, 0.2, 0.31, 0.05, 0.44, 0.13, 0.38]
    orders = list(itertools.permutations(["brightness", "saturation", "hue", "contrast"]))
    steps = []
    for i in range(8):
        f = {"brightness": rng.uniform(0.5, 1.5), "saturation": rng.uniform(0.5, 1.5), "contrast": rng.uniform(0.5, 1.5),
             "hue": rng.uniform(-0.15, 0.15)}
        steps.append([(n, float(f[n])) for n in orders[(i * 7) % 24]])
    steps[2], steps[3] = steps[2][:2], []
    fused = feed.train_images(frames, coef, radii, steps, 256, mirror)
    warped = feed.crop_images(frames, coef, 256, as_bytes=True, mirror=mirror)
    separate = feed.to_tensor(feed.color_jitter(feed.gaussian_blur(warped, radii), steps))
    assert fused.shape == (8, 3, 256, 256) and torch.equal(fused, separate)
    radii[1] = 2.5                                                       # box radius 1: the fused kernel is not taken
    again = feed.train_images(frames, coef, radii, steps, 256, mirror)
    assert torch.equal(again[0], fused[0]) and not torch.equal(again[1], fused[1])
    assert torch.equal(again, feed.to_tensor(feed.color_jitter(feed.gaussian_blur(warped, radii), steps)))


def test_train_batch_reproduces_the_upstream_item(cuda):
    """`feed.train_batch` on a batch holding the fixture's sample twice (+ one other frame in between): every entry of the
    unmodified upstream `Dataset.__getitem__` item with its filters on -- image, masks, point sets, SDF targets (GPU) and the
    host geometry -- at the fixture's position, and identical results for the repeated sample."""
    from hoisdf_b200 import feed
    from test_feed import product_sample, TARGET_KEYS, META_KEYS
    g = np.load(GOLDEN)
    seed = int(g["seed"])
    state = np.random.get_state()
    picks = [seed, seed + 1, seed]
    host = [product_sample(s) for s in picks]
    np.random.set_state(state)
    aug = [FO.synthetic_aug(s) for s in picks]
    frames = torch.from_numpy(np.stack([a[0] for a in aug])).to(cuda)
    hand_masks = torch.from_numpy(np.stack([a[1] for a in aug])).to(cuda)
    obj_masks = torch.from_numpy(np.stack([a[2] for a in aug])).to(cuda)
    rows = torch.from_numpy(np.concatenate([h[1] for h in host])).to(cuda)
    offsets = torch.from_numpy(np.cumsum([0] + [len(h[1]) for h in host]).astype(np.int64))
    inputs, targets, meta = feed.train_batch(frames, hand_masks, obj_masks, rows, offsets, [h[0] for h in host], 24, 8, 6.2, 5.8)
    assert inputs["img"].shape == (3, 3, 256, 256) and targets["hand_seg"].shape == (3, 128, 128)
    for i in (0, 2):
        assert np.array_equal(inputs["img"][i].cpu().numpy()[:, ::8], g["filt_img_rows"])
        assert np.array_equal(targets["hand_seg"][i].cpu().numpy(), g["filt_t_hand_seg"])
        assert np.array_equal(targets["obj_seg"][i].cpu().numpy(), g["filt_t_obj_seg"])
        for k in ("hand_sdf_points", "obj_sdf_points", "hand_pre_points", "obj_pre_points"):
            assert _close32(inputs[k][i].cpu().numpy(), g["filt_i_" + k]), k
        assert _close32(targets["hand_sdf"][i].cpu().numpy(), g["filt_t_hand_sdf"])
        assert _close32(targets["obj_sdf"][i].cpu().numpy(), g["filt_t_obj_sdf"])
        for k in TARGET_KEYS:
            assert np.allclose(targets[k][i].cpu().numpy(), g["filt_t_" + k], rtol=1e-5, atol=1e-5), k
        for k in META_KEYS:
            assert np.allclose(meta[k][i].cpu().numpy(), g["filt_m_" + k], rtol=1e-5, atol=1e-4), k
    for d in (inputs, targets, meta):
        for k, v in d.items():
            assert torch.equal(v[0], v[2]), k


def test_train_batch_reproduces_the_upstream_dexycb_item(cuda):
    """`feed.train_batch` on DexYCB material (the fixture's LEFT-hand sample + a right-hand one): mirrored warp -> blur ->
    jitter -> tensor, mirrored masks, x-flipped + rotated point sets incl. the `*_pre` sets, and the host geometry of one
    training sample of the unmodified upstream `dexycb.Dataset.__getitem__` with its filters on."""
    from hoisdf_b200 import feed
    from test_feed import dexycb_train_product_sample, DEX_TARGETS, EVAL_META
    g = np.load(GOLDEN)
    made = [dexycb_train_product_sample(int(g["seed"]), left=True), dexycb_train_product_sample(300, left=False)]
    rows = torch.from_numpy(np.concatenate([m[4] for m in made])).to(cuda)
    offsets = torch.from_numpy(np.cumsum([0] + [len(m[4]) for m in made]).astype(np.int64))
    inputs, targets, meta = feed.train_batch(
        torch.from_numpy(np.stack([m[1] for m in made])).to(cuda), torch.from_numpy(np.stack([m[2] for m in made])).to(cuda),
        torch.from_numpy(np.stack([m[3] for m in made])).to(cuda), rows, offsets, [m[0] for m in made], 24, 8, 6.2, 5.8)
    assert np.array_equal(inputs["img"][0].cpu().numpy()[:, ::8], g["dxt_img_rows"])
    assert np.array_equal(np.packbits(targets["hand_seg"][0].cpu().numpy().astype(np.uint8)), g["dxt_hand_seg"])
    assert np.array_equal(np.packbits(targets["obj_seg"][0].cpu().numpy().astype(np.uint8)), g["dxt_obj_seg"])
    for k in ("hand_sdf_points", "obj_sdf_points", "hand_pre_points", "obj_pre_points"):
        assert _close32(inputs[k][0].cpu().numpy(), g["dxt_" + k]), k
    for k in ("hand_sdf", "obj_sdf"):
        assert _close32(targets[k][0].cpu().numpy(), g["dxt_" + k]), k
    for k in DEX_TARGETS:
        assert np.allclose(targets[k][0].cpu().numpy(), g["dxt_" + k], rtol=1e-5, atol=1e-4), k
    for k in EVAL_META:
        assert np.allclose(meta[k][0].cpu().numpy(), g["dxt_" + k], rtol=1e-5, atol=1e-4), k
    assert meta["obj_cls"].shape == (2,)


def test_train_batch_feeds_the_training_step(cuda):
    """Raw DexYCB material -> `feed.train_batch` -> `Model.forward(..., "train")` -> weighted loss sum -> backward
    (main/train.py:104-131 from the frames on, BASELINE configs[3]'s path): the collated dicts are consumable as they are,
    every loss entry is finite and the parameters receive gradients."""
    from hoisdf_b200 import feed, synthetic as syn
    from hoisdf_b200.config import cfg
    from hoisdf_b200.model import get_model
    from hoisdf_b200.train import total_loss
    from test_feed import dexycb_train_product_sample
    old = (cfg.setting, cfg.dataset, cfg.num_samp_hand, cfg.num_samp_obj)
    try:
        cfg.set_setting("dexycb")
        type(cfg).dataset = "ho3d"
        type(cfg).num_samp_hand, type(cfg).num_samp_obj = 48, 16
        model = get_model("train", mano_buffers=syn.mano_buffers(31))
        model.load_state_dict(syn.full_state_dict(31, "dexycb"), strict=True)
        model = model.to(cuda).train()
        made = [dexycb_train_product_sample(s, n_hand=48, n_obj=16) for s in (300, 301)]
        rows = torch.from_numpy(np.concatenate([m[4] for m in made])).to(cuda)
        offsets = torch.from_numpy(np.cumsum([0] + [len(m[4]) for m in made]).astype(np.int64))
        inputs, targets, meta = feed.train_batch(
            torch.from_numpy(np.stack([m[1] for m in made])).to(cuda), torch.from_numpy(np.stack([m[2] for m in made])).to(cuda),
            torch.from_numpy(np.stack([m[3] for m in made])).to(cuda), rows, offsets, [m[0] for m in made], 48, 16,
            cfg.hand_sdf_scale, cfg.obj_sdf_scale)
        out = model(inputs, targets, meta, "train", 0, 0.0)
        total, parts = total_loss(out)
        assert torch.isfinite(total) and all(np.isfinite(float(v)) for v in parts.values()), parts
        total.backward()
        grads = [p.grad for p in model.parameters() if p.grad is not None]
        assert len(grads) > 100 and all(torch.isfinite(g).all() for g in grads)
    finally:
        cfg.set_setting(old[0])
        type(cfg).dataset = old[1]
        type(cfg).num_samp_hand, type(cfg).num_samp_obj = old[2], old[3]
