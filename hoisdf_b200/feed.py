"""Data feed, image path: the crop / warp that upstream's dataset runs per frame on a DataLoader
worker -- `data_crop` (data/ho3d.py:399-427, evaluation) and the image / mask part of `data_aug` (:351-381, training);
data/dexycb.py likewise -- for a whole batch of frames that are already in device memory.

Upstream warps every frame with PIL (`dataset_util.transform_img`, data/dataset_util.py:44-51: an affine `Image.transform` with
PIL's default NEAREST resampling), shrinks the two segmentation masks with `Image.resize((128, 128), NEAREST)` (cfg.output_hm_shape) and ships float
tensors; here the raw 8-bit frames are uploaded once and `hoisdf_image_crop_fwd` produces the
`(B, 3, res, res)` network input and the `(B, 128, 128)` masks, bit-exact with Pillow.  The geometry that comes with the crop --
bounding boxes, the fused crop window, the affine matrix, the updated camera intrinsics -- is a few dozen floating-point
operations per frame and stays on the host in numpy, with the arithmetic (float64 intermediates, `int()` truncations, float32
casts) of data/dataset_util.py so that `cam_intr`, `bbox_hand`, `bbox_obj` and the crop coefficients are the numbers upstream's
dataset returns.

The training-only filters between the warp and the tensor conversion (ho3d.py:355-364: PIL GaussianBlur, then
`dataset_util.color_jitter` = torchvision's four PIL adjustments in a shuffled order) run on the warped bytes on the GPU as
well (`gaussian_blur`, `color_jitter`; csrc/augment.cu), bit-exact with Pillow / torchvision.

Not covered: decoding the image files, the random draws of the geometric augmentation (the caller passes centre, scale and
angle; `draw_sdf_indices` / `draw_color_jitter` reproduce upstream's draws of the point indices and the jitter from the same
generator state), the MANO / object pose rotation (cv2.Rodrigues).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from ._capi import check, lib

__all__ = ["bbox_from_points", "fuse_boxes", "crop_affine", "apply_affine", "pil_coefficients", "resize_coefficients",
           "crop_geometry", "crop_geometry_dexycb", "crop_images", "crop_masks", "data_crop", "draw_sdf_indices", "sdf_point_sets",
           "gaussian_blur", "draw_color_jitter", "color_jitter", "to_tensor", "train_images",
           "draw_train_geometry", "train_geometry", "train_batch",
           "eval_geometry", "eval_batch", "dexycb_annotation", "dexycb_eval_geometry", "dexycb_eval_batch",
           "dexycb_train_geometry"]



def _require_gpu(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise RuntimeError("hoisdf_b200.feed runs on the GPU; got a %s tensor (no CPU fallback)" % t.device)


def _on(dev):
    return torch.cuda.device(dev)


def _stream() -> int:
    return ops._stream()

def bbox_from_points(points2d: np.ndarray, factor: float = 1.1) -> np.ndarray:
    """`dataset_util.get_bbox_joints` (data/dataset_util.py:106-116): box around 2-D points, centre truncated to an integer,
    half extents scaled by `factor`; float32 [x0, y0, x1, y1]."""
    pts = np.asarray(points2d)
    lo, hi = pts.min(0), pts.max(0)
    centre = np.asarray([int((hi[0] + lo[0]) / 2), int((hi[1] + lo[1]) / 2)])
    half = np.asarray([(hi[0] - lo[0]) * factor / 2, (hi[1] - lo[1]) * factor / 2])
    return np.array([*(centre - half), *(centre + half)], dtype=np.float32)


def fuse_boxes(box_a: np.ndarray, box_b: np.ndarray, img_size: Sequence[int], scale_factor: float = 1.0):
    """`dataset_util.fuse_bbox` (:319-332): the square window (integer centre, side) covering both boxes, clipped to the image
    (`img_size` = PIL's (width, height))."""
    both = np.concatenate((np.asarray(box_a).reshape(2, 2), np.asarray(box_b).reshape(2, 2)), axis=0)
    lo, hi = both.min(0), both.max(0)
    x0, y0 = max(0, lo[0]), max(0, lo[1])
    x1, y1 = min(hi[0], img_size[0]), min(hi[1], img_size[1])
    centre = np.asarray([int((x1 + x0) / 2), int((y1 + y0) / 2)])
    return centre, max(x1 - x0, y1 - y0) * scale_factor


def crop_affine(centre: np.ndarray, scale: float, res: int, rot: float = 0.0) -> np.ndarray:
    """`dataset_util.get_affine_transform(center, scale, [res, res], rot)[0]` (:54-66,96-103): rotation by `rot` radians about
    the image origin, then scale / translate the rotated centre to the middle of the crop; source pixel -> crop pixel, float32
    (3, 3).  `rot` = 0 is the evaluation crop, a drawn angle the training augmentation (ho3d.py:318-321)."""
    sn, cs = np.sin(rot), np.cos(rot)
    turn = np.zeros((3, 3))
    turn[0, :2] = [cs, -sn]
    turn[1, :2] = [sn, cs]
    turn[2, 2] = 1
    c = turn.dot(np.asarray(centre).tolist() + [1])[:2]
    m = np.zeros((3, 3))
    m[0, 0] = float(res) / scale
    m[1, 1] = float(res) / scale
    m[0, 2] = res * (-float(c[0]) / scale + 0.5)
    m[1, 2] = res * (-float(c[1]) / scale + 0.5)
    m[2, 2] = 1
    return m.dot(turn).astype(np.float32)


def apply_affine(points2d: np.ndarray, affine: np.ndarray) -> np.ndarray:
    """`dataset_util.transform_coords` (:37-41)."""
    pts = np.asarray(points2d)
    hom = np.concatenate([pts, np.ones([pts.shape[0], 1])], 1)
    return affine.dot(hom.transpose()).transpose()[:, :2]


def pil_coefficients(affine: np.ndarray) -> np.ndarray:
    """`dataset_util.transform_img` (:44-51): PIL's AFFINE `data` = the first two rows of the inverse (crop pixel -> source
    pixel), inverted in the matrix's own precision (float32 for upstream's matrices) and handed to C as doubles."""
    inv = np.linalg.inv(affine)
    return np.array([inv[0, 0], inv[0, 1], inv[0, 2], inv[1, 0], inv[1, 1], inv[1, 2]], dtype=np.float64)


def resize_coefficients(src_size: int, dst_size: int) -> np.ndarray:
    """What `Image.resize((dst, dst), Image.NEAREST)` of a (src, src) image hands to the same C routine (Pillow
    src/_imaging.c `_resize`: a = (src / dst, 0, 0, 0, src / dst, 0)) -- the masks' shrink, ho3d.py:369-371,377-379."""
    return np.array([src_size / dst_size, 0.0, 0.0, 0.0, src_size / dst_size, 0.0], dtype=np.float64)


def _fixed_point_ok(a: np.ndarray, size: int) -> bool:
    """Pillow takes its 16.16 fixed-point loop iff all four output corners map to |coordinate| < 32768."""
    for x, y in ((0, 0), (size, size), (0, size), (size, 0)):
        if not (abs(x * a[0] + y * a[1] + a[2]) < 32768.0 and abs(x * a[3] + y * a[4] + a[5]) < 32768.0):
            return False
    return True


def _warp(frames: torch.Tensor, coefficients: np.ndarray, res: int, divisor: float, as_bytes: bool, mirror=None) -> torch.Tensor:
    _require_gpu(frames)
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] not in (1, 3) or frames.stride(3) != 1 or \
            frames.stride(2) != frames.shape[3]:
        raise ValueError("frames must be (B, H, W, C) uint8 with packed pixels, C = 1 or 3")
    b, h, w, ch = frames.shape
    frame_pitch = frames.stride(0) if b > 1 else h * frames.stride(1)       # (a size-1 batch axis may carry any stride)
    coef = np.ascontiguousarray(np.asarray(coefficients, dtype=np.float64).reshape(b, 6))
    if not np.isfinite(coef).all():
        raise ValueError("non-finite crop coefficients")
    for a in coef:
        if (a[1] != 0.0 or a[3] != 0.0) and not _fixed_point_ok(a, res):
            raise ValueError("crop transform outside Pillow's fixed-point range (|source coordinate| >= 32768)")
    dev = frames.device
    with _on(dev):
        coef_d = torch.from_numpy(coef).to(dev)
        tables = torch.empty(b, 2, res, device=dev, dtype=torch.int32)
        mirror_d = None
        if mirror is not None:
            mirror_d = torch.as_tensor(np.asarray(mirror).astype(np.int32).reshape(b)).to(dev)
        if as_bytes:
            out = torch.empty(b, res, res, ch, device=dev, dtype=torch.uint8)
            args = (None, out.data_ptr())
        else:
            out = torch.empty(b, ch, res, res, device=dev, dtype=torch.float32)
            args = (out.data_ptr(), None)
        ops._count(2)
        check(lib.hoisdf_image_crop_fwd(frames.data_ptr(), b, h, w, ch, frames.stride(1), frame_pitch, coef_d.data_ptr(),
                                        ops._ptr(mirror_d), res, float(divisor), args[0], args[1], tables.data_ptr(), _stream()),
              "hoisdf_image_crop_fwd")
    return out


def crop_images(frames: torch.Tensor, coefficients: np.ndarray, res: int, as_bytes: bool = False, mirror=None) -> torch.Tensor:
    """frames: (B, H, W, 3) uint8 CUDA tensor (row-contiguous); coefficients: (B, 6) PIL AFFINE data (`pil_coefficients`).
    -> (B, 3, res, res) float32 in [0, 1] = `ToTensor()(np.asarray(img.transform(...)).astype(np.float32)) / 255.0`
    (ho3d.py:550,624), or with `as_bytes` the (B, res, res, 3) uint8 PIL images themselves (what the training feed hands to
    its blur / colour-jitter filters).  `mirror` (B,) bool: warp the left-right mirrored frame (dexycb.py:427-430, left
    hands) -- identical to warping `frames.flip(2)`."""
    if frames.dim() != 4 or frames.shape[3] != 3:
        raise ValueError("frames must be (B, H, W, 3) uint8")
    return _warp(frames, coefficients, res, 255.0, as_bytes, mirror)


def crop_masks(masks: torch.Tensor, coefficients: np.ndarray, res: int, out_res: int, mirror=None) -> torch.Tensor:
    """The segmentation masks of the training feed (ho3d.py:366-381, :551-552): masks (B, H, W) uint8 (mode "L") on the GPU ->
    `transform_img` to (res, res) with the frame's coefficients, `.resize((out_res, out_res), Image.NEAREST)`,
    `.astype(np.float32)` -> (B, out_res, out_res) float32.  ONE launch (`hoisdf_mask_crop_fwd`: one CTA per mask, Pillow's
    tables of both steps in shared memory); the same values as the two-step route through `hoisdf_image_crop_fwd`."""
    if masks.dim() != 3:
        raise ValueError("masks must be (B, H, W) uint8")
    _require_gpu(masks)
    if masks.dtype != torch.uint8 or masks.stride(2) != 1:
        raise ValueError("masks must be (B, H, W) uint8 with packed rows")
    b, h, w = masks.shape
    coef = np.ascontiguousarray(np.asarray(coefficients, dtype=np.float64).reshape(b, 6))
    if not np.isfinite(coef).all():
        raise ValueError("non-finite crop coefficients")
    for a in coef:
        if (a[1] != 0.0 or a[3] != 0.0) and not _fixed_point_ok(a, res):
            raise ValueError("crop transform outside Pillow's fixed-point range (|source coordinate| >= 32768)")
    frame_pitch = masks.stride(0) if b > 1 else h * masks.stride(1)
    dev = masks.device
    with _on(dev):
        coef_d = torch.from_numpy(coef).to(dev)
        mirror_d = None if mirror is None else torch.as_tensor(np.asarray(mirror).astype(np.int32).reshape(b)).to(dev)
        out = torch.empty(b, out_res, out_res, device=dev, dtype=torch.float32)
        ops._count(1)
        check(lib.hoisdf_mask_crop_fwd(masks.data_ptr(), b, h, w, masks.stride(1), frame_pitch, coef_d.data_ptr(),
                                       ops._ptr(mirror_d), res, out_res, out.data_ptr(), _stream()), "hoisdf_mask_crop_fwd")
    return out


def crop_geometry(cam_intr: np.ndarray, bbox_hand: np.ndarray, obj_p2d: np.ndarray, img_size: Sequence[int], res: int = 256
                  ) -> Tuple[np.ndarray, Dict[str, np.ndarray]]:
    """The host half of `data_crop` (data/ho3d.py:399-427) for a batch: -> ((B, 6) PIL coefficients, {"cam_intr" (B, 3, 3),
    "bbox_hand" (B, 4), "bbox_obj" (B, 4)} float32).  `img_size` = PIL's (width, height)."""
    b = len(cam_intr)
    coef = np.empty((b, 6), np.float64)
    K_out = np.empty((b, 3, 3), np.float32)
    hand_out = np.empty((b, 4), np.float32)
    obj_out = np.empty((b, 4), np.float32)
    for i in range(b):
        hand = np.asarray(bbox_hand[i]).copy().reshape(2, 2)
        p2d = np.asarray(obj_p2d[i])
        window_hand = bbox_from_points(hand, 1.5)
        window_obj = bbox_from_points(p2d, 1.5)
        box_hand = bbox_from_points(hand, 1.2)
        box_obj = bbox_from_points(p2d, 1.0)
        centre, scale = fuse_boxes(window_hand, window_obj, img_size)
        affine = crop_affine(centre, scale, res)
        hand_out[i] = apply_affine(box_hand.reshape(2, 2), affine).flatten()
        obj_out[i] = apply_affine(box_obj.reshape(2, 2), affine).flatten()
        K_out[i] = affine.dot(np.asarray(cam_intr[i]))
        coef[i] = pil_coefficients(affine)
    return coef, {"cam_intr": K_out, "bbox_hand": hand_out, "bbox_obj": obj_out}


def crop_geometry_dexycb(cam_intr: np.ndarray, joints_uv: np.ndarray, obj_p2d: np.ndarray, img_size: Sequence[int],
                         res: int = 256, heatmap_res: int = 128) -> Tuple[np.ndarray, Dict[str, np.ndarray]]:
    """The host half of DexYCB's `data_crop` (data/dexycb.py:355-404) for a batch: the window comes from the 2-D hand joints
    (factor 1.5; hand box factor 1.1), the intrinsics go through `get_affine_transform(..., K=K)`'s second matrix
    (dataset_util.py:67-91: the crop of the centre carried around the principal point -- without rotation the same centre up
    to float64 rounding, which is kept), the joints are scaled to heat-map pixels and the object corners normalised to the
    object box.  -> ((B, 6) PIL coefficients, {"cam_intr" (B, 3, 3) float64 as upstream, "bbox_hand", "bbox_obj" (B, 4),
    "joints_uv" (B, J, 2), "p2d" (B, N, 2)})."""
    b = len(cam_intr)
    coef = np.empty((b, 6), np.float64)
    out = {"cam_intr": [], "bbox_hand": [], "bbox_obj": [], "joints_uv": [], "p2d": []}
    for i in range(b):
        K = np.asarray(cam_intr[i])
        uv, p2d = np.asarray(joints_uv[i]), np.asarray(obj_p2d[i])
        centre, scale = fuse_boxes(bbox_from_points(uv, 1.5), bbox_from_points(p2d, 1.5), img_size)
        affine = crop_affine(centre, scale, res)
        # the principal-point form of the same crop: T^-1 R T centre with R = identity, evaluated as upstream does
        shift = np.eye(3)
        shift[0, 2], shift[1, 2] = -K[0, 2], -K[1, 2]
        back = shift.copy()
        back[:2, 2] *= -1
        carried = back.dot(np.eye(3)).dot(shift).dot(np.asarray(centre).tolist() + [1])
        post = crop_affine(carried[:2], scale, res)
        box_hand = apply_affine(bbox_from_points(uv, 1.1).reshape(2, 2), affine).flatten()
        box_obj = apply_affine(bbox_from_points(p2d, 1.0).reshape(2, 2), affine).flatten()
        corners = apply_affine(p2d, affine)
        span = box_obj.reshape(2, 2)
        out["cam_intr"].append(post.dot(K))
        out["bbox_hand"].append(box_hand)
        out["bbox_obj"].append(box_obj)
        out["joints_uv"].append(apply_affine(uv, affine) / res * heatmap_res)
        out["p2d"].append((corners - span[0, :]) / (span[1, :] - span[0, :]))
        coef[i] = pil_coefficients(affine)
    return coef, {k: np.stack(v) for k, v in out.items()}


def data_crop(frames: torch.Tensor, cam_intr: np.ndarray, bbox_hand: np.ndarray, obj_p2d: np.ndarray, res: int = 256
              ) -> Tuple[torch.Tensor, Dict[str, np.ndarray]]:
    """Batched `data_crop` (data/ho3d.py:399-427) + the tensor conversion of `__getitem__` (:624).
    frames (B, H, W, 3) uint8 on the GPU; cam_intr (B, 3, 3); bbox_hand (B, 4) = the annotation's hand box; obj_p2d (B, N, 2) =
    projected object box corners.  -> (img (B, 3, res, res) float32, {"cam_intr" (B, 3, 3), "bbox_hand" (B, 4), "bbox_obj" (B, 4)}
    float32) -- the `inputs["img"]` and the geometric `meta_info` entries of upstream's evaluation sample."""
    _, h, w, _ = frames.shape
    coef, meta = crop_geometry(cam_intr, bbox_hand, obj_p2d, (w, h), res)
    return crop_images(frames, coef, res), meta


def draw_sdf_indices(sdf_rows: np.ndarray, n_hand_rows: int, n_hand: int, n_obj: int, dist: Optional[float] = None
                     ) -> np.ndarray:
    """Upstream's draws of one sample (data/ho3d.py:462-482, data/dexycb.py:519-541) from numpy's GLOBAL legacy generator, in
    upstream's order, so that after the same `np.random.seed` the indices ARE upstream's: `np.random.choice(n, ...)` consumes
    the generator exactly as upstream's `np.random.choice(list(range(n)), ...)` does (both reduce to `permutation(n)[:size]`)
    without building a Python list of every row number per sample -- the cost SURVEY section 8 f-4 names.
    sdf_rows: the frame's (N, 6) `.npy` rows on the host, hand rows first; `dist` = cfg.points_filter_dist for the training
    sample's near-surface `*_pre` draws, None for an evaluation sample.  -> `all_idx` int64."""
    n_rows = sdf_rows.shape[0]
    draws = [np.random.choice(n_hand_rows, size=int(n_hand), replace=False),
             n_hand_rows + np.random.choice(n_rows - n_hand_rows, size=int(n_obj), replace=False)]
    if dist is not None:
        draws.append(np.random.choice(np.where(np.abs(sdf_rows[:n_hand_rows, 3]) < dist)[0], size=n_hand, replace=False))
        draws.append(np.random.choice(np.where(np.abs(sdf_rows[n_hand_rows:, 4]) < dist)[0] + n_hand_rows, size=n_obj,
                                      replace=False))
    return np.concatenate(draws).astype(np.int64)


def sdf_point_sets(rows: torch.Tensor, row_offsets: torch.Tensor, index: torch.Tensor, n_hand: int, n_obj: int,
                   hand_root: torch.Tensor, obj_centre: torch.Tensor, hand_scale: float, obj_scale: float,
                   rot: Optional[torch.Tensor] = None, flip: Optional[torch.Tensor] = None
                   ) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """The SDF point sets of a batch of samples (data/ho3d.py:484-486,333,524-548,561-579; data/dexycb.py:515-548) from the
    packed `.npy` rows of the frames, resident on the GPU.
    rows (total, 6) float32 = the frames' `[x, y, z, sdf_hand, sdf_obj, label]` rows back to back, row_offsets (B + 1,) int64;
    index (B, n_sel) int64 = upstream's `all_idx` per frame (the caller's `np.random.choice` draws), n_sel = n_hand + n_obj
    (evaluation) or twice that (training: + the near-surface `*_pre` draws); hand_root / obj_centre (B, 3) float32 (`mano_root`,
    `obj_center_cam`); rot (B, 3, 3) float32 = the augmentation's `rot_mat` or None; flip (B,) = dexycb's `do_flip` or None.
    -> (inputs, targets) with upstream's keys: `hand_sdf_points`, `obj_sdf_points` (, `hand_pre_points`, `obj_pre_points`)
    (B, n, 3); `hand_sdf`, `obj_sdf` (B, n)."""
    _require_gpu(rows)
    dev = rows.device
    if rows.dtype != torch.float32 or rows.dim() != 2 or rows.shape[1] != 6 or not rows.is_contiguous():
        raise ValueError("rows must be a contiguous (total, 6) float32 tensor")
    b, n_sel = index.shape
    per = n_hand + n_obj
    if n_sel not in (per, 2 * per):
        raise ValueError("index must hold n_hand + n_obj draws per frame, or twice that with the *_pre draws")
    if row_offsets.shape != (b + 1,):
        raise ValueError("row_offsets must hold batch + 1 entries")

    def prep(t, dtype, shape, what):
        if t is None:
            return None
        t = t.to(device=dev, dtype=dtype).contiguous()
        if tuple(t.shape) != shape:
            raise ValueError("%s must have shape %s" % (what, shape))
        return t

    row_offsets = prep(row_offsets, torch.int64, (b + 1,), "row_offsets")
    index = prep(index, torch.int64, (b, n_sel), "index")
    hand_root = prep(hand_root, torch.float32, (b, 3), "hand_root")
    obj_centre = prep(obj_centre, torch.float32, (b, 3), "obj_centre")
    rot = prep(rot, torch.float32, (b, 3, 3), "rot")
    flip = prep(flip, torch.int32, (b,), "flip")
    with _on(dev):
        new = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.float32)      # noqa: E731
        hp, op_, hs, os_ = new(b, n_hand, 3), new(b, n_obj, 3), new(b, n_hand), new(b, n_obj)
        pre = n_sel == 2 * per
        hpre, opre = (new(b, n_hand, 3), new(b, n_obj, 3)) if pre else (None, None)
        status = torch.zeros(1, device=dev, dtype=torch.int32)
        ops._count(1)
        check(lib.hoisdf_sdf_rows_fwd(rows.data_ptr(), row_offsets.data_ptr(), index.data_ptr(), b, n_sel, n_hand, n_obj,
                                      ops._ptr(rot), ops._ptr(flip), hand_root.data_ptr(), obj_centre.data_ptr(),
                                      float(hand_scale), float(obj_scale), hp.data_ptr(), op_.data_ptr(), ops._ptr(hpre),
                                      ops._ptr(opre), hs.data_ptr(), os_.data_ptr(), status.data_ptr(), _stream()),
              "hoisdf_sdf_rows_fwd")
        if int(status.item()) != 0:
            raise IndexError("sdf_point_sets: an index lies outside its frame's rows")
    inputs = {"hand_sdf_points": hp, "obj_sdf_points": op_}
    if pre:
        inputs.update(hand_pre_points=hpre, obj_pre_points=opre)
    return inputs, {"hand_sdf": hs, "obj_sdf": os_}


# ---------------------------------------------------------------------------------------------- photometric augmentation
JITTER_OPS = {"brightness": 1, "saturation": 2, "hue": 3, "contrast": 4}


def _bytes_image(images: torch.Tensor, channels=(1, 3)) -> Tuple[int, int, int, int]:
    _require_gpu(images)
    if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[3] not in channels or not images.is_contiguous():
        raise ValueError("images must be a contiguous (B, H, W, C) uint8 tensor, C in %s" % (channels,))
    return tuple(images.shape)


def gaussian_blur(images: torch.Tensor, radii: Sequence[float]) -> torch.Tensor:
    """`img.filter(ImageFilter.GaussianBlur(radius))` (ho3d.py:355-357) per sample: images (B, H, W, C) uint8 on the GPU, radii (B,)
    = upstream's `random.random() * self.blur_radius` draws.  -> the blurred bytes, as Pillow produces them."""
    b, h, w, ch = _bytes_image(images)
    params = _blur_tables(radii, b)
    dev = images.device
    with _on(dev):
        params_d = torch.from_numpy(params.view(np.int32)).to(dev)
        out, scratch = torch.empty_like(images), torch.empty_like(images)
        ops._count(2)
        check(lib.hoisdf_gaussian_blur_u8(images.data_ptr(), out.data_ptr(), scratch.data_ptr(), b, h, w, ch,
                                          params_d.data_ptr(), 3, _stream()), "hoisdf_gaussian_blur_u8")
    return out


def draw_color_jitter(brightness: float = 0, contrast: float = 0, saturation: float = 0, hue: float = 0):
    """The draws of `dataset_util.color_jitter` (data/dataset_util.py:144-201) from Python's GLOBAL `random` generator, in
    upstream's order: four `random.uniform` factors (brightness, contrast, saturation, hue; a range of 0 draws nothing), then
    `random.shuffle` of the adjustments listed as (brightness, saturation, hue, contrast) -- after the same `random.seed` the
    sequence IS upstream's.  -> list of (name, factor) in application order, the `steps` entry of `color_jitter`."""
    import random
    b = random.uniform(max(0, 1 - brightness), 1 + brightness) if brightness > 0 else None
    c = random.uniform(max(0, 1 - contrast), 1 + contrast) if contrast > 0 else None
    sat = random.uniform(max(0, 1 - saturation), 1 + saturation) if saturation > 0 else None
    hu = random.uniform(-hue, hue) if hue > 0 else None
    steps = [(name, f) for name, f in (("brightness", b), ("saturation", sat), ("hue", hu), ("contrast", c)) if f is not None]
    random.shuffle(steps)
    return steps


def _jitter_tables(steps, b: int) -> Tuple[np.ndarray, np.ndarray]:
    """(B, 4) op codes and factors of `hoisdf_color_jitter_u8` / `hoisdf_train_image_fwd` from per-sample (name, factor) lists."""
    if len(steps) != b:
        raise ValueError("one step list per sample")
    codes = np.zeros((b, 4), np.int32)
    factors = np.zeros((b, 4), np.float32)
    for i, seq in enumerate(steps):
        if len(seq) > 4:
            raise ValueError("at most four adjustments per sample")
        for j, (name, f) in enumerate(seq):
            codes[i, j] = JITTER_OPS[name]
            if name == "hue":
                if not -0.5 <= f <= 0.5:
                    raise ValueError("hue_factor (%s) is not in [-0.5, 0.5]." % f)       # torchvision raises the same
                factors[i, j] = float(np.int32(f * 255).astype(np.uint8))                # torchvision's byte shift of H
            else:
                factors[i, j] = f
    return codes, factors


def _blur_tables(radii, b: int) -> np.ndarray:
    """(B, 3) uint32 {n, ww, fw} of `hoisdf_gaussian_blur_params` for the samples' radii."""
    import ctypes
    params = np.zeros((b, 3), np.uint32)
    for i, r in enumerate(np.asarray(radii, dtype=np.float64).reshape(b)):
        check(lib.hoisdf_gaussian_blur_params(float(r), 3, params[i].ctypes.data_as(ctypes.c_void_p)), "hoisdf_gaussian_blur_params")
    return params


def train_images(frames: torch.Tensor, coefficients: np.ndarray, radii: Sequence[float], steps, res: int = 256, mirror=None
                 ) -> torch.Tensor:
    """The network input of a batch of TRAINING frames (ho3d.py:351-364,550): warp -> GaussianBlur -> colour jitter ->
    ToTensor / 255 -- in ONE launch (`hoisdf_train_image_fwd`: one CTA per frame, the warped image resident in shared memory)
    when the crop fits an SM's shared memory and every blur has box radius < 1 (all of upstream's draws), otherwise through
    the step-by-step calls; the bytes are the same either way.  frames (B, H, W, 3) uint8 on the GPU; coefficients (B, 6);
    radii (B,); steps[b] = `draw_color_jitter(...)`; mirror (B,) bool or None.  -> (B, 3, res, res) float32."""
    _require_gpu(frames)
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3 or frames.stride(3) != 1 or frames.stride(2) != 3:
        raise ValueError("frames must be (B, H, W, 3) uint8 with packed RGB pixels")
    b, h, w, _ = frames.shape
    blur = _blur_tables(radii, b)
    codes, factors = _jitter_tables(steps, b)
    smem = lib.hoisdf_train_image_smem_bytes(res)
    if (blur[:, 0] != 0).any() or smem < 0 or smem > 227 * 1024:
        warped = crop_images(frames, coefficients, res, as_bytes=True, mirror=mirror)
        return to_tensor(color_jitter(gaussian_blur(warped, radii), steps))
    coef = np.ascontiguousarray(np.asarray(coefficients, dtype=np.float64).reshape(b, 6))
    if not np.isfinite(coef).all():
        raise ValueError("non-finite crop coefficients")
    for a in coef:
        if (a[1] != 0.0 or a[3] != 0.0) and not _fixed_point_ok(a, res):
            raise ValueError("crop transform outside Pillow's fixed-point range (|source coordinate| >= 32768)")
    frame_pitch = frames.stride(0) if b > 1 else h * frames.stride(1)
    dev = frames.device
    with _on(dev):
        coef_d = torch.from_numpy(coef).to(dev)
        mirror_d = None if mirror is None else torch.as_tensor(np.asarray(mirror).astype(np.int32).reshape(b)).to(dev)
        blur_d = torch.from_numpy(blur.view(np.int32)).to(dev)
        codes_d, factors_d = torch.from_numpy(codes).to(dev), torch.from_numpy(factors).to(dev)
        out = torch.empty(b, 3, res, res, device=dev, dtype=torch.float32)
        ops._count(1)
        check(lib.hoisdf_train_image_fwd(frames.data_ptr(), b, h, w, frames.stride(1), frame_pitch, coef_d.data_ptr(),
                                         ops._ptr(mirror_d), blur_d.data_ptr(), codes_d.data_ptr(), factors_d.data_ptr(), res,
                                         out.data_ptr(), None, _stream()), "hoisdf_train_image_fwd")
    return out


def color_jitter(images: torch.Tensor, steps: Sequence[Sequence[Tuple[str, float]]]) -> torch.Tensor:
    """`dataset_util.color_jitter` for a batch: images (B, H, W, 3) uint8 on the GPU, steps[b] = that sample's adjustments in
    application order (`draw_color_jitter`), each ("brightness" | "saturation" | "hue" | "contrast", factor as torchvision's
    `adjust_*` receives it).  -> the jittered bytes, as torchvision's PIL branch produces them."""
    b, h, w, _ = _bytes_image(images, (3,))
    codes, factors = _jitter_tables(steps, b)
    dev = images.device
    with _on(dev):
        out = torch.empty_like(images)
        sums = torch.empty(4 * b, device=dev, dtype=torch.int64)
        codes_d, factors_d = torch.from_numpy(codes).to(dev), torch.from_numpy(factors).to(dev)     # (kept alive past the launch)
        ops._count(8)
        check(lib.hoisdf_color_jitter_u8(images.data_ptr(), out.data_ptr(), b, h, w, codes_d.data_ptr(), factors_d.data_ptr(),
                                         sums.data_ptr(), _stream()), "hoisdf_color_jitter_u8")
    return out


def to_tensor(images: torch.Tensor) -> torch.Tensor:
    """`ToTensor()(np.asarray(img).astype(np.float32)) / 255.0` (ho3d.py:550) for square (B, S, S, 3) uint8 images on the GPU:
    -> (B, 3, S, S) float32 with IEEE division (torch's CUDA division by a scalar multiplies by the reciprocal instead)."""
    b, h, w, _ = _bytes_image(images, (3,))
    if h != w:
        raise ValueError("to_tensor takes the square crops of the feed")
    return _warp(images, np.tile(np.array([1.0, 0, 0, 0, 1.0, 0]), (b, 1)), h, 255.0, False)


# ---------------------------------------------------------------------------------------------- training sample, host geometry
HO3D_MASKED_OBJECTS = ("021_bleach_cleanser", "006_mustard_bottle", "010_potted_meat_can")        # ho3d.py:554-558,631-635


def _affine_about_principal_point(centre, scale, res, turn, K) -> np.ndarray:
    """The second matrix of `get_affine_transform(..., K=K)` (dataset_util.py:67-91): the un-rotated crop of the centre after
    it was turned about the principal point (T^-1 R T centre) -- what the intrinsics are multiplied with."""
    shift = np.eye(3)
    shift[0, 2], shift[1, 2] = -K[0, 2], -K[1, 2]
    back = shift.copy()
    back[:2, 2] *= -1
    carried = back.dot(turn).dot(shift).dot(np.asarray(centre).tolist() + [1])
    return crop_affine(carried[:2], scale, res)


def draw_train_geometry(centre: np.ndarray, scale: float, center_jittering: float = 0.1, scale_jittering: float = 0.2,
                        max_rot: float = np.pi, dataset: str = "ho3d") -> Tuple[np.ndarray, float, float]:
    """The three geometric draws of `data_aug` (ho3d.py:306-319, dexycb.py:253-274) from the GLOBAL generators in upstream's
    order (centre offsets, scale factor, angle) applied to the fused window -> (centre, scale, rot).  HO3D draws the angle
    uniformly in +-max_rot; DexYCB takes a clipped normal (x 30 degrees, scaled by max_rot / 180) with probability 0.6 --
    that coin is Python's `random.random()`, drawn BEFORE the blur radius -- and no rotation otherwise."""
    import random
    centre = centre + center_jittering * scale * np.random.uniform(low=-1, high=1, size=2)
    factor = np.clip(scale_jittering * np.random.randn() + 1, 1 - scale_jittering, 1 + scale_jittering)
    if dataset == "dexycb":
        rot = np.clip(np.random.randn(), -2.0, 2.0) * 30 if random.random() <= 0.6 else 0
        return centre, scale * factor, rot * max_rot / 180
    return centre, scale * factor, np.random.uniform(low=-max_rot, high=max_rot)


def train_geometry(cam_intr: np.ndarray, joints_uv: np.ndarray, joints_3d: np.ndarray, mano_param: np.ndarray,
                   obj_p2d: np.ndarray, obj_p3d: np.ndarray, obj_rot: np.ndarray, obj_trans: np.ndarray, centre: np.ndarray,
                   scale: float, rot: float, obj_depth_mean_value: Optional[float], res: int = 256, heatmap_res: int = 128,
                   coord_change: Optional[np.ndarray] = None, hand_box_factor: float = 1.2, obj_name: Optional[str] = None
                   ) -> Dict[str, np.ndarray]:
    """Everything of one HO3D training sample that is not pixels or SDF rows (data_aug, ho3d.py:318-349, and the tail of
    `__getitem__`, :519-523,553,568-587), for the drawn (centre, scale, rot): a few dozen floating-point operations in upstream's
    own precision mixture (float64 matrices cast to float32, cv2.Rodrigues for the two rotations), kept on the host.
    -> {"coef" (6,) PIL coefficients of the warp, "rot_mat" (3, 3) for `sdf_point_sets`, and upstream's entries: `joint_coord`,
    `joint_cam_no_trans`, `mano_param`, `obj_rot`, `rel_obj_trans` (targets); `cam_intr`, `mano_root`, `obj_center_cam`,
    `bbox_hand`, `bbox_obj` (meta_info; + `obj_mask` when `obj_name` is given); `p2d`, `p3d` (the normalised corners upstream
    computes and drops)}.
    DexYCB's `data_aug` (dexycb.py:219-354) is the same with `coord_change = np.eye(3)`, `hand_box_factor = 1.1` and the
    object centre at the ROOT joint's depth (`obj_depth_mean_value = None`; dexycb.py:589-592)."""
    import cv2
    if coord_change is None:
        coord_change = np.array([[1.0, 0.0, 0.0], [0, -1.0, 0.0], [0.0, 0.0, -1.0]], dtype=np.float32)     # ho3d.py:70-72
    K = np.asarray(cam_intr).copy()
    mano_param = np.asarray(mano_param).copy()
    sn, cs = np.sin(rot), np.cos(rot)
    turn = np.zeros((3, 3))
    turn[0, :2] = [cs, -sn]
    turn[1, :2] = [sn, cs]
    turn[2, 2] = 1
    affine = crop_affine(centre, scale, res, rot)
    post = _affine_about_principal_point(centre, scale, res, turn, K)
    rot_mat = turn.astype(np.float32)
    # `dataset_util.rotation_angle` (:106-111): global hand rotation from OpenGL to camera axes, then the augmentation's turn
    per_rdg, _ = cv2.Rodrigues(mano_param[:3])
    resrot, _ = cv2.Rodrigues(np.dot(np.dot(rot_mat, coord_change), per_rdg))
    mano_param[:3] = resrot[:, 0].astype(np.float32)
    uv = apply_affine(joints_uv, affine)
    joints_3d = np.asarray(joints_3d).dot(rot_mat.T)
    p3d = np.asarray(obj_p3d).dot(rot_mat.T)
    new_obj_rot = cv2.Rodrigues(rot_mat.dot(cv2.Rodrigues(np.asarray(obj_rot))[0]))[0].squeeze()
    new_obj_trans = rot_mat.dot(np.asarray(obj_trans))
    K = post.dot(K)
    p2d = apply_affine(obj_p2d, affine)
    bbox_hand = bbox_from_points(uv, hand_box_factor)
    uv = uv / res * heatmap_res
    bbox_obj = bbox_from_points(p2d, 1.0)
    span = bbox_obj.reshape(2, 2)
    p2d = (p2d - span[0, :]) / (span[1, :] - span[0, :])
    hand_root = joints_3d[0].copy()
    joints_3d = joints_3d - hand_root[None]
    # `get_center_cam` (:343-350): back-projection of the object box centre at the mean object depth
    depth = hand_root[-1] if obj_depth_mean_value is None else obj_depth_mean_value
    c = np.asarray([int((bbox_obj[2] + bbox_obj[0]) / 2), int((bbox_obj[3] + bbox_obj[1]) / 2), depth])
    centre_cam = np.array([(c[0] - K[0, 2]) / K[0, 0] * c[2], (c[1] - K[1, 2]) / K[1, 1] * c[2], c[2]]).astype(np.float32)
    extra = {} if obj_name is None else {"obj_mask": obj_name in HO3D_MASKED_OBJECTS}                     # ho3d.py:554-558
    return {**extra, "coef": pil_coefficients(affine), "rot_mat": rot_mat, "joint_coord": uv.astype(np.float32),
            "joint_cam_no_trans": joints_3d * 1000, "mano_param": mano_param, "obj_rot": new_obj_rot,
            "rel_obj_trans": new_obj_trans.astype(np.float32) - centre_cam, "cam_intr": K, "mano_root": hand_root,
            "obj_center_cam": centre_cam, "bbox_hand": bbox_hand, "bbox_obj": bbox_obj, "p2d": p2d, "p3d": p3d - centre_cam[None]}


def train_batch(frames: torch.Tensor, hand_masks: torch.Tensor, obj_masks: torch.Tensor, rows: torch.Tensor,
                row_offsets: torch.Tensor, samples: Sequence[Dict], n_hand: int, n_obj: int, hand_sdf_scale: float,
                obj_sdf_scale: float, res: int = 256, heatmap_res: int = 128):
    """One collated training batch (HO3D, or DexYCB when the sample dicts come from `dexycb_train_geometry` and carry "flip")
    -- what upstream's DataLoader hands to `main/train.py:104-108` -- from the raw material
    on the GPU: frames (B, H, W, 3) uint8, hand / object masks (B, H, W) uint8 (the unpacked bits), the frames' packed SDF rows
    (`sdf_point_sets`), and per frame the host results `samples[b]` = `train_geometry(...)`'s dict + "index" (`draw_sdf_indices`),
    "blur_radius" (`random.random() * blur_radius`) and "jitter" (`draw_color_jitter(...)`).
    -> (inputs, targets, meta_info): dicts of CUDA tensors with upstream's keys (ho3d.py:561-587), batch-first."""
    dev = frames.device
    b = frames.shape[0]
    if len(samples) != b:
        raise ValueError("one sample dict per frame")
    coef = np.stack([s["coef"] for s in samples])
    mirror = np.array([bool(s.get("flip", False)) for s in samples])           # DexYCB left hands (dexycb.py:427-430,479-481,547)
    img = train_images(frames, coef, [s["blur_radius"] for s in samples], [s["jitter"] for s in samples], res, mirror)
    stack = lambda key: torch.from_numpy(np.stack([np.asarray(s[key]) for s in samples])).to(dev)  # noqa: E731  (dtypes as upstream's collate)
    inputs, targets = sdf_point_sets(rows, row_offsets, torch.from_numpy(np.stack([s["index"] for s in samples])), n_hand, n_obj,
                                     stack("mano_root"), stack("obj_center_cam"), hand_sdf_scale, obj_sdf_scale,
                                     rot=stack("rot_mat"), flip=torch.from_numpy(mirror.astype(np.int32)))
    inputs["img"] = img
    segs = crop_masks(torch.cat([hand_masks, obj_masks]), np.concatenate([coef, coef]), res, heatmap_res,
                      mirror=np.concatenate([mirror, mirror]))                 # both mask sets in one launch
    targets.update(hand_seg=segs[:len(samples)], obj_seg=segs[len(samples):])
    for key in ("joint_coord", "joint_cam_no_trans", "obj_rot", "rel_obj_trans", "mano_param"):
        targets[key] = stack(key)
    meta = {key: stack(key) for key in ("cam_intr", "mano_root", "obj_center_cam", "bbox_hand", "bbox_obj")}
    if "obj_cls" in samples[0]:
        meta["obj_cls"] = torch.tensor([int(s["obj_cls"]) for s in samples], device=dev)
    if "obj_mask" in samples[0]:                                               # ho3d.py:554-558,583 (caller-provided flag)
        meta["obj_mask"] = torch.tensor([bool(s["obj_mask"]) for s in samples], device=dev)
    return inputs, targets, meta


# ---------------------------------------------------------------------------------------------- evaluation sample (main/test.py's feed)
def eval_geometry(annotation: Dict, obj_bbox3d: np.ndarray, img_size: Sequence[int], obj_depth_mean_value: float, res: int = 256
                  ) -> Dict[str, np.ndarray]:
    """Everything of one HO3D EVALUATION sample that is not pixels (the `else` branch of `__getitem__`, ho3d.py:591-660): from the
    frame's annotation dict (`meta/*.pkl`: camMat, objRot, objTrans, handBoundingBox, handJoints3D, objName) and the object's 3-D
    box corners -- corner projection (ho3d_util.py:44-63), `data_crop`'s geometry, the root joint in camera axes, the object
    centre at the mean depth, the object pose in OpenCV axes (dataset_util.py:27-35).
    -> {"coef" (6,), `obj_rot`, `rel_obj_trans` (targets), `cam_intr`, `mano_root`, `obj_center_cam`, `bbox_hand`, `bbox_obj`,
    `obj_mask`, `obj_cls` (meta_info)}."""
    import cv2
    K = np.array(annotation["camMat"], dtype=np.float32)
    pose = np.zeros((4, 4))
    pose[:3, 3] = annotation["objTrans"]
    pose[3, 3] = 1
    pose[:3, :3] = cv2.Rodrigues(annotation["objRot"].reshape((3,)))[0]
    pose[1, :] = -pose[1, :]                                   # OpenGL -> camera axes
    pose[2, :] = -pose[2, :]
    corners = np.array(obj_bbox3d)
    uv = np.matmul(np.array(K), np.matmul(pose[:3, :3], corners.T) + pose[:3, 3].reshape(-1, 1)).T
    p2d = uv[:, :2] / uv[:, -1:]
    flip = np.array([[1.0, 0.0, 0.0], [0, -1.0, 0.0], [0.0, 0.0, -1.0]], dtype=np.float32)
    root = np.array(annotation["handJoints3D"], dtype=np.float32).dot(flip.T)
    coef, geo = crop_geometry(K[None], np.array(annotation["handBoundingBox"], dtype=np.float32)[None], p2d[None], img_size, res)
    K, bbox_obj = geo["cam_intr"][0], geo["bbox_obj"][0]
    c = np.asarray([int((bbox_obj[2] + bbox_obj[0]) / 2), int((bbox_obj[3] + bbox_obj[1]) / 2), obj_depth_mean_value])
    centre_cam = np.array([(c[0] - K[0, 2]) / K[0, 0] * c[2], (c[1] - K[1, 2]) / K[1, 1] * c[2], c[2]]).astype(np.float32)
    change = np.array([[1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.0, -1.0]])
    obj_pose = annotation["objRot"].squeeze()
    obj_rot = obj_pose.copy()
    obj_rot[:3] = cv2.Rodrigues(change.dot(cv2.Rodrigues(obj_pose[:3])[0]))[0][:, 0]
    obj_trans = annotation["objTrans"].copy().dot(change.T).astype(np.float32) - centre_cam
    name = annotation["objName"]
    return {"coef": coef[0], "obj_rot": obj_rot, "rel_obj_trans": obj_trans.astype(np.float32), "cam_intr": K, "mano_root": root,
            "obj_center_cam": centre_cam, "bbox_hand": geo["bbox_hand"][0], "bbox_obj": bbox_obj, "obj_cls": name,
            "obj_mask": name in HO3D_MASKED_OBJECTS}


def eval_batch(frames: torch.Tensor, samples: Sequence[Dict], res: int = 256):
    """One collated HO3D evaluation batch -- what `main/test.py:120-126` feeds to `model(inputs, targets, meta_info, "eval")` --
    from the decoded frames (B, H, W, 3) uint8 on the GPU and `eval_geometry`'s per-frame dicts.
    -> (inputs {"img"}, targets {"obj_rot", "rel_obj_trans"}, meta_info {"cam_intr", "mano_root", "obj_center_cam", "bbox_hand",
    "bbox_obj", "obj_mask" (tensors), "obj_cls", "hand_type" (lists, as default_collate leaves strings)})."""
    dev = frames.device
    if len(samples) != frames.shape[0]:
        raise ValueError("one sample dict per frame")
    img = crop_images(frames, np.stack([s["coef"] for s in samples]), res)
    stack = lambda key: torch.from_numpy(np.stack([np.asarray(s[key]) for s in samples])).to(dev)  # noqa: E731  (dtypes as upstream's collate)
    meta = {key: stack(key) for key in ("cam_intr", "mano_root", "obj_center_cam", "bbox_hand", "bbox_obj")}
    meta["obj_mask"] = torch.tensor([bool(s["obj_mask"]) for s in samples], device=dev)
    meta["obj_cls"] = [s["obj_cls"] for s in samples]
    meta["hand_type"] = ["right"] * len(samples)
    return {"img": img}, {"obj_rot": stack("obj_rot"), "rel_obj_trans": stack("rel_obj_trans")}, meta


# ---------------------------------------------------------------------------------------------- DexYCB evaluation sample
def dexycb_annotation(sample_info: Dict, components_right: np.ndarray, components_left: np.ndarray, handmean: np.ndarray,
                      obj_bbox3d: np.ndarray, width: int) -> Dict:
    """The annotation algebra at the head of DexYCB's `__getitem__` (data/dexycb.py:409-514), before any crop: intrinsics from
    the annotation, the MANO pose from PCA to axis-angle (right or left components), the left-hand mirror (x-flip of the joints,
    principal point and object pose; re-projection of the object corners).  `sample_info` = the entry of DexYCB's `sample_dict`;
    `obj_bbox3d` = the grasped object's box corners.  -> {"flip", "cam_intr" (float64), "joints_uv", "joints_3d", "mano_param"
    (58,), "obj_p2d", "obj_p3d", "obj_rot", "obj_trans", "obj_cls"}."""
    import cv2
    flip = sample_info["mano_side"] == "left"
    intr = sample_info["intrinsics"]
    K = np.zeros((3, 3))
    K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[2, 2] = intr["fx"], intr["fy"], intr["ppx"], intr["ppy"], 1
    pca = np.array(sample_info["pose_m"], dtype=np.float32).squeeze()
    betas = np.array(sample_info["mano_betas"], dtype=np.float32)
    joints_3d = np.array(sample_info["joint_3d"], dtype=np.float32).squeeze()
    joints_uv = np.array(sample_info["joint_2d"], dtype=np.float32).squeeze()
    comps = components_left if flip else components_right
    pose = np.concatenate((pca[0:3], np.matmul(pca[3:48], comps.copy()), pca[48:]), axis=0)
    if flip:
        turned = pose[:48].reshape(-1, 3)
        turned[:, 1:] *= -1
        pose[0:48] = turned.reshape(-1)
        joints_3d[:, 0] *= -1
        joints_uv[:, 0] = np.array(width, dtype=np.float32) - joints_uv[:, 0] - 1
    mano_param = np.concatenate((np.concatenate((pose[:3], pose[3:48] + handmean.copy()), axis=0), betas))
    grasp = np.array(sample_info["pose_y"][sample_info["ycb_grasp_ind"]], dtype=np.float32)
    corners = np.asarray(obj_bbox3d).copy()

    def project(rt):
        cam = (np.matmul(rt[:3, :3], corners.T) + rt[:3, 3].reshape(-1, 1)).T
        uv = np.matmul(K, np.matmul(rt[:3, :3], corners.T) + rt[:3, 3].reshape(-1, 1)).T
        return cam, uv[:, :2] / uv[:, -1:]

    p3d, p2d = project(grasp)
    obj_rot = cv2.Rodrigues(grasp[:, :3])[0].squeeze()
    obj_trans = grasp[:, 3]
    if flip:
        K[0, 2] = width - K[0, 2] - 1
        obj_trans[0] *= -1
        obj_rot[1:] *= -1
        p3d, p2d = project(np.concatenate([cv2.Rodrigues(obj_rot)[0], obj_trans[:, None]], axis=1))
    return {"flip": flip, "cam_intr": K, "joints_uv": joints_uv, "joints_3d": joints_3d, "mano_param": mano_param, "obj_p2d": p2d,
            "obj_p3d": p3d, "obj_rot": obj_rot, "obj_trans": obj_trans,
            "obj_cls": sample_info["ycb_ids"][sample_info["ycb_grasp_ind"]]}


def dexycb_eval_geometry(sample_info: Dict, components_right: np.ndarray, components_left: np.ndarray, handmean: np.ndarray,
                         obj_bbox3d: np.ndarray, img_size: Sequence[int], res: int = 256, heatmap_res: int = 128
                         ) -> Dict[str, np.ndarray]:
    """Everything of one DexYCB TEST sample that is not pixels or SDF rows (data/dexycb.py:409-514,584-596,627-655):
    `dexycb_annotation`, then `data_crop`'s geometry, root joint, object centre at the ROOT's depth.
    -> {"coef" (6,), "flip" (bool: pass as `mirror` to the warps and `flip` to `sdf_point_sets`), and upstream's entries:
    `joint_coord`, `joint_cam_no_trans`, `obj_rot`, `rel_obj_trans`, `mano_param` (targets), `cam_intr`, `mano_root`,
    `obj_center_cam`, `bbox_hand`, `bbox_obj` (float64 here, as upstream), `obj_cls` (meta_info)}."""
    a = dexycb_annotation(sample_info, components_right, components_left, handmean, obj_bbox3d, img_size[0])
    joints_3d = a["joints_3d"]
    coef, geo = crop_geometry_dexycb(a["cam_intr"][None], a["joints_uv"][None], a["obj_p2d"][None], img_size, res, heatmap_res)
    K, bbox_obj = geo["cam_intr"][0], geo["bbox_obj"][0]
    root = joints_3d[0].copy()
    c = np.asarray([int((bbox_obj[2] + bbox_obj[0]) / 2), int((bbox_obj[3] + bbox_obj[1]) / 2), root[-1]])
    centre_cam = np.array([(c[0] - K[0, 2]) / K[0, 0] * c[2], (c[1] - K[1, 2]) / K[1, 1] * c[2], c[2]]).astype(np.float32)
    return {"coef": coef[0], "flip": a["flip"], "joint_coord": geo["joints_uv"][0].astype(np.float32),
            "joint_cam_no_trans": (joints_3d - root[None]) * 1000, "obj_rot": a["obj_rot"],
            "rel_obj_trans": (a["obj_trans"].astype(np.float32) - centre_cam).astype(np.float32),
            "mano_param": a["mano_param"].astype(np.float32), "cam_intr": K.astype(np.float32), "mano_root": root,
            "obj_center_cam": centre_cam, "bbox_hand": geo["bbox_hand"][0], "bbox_obj": bbox_obj, "obj_cls": a["obj_cls"],
            "p2d": geo["p2d"][0], "p3d": a["obj_p3d"] - centre_cam[None]}


def dexycb_train_geometry(sample_info: Dict, components_right: np.ndarray, components_left: np.ndarray, handmean: np.ndarray,
                          obj_bbox3d: np.ndarray, img_size: Sequence[int], res: int = 256, heatmap_res: int = 128,
                          center_jittering: float = 0.1, scale_jittering: float = 0.2, max_rot: float = np.pi
                          ) -> Dict[str, np.ndarray]:
    """The host half of one DexYCB TRAINING sample (dexycb.py:409-514, `data_aug` :219-354, :584-657): `dexycb_annotation`, the
    fused window, the geometric draws in upstream's order (`draw_train_geometry(dataset="dexycb")` -- call AFTER
    `draw_sdf_indices`, BEFORE the blur / jitter draws) and `train_geometry` with DexYCB's constants.  -> `train_geometry`'s
    dict + "flip", "obj_cls"; `cam_intr` / `mano_param` cast to float32 as upstream's dict does (:645,649)."""
    a = dexycb_annotation(sample_info, components_right, components_left, handmean, obj_bbox3d, img_size[0])
    centre, scale = fuse_boxes(bbox_from_points(a["joints_uv"], 1.5), bbox_from_points(a["obj_p2d"], 1.5), img_size)
    centre, scale, rot = draw_train_geometry(centre, scale, center_jittering, scale_jittering, max_rot, "dexycb")
    g = train_geometry(a["cam_intr"], a["joints_uv"], a["joints_3d"], a["mano_param"], a["obj_p2d"], a["obj_p3d"], a["obj_rot"],
                       a["obj_trans"], centre, scale, rot, None, res, heatmap_res, np.eye(3), 1.1)
    g.update(flip=a["flip"], obj_cls=a["obj_cls"], cam_intr=g["cam_intr"].astype(np.float32),
             mano_param=g["mano_param"].astype(np.float32))
    return g


def dexycb_eval_batch(frames: torch.Tensor, hand_masks: torch.Tensor, obj_masks: torch.Tensor, rows: torch.Tensor,
                      row_offsets: torch.Tensor, samples: Sequence[Dict], n_hand: int, n_obj: int, hand_sdf_scale: float,
                      obj_sdf_scale: float, res: int = 256, heatmap_res: int = 128):
    """One collated DexYCB test batch (BASELINE configs[2]'s feed; data/dexycb.py:627-657) from the raw material on the GPU:
    frames (B, H, W, 3) uint8 and masks (B, H, W) uint8 UN-mirrored, the frames' packed SDF rows, and per frame
    `dexycb_eval_geometry(...)`'s dict + "index" (`draw_sdf_indices(rows, n_hand_rows, n_hand, n_obj)`).
    -> (inputs, targets, meta_info) with upstream's keys (`hand_pre_points` / `obj_pre_points` are False, as upstream)."""
    dev = frames.device
    if len(samples) != frames.shape[0]:
        raise ValueError("one sample dict per frame")
    coef = np.stack([s["coef"] for s in samples])
    mirror = np.array([bool(s["flip"]) for s in samples])
    stack = lambda key: torch.from_numpy(np.stack([np.asarray(s[key]) for s in samples])).to(dev)  # noqa: E731
    inputs, targets = sdf_point_sets(rows, row_offsets, torch.from_numpy(np.stack([s["index"] for s in samples])), n_hand, n_obj,
                                     stack("mano_root"), stack("obj_center_cam"), hand_sdf_scale, obj_sdf_scale,
                                     flip=torch.from_numpy(mirror.astype(np.int32)))
    inputs.update(img=crop_images(frames, coef, res, mirror=mirror), hand_pre_points=False, obj_pre_points=False)
    segs = crop_masks(torch.cat([hand_masks, obj_masks]), np.concatenate([coef, coef]), res, heatmap_res,
                      mirror=np.concatenate([mirror, mirror]))                 # both mask sets in one launch
    targets.update(hand_seg=segs[:len(samples)], obj_seg=segs[len(samples):])
    for key in ("joint_coord", "joint_cam_no_trans", "obj_rot", "rel_obj_trans", "mano_param"):
        targets[key] = stack(key)
    meta = {key: stack(key) for key in ("cam_intr", "mano_root", "obj_center_cam", "bbox_hand", "bbox_obj")}
    meta["obj_cls"] = torch.tensor([int(s["obj_cls"]) for s in samples], device=dev)
    return inputs, targets, meta
