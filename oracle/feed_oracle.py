"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the image path of upstream's evaluation data feed:
`data_crop` of data/ho3d.py:401-427 with the helpers of data/dataset_util.py and the tensor conversion of ho3d.py:624, on
numpy + Pillow (the third-party dependency that defines the resampling: `Image.transform(..., Image.AFFINE, ...)`, default
NEAREST; pinned here against Pillow 12.2.0).  Only tests/ may import this; the product path is hoisdf_b200/feed.py +
csrc/feed.cu.

Pinned to the unmodified upstream functions by tests/test_oracle_vs_reference.py (live, where /root/reference exists) and by the
fixture tests/golden/feed_crop_seed5.npz (generated from the upstream functions by oracle/make_golden.py:feed_case)."""
import numpy as np
from PIL import Image


def get_bbox_joints(joints2d, bbox_factor=1.1):
    """data/dataset_util.py:106-116"""
    mn, mx = joints2d.min(0), joints2d.max(0)
    c = np.asarray([int((mx[0] + mn[0]) / 2), int((mx[1] + mn[1]) / 2)])
    d = np.asarray([(mx[0] - mn[0]) * bbox_factor / 2, (mx[1] - mn[1]) * bbox_factor / 2])
    return np.array([*(c - d), *(c + d)], dtype=np.float32)


def fuse_bbox(bbox_1, bbox_2, img_shape, scale_factor=1.0):
    """data/dataset_util.py:319-332"""
    bbox = np.concatenate((bbox_1.reshape(2, 2), bbox_2.reshape(2, 2)), axis=0)
    min_x, min_y = bbox.min(0)
    min_x, min_y = max(0, min_x), max(0, min_y)
    max_x, max_y = bbox.max(0)
    max_x, max_y = min(max_x, img_shape[0]), min(max_y, img_shape[1])
    center = np.asarray([int((max_x + min_x) / 2), int((max_y + min_y) / 2)])
    return center, max(max_x - min_x, max_y - min_y) * scale_factor


def get_affine_transform(center, scale, res, rot=0.0):
    """data/dataset_util.py:54-66,96-103 (the K = None form): rotation about the origin, then scale / translate into the crop."""
    rot_mat = np.zeros((3, 3))
    sn, cs = np.sin(rot), np.cos(rot)
    rot_mat[0, :2] = [cs, -sn]
    rot_mat[1, :2] = [sn, cs]
    rot_mat[2, 2] = 1
    c = rot_mat.dot(center.tolist() + [1])[:2]
    t = np.zeros((3, 3))
    t[0, 0] = float(res[0]) / scale
    t[1, 1] = float(res[1]) / scale
    t[0, 2] = res[1] * (-float(c[0]) / scale + 0.5)
    t[1, 2] = res[0] * (-float(c[1]) / scale + 0.5)
    t[2, 2] = 1
    return t.dot(rot_mat).astype(np.float32), rot_mat.astype(np.float32)


def transform_coords(pts, affine_trans):
    """data/dataset_util.py:37-41"""
    hom2d = np.concatenate([pts, np.ones([np.array(pts).shape[0], 1])], 1)
    return affine_trans.dot(hom2d.transpose()).transpose()[:, :2]


def transform_img(img, affine_trans, res):
    """data/dataset_util.py:44-51 (img: PIL image)"""
    trans = np.linalg.inv(affine_trans)
    return img.transform(tuple(res), Image.AFFINE,
                         (trans[0, 0], trans[0, 1], trans[0, 2], trans[1, 0], trans[1, 1], trans[1, 2]))


def data_crop(img_u8, K, bbox_hand, p2d, inp_res=256):
    """data/ho3d.py:401-427 + :624.  img_u8 (H, W, 3) uint8 -> (img (3, res, res) float32, K, bbox_hand, bbox_obj)."""
    img = Image.fromarray(img_u8)
    K, bbox_hand = K.copy(), bbox_hand.copy()
    crop_hand = get_bbox_joints(bbox_hand.reshape(2, 2), bbox_factor=1.5)
    crop_obj = get_bbox_joints(p2d, bbox_factor=1.5)
    bbox_hand = get_bbox_joints(bbox_hand.reshape(2, 2), bbox_factor=1.2)
    bbox_obj = get_bbox_joints(p2d, bbox_factor=1.0)
    center, scale = fuse_bbox(crop_hand, crop_obj, img.size)
    affinetrans, _ = get_affine_transform(center, scale, [inp_res, inp_res])
    bbox_hand = transform_coords(bbox_hand.reshape(2, 2), affinetrans).flatten()
    bbox_obj = transform_coords(bbox_obj.reshape(2, 2), affinetrans).flatten()
    img = transform_img(img, affinetrans, [inp_res, inp_res]).crop((0, 0, inp_res, inp_res))
    K = affinetrans.dot(K)
    tensor = np.ascontiguousarray(np.asarray(img).astype(np.float32).transpose(2, 0, 1)) / np.float32(255.0)
    return tensor, K, bbox_hand.astype(np.float32), bbox_obj.astype(np.float32)


def synthetic_frame(seed, h=480, w=640):
    """A deterministic frame + annotation in the HO3D layout (numpy PCG64: identical on every machine)."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    K = np.array([[614.6, 0.0, w / 2 + rng.uniform(-20, 20)], [0.0, 614.2, h / 2 + rng.uniform(-20, 20)], [0, 0, 1]], np.float32)
    cx, cy = rng.uniform(0.3 * w, 0.7 * w), rng.uniform(0.3 * h, 0.7 * h)
    bw, bh = rng.uniform(60, 160), rng.uniform(60, 160)
    bbox_hand = np.array([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], np.float32)
    p2d = (np.array([cx, cy]) + rng.uniform(-140, 140, (8, 2))).astype(np.float32)
    return img, K, bbox_hand, p2d


def aug_warp(img_u8, hand_seg_u8, obj_seg_u8, center, scale, rot, inp_res=256, heatmap_res=128):
    """The image / mask part of the training augmentation, data/ho3d.py:318-321 (affine with the drawn rotation), :351-353
    (frame warp + crop), :366-381 (mask warp + crop + NEAREST shrink), :550-552 (tensor conversion) -- without the random
    blur / colour jitter between them (ho3d.py:355-364).  -> (pil_bytes (res, res, 3) uint8 = the image handed to the blur,
    img (3, res, res) float32, hand_seg (hm, hm) float32, obj_seg (hm, hm) float32, affinetrans (3, 3) float32)."""
    affinetrans, _ = get_affine_transform(np.asarray(center), scale, [inp_res, inp_res], rot=rot)
    img = transform_img(Image.fromarray(img_u8), affinetrans, [inp_res, inp_res]).crop((0, 0, inp_res, inp_res))
    segs = []
    for seg in (hand_seg_u8, obj_seg_u8):
        s = transform_img(Image.fromarray(seg), affinetrans, [inp_res, inp_res]).crop((0, 0, inp_res, inp_res))
        segs.append(np.asarray(s.resize((heatmap_res, heatmap_res), Image.NEAREST)).astype(np.float32))
    pil_bytes = np.asarray(img)
    tensor = np.ascontiguousarray(pil_bytes.astype(np.float32).transpose(2, 0, 1)) / np.float32(255.0)
    return pil_bytes, tensor, segs[0], segs[1], affinetrans


def synthetic_aug(seed, h=480, w=640):
    """Frame, two blocky 0/1 masks (np.unpackbits output in upstream, ho3d.py:446-451) and an augmentation draw
    (centre jitter, scale jitter, rotation up to +-pi as cfg allows) -- deterministic."""
    rng = np.random.default_rng(1000 + seed)
    img, K, bbox_hand, p2d = synthetic_frame(seed, h, w)
    coarse = rng.integers(0, 2, (2, h // 8, w // 8), dtype=np.uint8)
    masks = np.repeat(np.repeat(coarse, 8, axis=1), 8, axis=2)
    crop_hand = get_bbox_joints(bbox_hand.reshape(2, 2), 1.5)
    crop_obj = get_bbox_joints(p2d, 1.5)
    center, scale = fuse_bbox(crop_hand, crop_obj, (w, h))
    center = center + 0.1 * scale * rng.uniform(-1, 1, 2)
    scale = scale * float(np.clip(0.2 * rng.standard_normal() + 1, 0.8, 1.2))
    rot = float(rng.uniform(-np.pi, np.pi))
    return img, masks[0], masks[1], center, scale, rot


def sdf_point_sets(sdf_data, all_idx, n_hand, n_obj, hand_root, obj_center_cam, hand_sdf_scale, obj_sdf_scale, rot_mat=None,
                   do_flip=False):
    """The SDF part of `__getitem__`: data/ho3d.py:484-486 (row gather), :333 (rotation inside data_aug), :524-548
    (normalisation), :561-579 (dict entries); data/dexycb.py:544-548 (mirror flip before the augmentation).
    sdf_data (N, 6) float32 = the frame's .npy; all_idx = concatenated draws.  -> (inputs, targets) dicts of numpy arrays."""
    sdf_data = sdf_data[all_idx]
    sdf_points = sdf_data[:, :5]
    if do_flip:
        sdf_points[:, 0] *= -1
    if rot_mat is not None:
        sdf_points = sdf_points.copy()
        sdf_points[:, :3] = sdf_points[:, :3].dot(rot_mat.T)
    hand_sdf_points = sdf_points[:int(n_hand)]
    obj_sdf_points = sdf_points[int(n_hand):int(n_hand + n_obj)]
    hand_sdf_points[:, :3] = hand_sdf_points[:, :3] - hand_root[None]
    hand_sdf_points = hand_sdf_points * hand_sdf_scale
    obj_sdf_points[:, :3] = obj_sdf_points[:, :3] - obj_center_cam[None]
    obj_sdf_points = obj_sdf_points * obj_sdf_scale
    inputs = {"hand_sdf_points": hand_sdf_points[:, :3], "obj_sdf_points": obj_sdf_points[:, :3]}
    if len(all_idx) == 2 * (n_hand + n_obj):
        hand_pre_points = sdf_points[int(n_hand + n_obj):int(n_hand * 2 + n_obj)]
        obj_pre_points = sdf_points[int(n_hand * 2 + n_obj):]
        hand_pre_points[:, :3] = hand_pre_points[:, :3] - hand_root[None]
        hand_pre_points = hand_pre_points * hand_sdf_scale
        obj_pre_points[:, :3] = obj_pre_points[:, :3] - obj_center_cam[None]
        obj_pre_points = obj_pre_points * obj_sdf_scale
        inputs.update(hand_pre_points=hand_pre_points[:, :3], obj_pre_points=obj_pre_points[:, :3])
    return inputs, {"hand_sdf": hand_sdf_points[:, 3], "obj_sdf": obj_sdf_points[:, 4]}


def synthetic_sdf_frame(seed, n_hand, n_obj, train=True, dist=0.02):
    """A packed SDF file in upstream's layout (tool/pre_process_sdf.py:140-147: hand rows then object rows,
    [x, y, z, sdf_hand, sdf_obj, label] float32) + the draws of ho3d.py:462-482 from a seeded generator + a rotation.
    -> (rows, number of hand rows, all_idx, rot_mat, hand_root, obj_center_cam)"""
    rng = np.random.default_rng(2000 + seed)
    nh, no = int(rng.integers(3 * n_hand, 5 * n_hand)), int(rng.integers(3 * n_obj, 5 * n_obj))
    data = np.concatenate([rng.uniform(-0.15, 0.15, (nh + no, 3)), rng.normal(0, 0.03, (nh + no, 2)),
                           rng.integers(0, 6, (nh + no, 1))], axis=1).astype(np.float32)
    draws = [rng.choice(nh, n_hand, replace=False), nh + rng.choice(no, n_obj, replace=False)]
    if train:
        draws.append(rng.choice(np.where(np.abs(data[:nh, 3]) < dist)[0], n_hand, replace=False))
        draws.append(rng.choice(np.where(np.abs(data[nh:, 4]) < dist)[0] + nh, n_obj, replace=False))
    th = rng.uniform(-np.pi, np.pi)
    rot = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]]).astype(np.float32)
    root = rng.uniform(-0.1, 0.1, 3).astype(np.float32)
    centre = rng.uniform(-0.1, 0.1, 3).astype(np.float32)
    return data, nh, np.concatenate(draws).astype(np.int64), rot, root, centre


def synthetic_annotation(seed):
    """The per-frame annotation arrays `ho3d.Dataset.__getitem__` reads in training mode (ho3d.py:436-459), synthetic and
    deterministic: intrinsics, 2-D / 3-D hand joints, MANO parameters, projected / 3-D object corners, object pose."""
    _, K, _, p2d = synthetic_frame(seed)
    rng = np.random.default_rng(3000 + seed)
    joints_3d = rng.uniform(-0.1, 0.1, (21, 3)).astype(np.float32) + np.array([0, 0, 0.6], np.float32)
    uvw = joints_3d.dot(K.T)
    return {"cam_intr": K, "joints_uv": (uvw[:, :2] / uvw[:, 2:]).astype(np.float32), "joints_3d": joints_3d,
            "mano_param": rng.uniform(-0.5, 0.5, 61).astype(np.float32), "obj_p2d": p2d,
            "obj_p3d": rng.uniform(-0.1, 0.1, (21, 3)).astype(np.float32) + np.array([0, 0, 0.6], np.float32),
            "obj_rot": rng.uniform(-1, 1, 3).astype(np.float32), "obj_trans": np.array([0.02, -0.03, 0.6], np.float32),
            "obj_depth_mean_value": 0.7}


def synthetic_eval_annotation(seed):
    """One `meta/*.pkl` dict of an HO3D evaluation frame (the keys ho3d.py:606-631 reads) + the object's 3-D box corners,
    synthetic: OpenGL-axes object pose in front of the camera, a hand box, a root joint."""
    img, K, bbox_hand, _ = synthetic_frame(seed)
    rng = np.random.default_rng(4000 + seed)
    half = rng.uniform(0.03, 0.09, 3)
    signs = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=np.float64)
    corners = np.concatenate([signs * half, np.zeros((1, 3))]).astype(np.float32)          # 8 corners + the centre
    ann = {"camMat": K.astype(np.float64), "objName": ["003_cracker_box", "021_bleach_cleanser"][seed % 2],
           "objRot": rng.uniform(-1.5, 1.5, (3, 1)), "objTrans": np.array([rng.uniform(-0.08, 0.08), rng.uniform(-0.08, 0.08),
                                                                            -rng.uniform(0.5, 0.8)]),
           "handBoundingBox": [float(v) for v in bbox_hand], "handJoints3D": rng.uniform(-0.1, 0.1, 3) + np.array([0, 0, -0.6])}
    return img, ann, corners


def synthetic_dexycb_sample(seed, left=None):
    """One `sample_dict` entry of DexYCB in the layout `dexycb.Dataset.__getitem__` reads (dexycb.py:411-418,434-437,487-498) +
    the arrays the dataset object holds (MANO PCA components / mean, the object's 3-D box corners), synthetic and deterministic.
    -> (frame (480, 640, 3) uint8, hand mask, obj mask, sample_info dict, holders dict)."""
    img, hand_mask, obj_mask = synthetic_aug(seed)[:3]
    _, K, _, _ = synthetic_frame(seed)
    rng = np.random.default_rng(5000 + seed)
    left = bool(seed % 2) if left is None else left
    joints_3d = rng.uniform(-0.09, 0.09, (21, 3)) + np.array([rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), 0.65])
    uvw = joints_3d.dot(K.astype(np.float64).T)
    half = rng.uniform(0.03, 0.08, 3)
    signs = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=np.float64)
    corners = np.concatenate([signs * half, np.zeros((1, 3))]).astype(np.float32)
    th = rng.uniform(-1, 1, 3)
    import cv2
    R = cv2.Rodrigues(th)[0]
    pose_y = np.concatenate([R, np.array([[rng.uniform(-0.06, 0.06)], [rng.uniform(-0.06, 0.06)], [0.7]])], axis=1)
    info = {"mano_side": "left" if left else "right", "color_file": "frame_%d.jpg.png" % seed,
            "intrinsics": {"fx": float(K[0, 0]), "fy": float(K[1, 1]), "ppx": float(K[0, 2]), "ppy": float(K[1, 2])},
            "pose_m": rng.uniform(-0.6, 0.6, (1, 51)).tolist(), "mano_betas": rng.uniform(-1, 1, 10).tolist(),
            "joint_3d": joints_3d[None].tolist(), "joint_2d": (uvw[:, :2] / uvw[:, 2:])[None].tolist(),
            "pose_y": [np.zeros((3, 4)).tolist(), pose_y.tolist()], "ycb_grasp_ind": 1, "ycb_ids": [7, 3 + seed % 5]}
    holders = {"components_right": rng.standard_normal((45, 45)).astype(np.float32) * 0.2,
               "components_left": rng.standard_normal((45, 45)).astype(np.float32) * 0.2,
               "handmean": rng.uniform(-0.2, 0.2, 45).astype(np.float32), "obj_bbox3d": {info["ycb_ids"][1]: corners}}
    return img, hand_mask, obj_mask, info, holders
