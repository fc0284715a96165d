"""Import the UNMODIFIED upstream HOISDF reference (read-only at /root/reference) on a GPU-less host.

TEST INFRASTRUCTURE ONLY.  This file exists so that (a) `oracle/make_golden.py` can generate the
committed golden vectors under `tests/golden/` and (b) `tests/test_oracle_vs_reference.py` can pin the
CPU restatement in `oracle/hoisdf_oracle.py` against the real thing.  Nothing in `hoisdf_b200/`
(the product) may import it, and it is never used on the GPU box (`/root/reference` does not exist
there -> `available()` is False and the callers skip).

Shims, all applied outside the read-only tree:
  1. `torchvision.models.resnet.model_urls = {}`  (common/nets/resnet.py:9 imports a removed name)
  2. `torch.Tensor.cuda = identity` on a host without CUDA (main/model.py:132,275-281,301-302;
     common/nets/sdf_net.py:122 hard-code `.cuda()`)
  3. a synthetic `ManoLayer` (manopth/manopth/manolayer.py:14 needs chumpy + the licensed MANO pkl):
     built with `ManoLayer.__new__` + the buffers `forward` reads (manolayer.py:74-106).
"""
from __future__ import annotations

import os
import sys
import tempfile

REFERENCE_ROOT = os.environ.get("HOISDF_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "main", "model.py"))


_loaded = {}


def load(setting: str = "ho3d"):
    """Return a namespace dict {cfg, M (main.model), modules...} for the reference."""
    if "ns" in _loaded:
        ns = _loaded["ns"]
        _apply_setting(ns["cfg"], setting)
        return ns
    if not available():
        raise RuntimeError("upstream reference not present at %s" % REFERENCE_ROOT)
    import torch
    import torchvision.models.resnet as tv_resnet

    if not hasattr(tv_resnet, "model_urls"):
        tv_resnet.model_urls = {}
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # shim 2
    # main/config.py:197 creates ./outputs/log relative to the CWD at import time
    scratch = tempfile.mkdtemp(prefix="hoisdf_ref_cwd_")
    old = os.getcwd()
    os.chdir(scratch)
    try:
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        # manopth's ManoLayer module imports chumpy-dependent code at import time; stub the loader
        import types

        stub = types.ModuleType("manopth.mano.webuser.smpl_handpca_wrapper_HAND_only")
        stub.ready_arguments = lambda *a, **k: (_ for _ in ()).throw(
            RuntimeError("MANO pkl not available in this container")
        )
        sys.modules.setdefault("manopth.mano.webuser.smpl_handpca_wrapper_HAND_only", stub)
        from main.config import cfg  # noqa

        _apply_setting(cfg, setting)
        import main.model as M  # noqa
        from manopth.manopth.manolayer import ManoLayer  # noqa
    finally:
        os.chdir(old)
    ns = {"cfg": cfg, "M": M, "ManoLayer": ManoLayer}
    _loaded["ns"] = ns
    return ns


def _apply_setting(cfg, setting):
    cls = type(cfg)
    if setting == "ho3d":
        cls.setting = "ho3d"
        cls.dataset = "ho3d"
        cls.use_big_decoder = True
    elif setting == "dexycb":
        cls.setting = "dexycb"
        cls.dataset = "dexycb"
        cls.use_big_decoder = False
    else:
        raise ValueError(setting)
    cls.use_inverse_kinematics = False
    cfg.calc_mutliscale_dim(cls.use_big_decoder, cls.resnet_type)


def synthetic_mano_layer(ns, mano_buffers):
    """Shim 3: a ManoLayer carrying caller-provided buffers (no pkl, no chumpy)."""
    import torch

    ManoLayer = ns["ManoLayer"]
    layer = ManoLayer.__new__(ManoLayer)
    torch.nn.Module.__init__(layer)
    layer.center_idx = 0
    layer.robust_rot = False
    layer.rot = 3
    layer.flat_hand_mean = True
    layer.side = "right"
    layer.use_pca = False
    layer.joint_rot_mode = "axisang"
    layer.root_rot_mode = "axisang"
    layer.ncomps = 45
    for k, v in mano_buffers.items():
        layer.register_buffer(k, v.clone())
    layer.kintree_parents = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]
    return layer


def build_model(ns, mano_buffers):
    """Mirror of main/model.py:683-760 (`get_model`) without the pkl-dependent ManoLayer."""
    cfg, M = ns["cfg"], ns["M"]
    backbone = M.BackboneNet()
    decoder = M.DecoderNet_big() if cfg.use_big_decoder else M.DecoderNet()
    hand_sdf = M.SDFDecoder(latent_size=cfg.hidden_dim, point_feat_size=cfg.PointFeatSize,
                            use_classifier=cfg.ClassifierBranch)
    obj_sdf = M.SDFDecoder(latent_size=cfg.hidden_dim, point_feat_size=cfg.PointFeatSize,
                           use_classifier=cfg.ClassifierBranch)
    hand_tr = M.Transformer(d_model=cfg.hidden_dim, dropout=cfg.dropout, nhead=cfg.nheads,
                            dim_feedforward=cfg.dim_feedforward, num_encoder_layers=cfg.enc_layers,
                            num_decoder_layers=cfg.dec_layers, normalize_before=cfg.pre_norm,
                            return_intermediate_dec=True)
    obj_tr = M.VoteTransformer(d_model=cfg.hidden_dim, dropout=cfg.dropout, nhead=cfg.nheads,
                               dim_feedforward=cfg.dim_feedforward,
                               num_encoder_layers=cfg.enc_layers // 2,
                               normalize_before=cfg.pre_norm, return_intermediate_dec=True)
    mano = synthetic_mano_layer(ns, mano_buffers)
    return M.Model(backbone, decoder, hand_sdf, obj_sdf, hand_tr, obj_tr, mano)


def load_data_modules():
    """Upstream `data.dataset_util` and `data.ho3d` (for the data-feed oracle).
    Shim 4: `libyana.meshutils.meshio` and `pytorch3d.io` (data/dataset_util.py:14,16: mesh readers, not on the crop path) are
    absent from this image and are replaced by empty modules; the functions under test never touch them."""
    if "data" in _loaded:
        return _loaded["data"]
    load("ho3d")
    import types

    for name, attrs in (("libyana", {}), ("libyana.meshutils", {}), ("libyana.meshutils.meshio", {}),
                        ("pytorch3d", {}), ("pytorch3d.io", {"load_obj": None})):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                mod = types.ModuleType(name)
                for k, v in attrs.items():
                    setattr(mod, k, v)
                sys.modules[name] = mod
    sys.modules["libyana"].meshutils = sys.modules["libyana.meshutils"]
    sys.modules["libyana.meshutils"].meshio = sys.modules["libyana.meshutils.meshio"]
    import data.dataset_util as DU
    import data.ho3d as H

    _loaded["data"] = {"dataset_util": DU, "ho3d": H}
    return _loaded["data"]


def ho3d_train_item(seed, n_hand=24, n_obj=8, filters=False):
    """ONE training sample through the UNMODIFIED upstream `data.ho3d.Dataset.__getitem__` (data/ho3d.py:432-589, mode "train"),
    on a synthetic frame written to a scratch directory (PNG + packed SDF .npy in upstream's layouts).  The dataset object is
    made with `__new__` (its constructor reads the HO3D tree, absent here) and given exactly the attributes `__getitem__` /
    `data_aug` read; with `filters=False` blur and colour jitter are switched off through upstream's own parameters (radius /
    ranges 0: PIL's GaussianBlur(0) and the empty jitter list are identities), with `filters=True` they run with the
    constructor's defaults (ho3d.py:34-41); every random draw is upstream's, from generators seeded with `seed`.
    Returns (inputs, targets, meta_info, taps): taps = the draws (`np.random.choice` results, affine arguments) and the raw
    arrays a restatement needs to reproduce the item."""
    import random

    import numpy as np
    from PIL import Image
    import torchvision.transforms as transforms

    from oracle import feed_oracle as FO

    mods = load_data_modules()
    DU, H = mods["dataset_util"], mods["ho3d"]
    scratch = tempfile.mkdtemp(prefix="hoisdf_feed_")
    img, hand_mask, obj_mask, _, _, _ = FO.synthetic_aug(seed)
    sdf, nh, _, _, _, _ = FO.synthetic_sdf_frame(seed, n_hand, n_obj)
    no = len(sdf) - nh
    ann = FO.synthetic_annotation(seed)
    Image.fromarray(img).save(os.path.join(scratch, "frame.png"))
    np.save(os.path.join(scratch, "sdf.npy"), sdf)
    ds = H.Dataset.__new__(H.Dataset)
    ds.mode = "train"
    ds.image_paths = [os.path.join(scratch, "frame.png")]
    ds.K = [ann["cam_intr"]]
    ds.joints_uv = [ann["joints_uv"]]
    ds.mano_params = [ann["mano_param"]]
    ds.joints_3d = [ann["joints_3d"]]
    ds.hand_segs = [np.packbits(hand_mask)]
    ds.obj_segs = [np.packbits(obj_mask)]
    ds.obj_p2ds = [ann["obj_p2d"]]
    ds.obj_p3ds = [ann["obj_p3d"]]
    ds.obj_rot_list = [ann["obj_rot"]]
    ds.obj_trans_list = [ann["obj_trans"]]
    ds.obj_cls_list = ["003_cracker_box"]
    ds.sdf_paths = [os.path.join(scratch, "sdf.npy")]
    ds.sdf_indexes = [np.array([nh, no])]
    ds.num_samp_hand, ds.num_samp_obj = n_hand, n_obj
    ds.dist = 0.02
    ds.hand_sdf_scale, ds.obj_sdf_scale = 6.2, 5.8
    ds.obj_depth_mean_value = ann["obj_depth_mean_value"]
    ds.inp_res, ds.heatmap_res = 256, 128          # main/config.py:111-112
    ds.transform = transforms.ToTensor()
    ds.coord_change_mat = np.array([[1.0, 0.0, 0.0], [0, -1.0, 0.0], [0.0, 0.0, -1.0]], dtype=np.float32)
    ds.hue = ds.contrast = ds.brightness = ds.saturation = 0
    ds.blur_radius = 0
    if filters:
        ds.hue, ds.saturation, ds.contrast, ds.brightness, ds.blur_radius = 0.15, 0.5, 0.5, 0.5, 0.5
    ds.scale_jittering, ds.center_jittering, ds.max_rot = 0.2, 0.1, np.pi
    taps = {"draws": [], "affine": [], "sdf": sdf, "n_hand_rows": nh, "frame": img, "hand_mask": hand_mask,
            "obj_mask": obj_mask, "hand_sdf_scale": ds.hand_sdf_scale, "obj_sdf_scale": ds.obj_sdf_scale}
    taps["raw"] = {"cam_intr": ds.K[0].copy(), "joints_uv": ds.joints_uv[0].copy(), "joints_3d": ds.joints_3d[0].copy(),
                   "mano_param": ds.mano_params[0].copy(), "obj_p2d": ds.obj_p2ds[0].copy(), "obj_p3d": ds.obj_p3ds[0].copy(),
                   "obj_rot": ds.obj_rot_list[0].copy(), "obj_trans": ds.obj_trans_list[0].copy(),
                   "obj_depth_mean_value": ds.obj_depth_mean_value}
    real_choice, real_affine = np.random.choice, DU.get_affine_transform

    def tap_choice(*a, **k):
        out = real_choice(*a, **k)
        taps["draws"].append(np.asarray(out).copy())
        return out

    def tap_affine(center, scale, res, rot=0, K=None):
        out = real_affine(center, scale, res, rot=rot, K=K)
        taps["affine"].append({"center": np.asarray(center).copy(), "scale": float(scale), "rot": float(rot),
                               "affinetrans": out[0].copy(), "rot_mat": out[-1].copy()})
        return out

    np.random.seed(seed)
    random.seed(seed)
    np.random.choice, DU.get_affine_transform = tap_choice, tap_affine
    try:
        inputs, targets, meta = ds[0]
    finally:
        np.random.choice, DU.get_affine_transform = real_choice, real_affine
        import shutil
        shutil.rmtree(scratch, ignore_errors=True)
    return inputs, targets, meta, taps


def ho3d_eval_item(seed):
    """ONE evaluation sample through the UNMODIFIED upstream `data.ho3d.Dataset.__getitem__` (the `else` branch, ho3d.py:591-660;
    what main/test.py's loader yields): a synthetic sequence directory (rgb/0000.png + meta/0000.pkl) in a scratch root, the
    dataset object made with `__new__` and given the attributes that branch reads.  Returns (inputs, targets, meta_info)."""
    import pickle
    import shutil

    import numpy as np
    from PIL import Image
    import torchvision.transforms as transforms

    from oracle import feed_oracle as FO

    H = load_data_modules()["ho3d"]
    img, ann, corners = FO.synthetic_eval_annotation(seed)
    root = tempfile.mkdtemp(prefix="hoisdf_feed_eval_")
    seq = os.path.join(root, "evaluation", "SEQ1")
    os.makedirs(os.path.join(seq, "rgb"))
    os.makedirs(os.path.join(seq, "meta"))
    Image.fromarray(img).save(os.path.join(seq, "rgb", "0000.png"))
    with open(os.path.join(seq, "meta", "0000.pkl"), "wb") as f:
        pickle.dump(ann, f)
    ds = H.Dataset.__new__(H.Dataset)
    ds.mode, ds.root, ds.set_list = "evaluation", root, ["SEQ1/0000"]
    ds.obj_bbox3d = {ann["objName"]: corners}
    ds.coord_change_mat = np.array([[1.0, 0.0, 0.0], [0, -1.0, 0.0], [0.0, 0.0, -1.0]], dtype=np.float32)
    ds.obj_depth_mean_value, ds.inp_res = 0.7, 256
    ds.transform = transforms.ToTensor()
    try:
        return ds[0]
    finally:
        shutil.rmtree(root, ignore_errors=True)


def dexycb_test_item(seed, n_hand=24, n_obj=8, left=None, mode="test", filters=False):
    """ONE sample through the UNMODIFIED upstream `data.dexycb.Dataset.__getitem__` in test mode (dexycb.py:409-657; the feed
    of BASELINE configs[2]'s evaluation): synthetic colour file + packed SDF .npy in a scratch directory, dataset object made
    with `__new__` and given the attributes that method reads.  Returns (inputs, targets, meta_info, taps) -- taps: the draws."""
    import shutil

    import numpy as np
    from PIL import Image
    import torchvision.transforms as transforms

    from oracle import feed_oracle as FO

    load_data_modules()
    import data.dexycb as D

    img, hand_mask, obj_mask, info, holders = FO.synthetic_dexycb_sample(seed, left)
    sdf, nh = FO.synthetic_sdf_frame(seed, n_hand, n_obj)[:2]
    scratch = tempfile.mkdtemp(prefix="hoisdf_feed_dex_")
    Image.fromarray(img).save(os.path.join(scratch, info["color_file"]))
    np.save(os.path.join(scratch, "sdf.npy"), sdf)
    ds = D.Dataset.__new__(D.Dataset)
    ds.mode = mode                      # "test", or "train": data_aug with the constructor's defaults (dexycb.py:29-40)
    ds.dist = 0.02
    ds.scale_jittering, ds.center_jittering, ds.max_rot = 0.2, 0.1, np.pi
    ds.hue = ds.contrast = ds.brightness = ds.saturation = ds.blur_radius = 0
    if filters:
        ds.hue, ds.saturation, ds.contrast, ds.brightness, ds.blur_radius = 0.15, 0.5, 0.5, 0.5, 0.5
    ds.sample_dict, ds.sample_list_processed = {"k": info}, ["k"]
    ds.image_fast_path = scratch
    ds.mano_handcomponent_right, ds.mano_handcomponent_left = holders["components_right"], holders["components_left"]
    ds.mano_handmean = holders["handmean"]
    ds.hand_segs, ds.obj_segs = [np.packbits(hand_mask)], [np.packbits(obj_mask)]
    ds.obj_bbox3d = holders["obj_bbox3d"]
    ds.sdf_path_list, ds.sdf_index_list = [os.path.join(scratch, "sdf.npy")], [np.array([nh, len(sdf) - nh])]
    ds.num_samp_hand, ds.num_samp_obj = n_hand, n_obj
    ds.inp_res, ds.heatmap_res = 256, 128          # main/config.py:111-112
    ds.hand_sdf_scale, ds.obj_sdf_scale = 6.2, 5.8
    ds.transform = transforms.ToTensor()
    taps = {"draws": [], "sdf": sdf, "n_hand_rows": nh}
    real_choice = np.random.choice

    def tap_choice(*a, **k):
        out = real_choice(*a, **k)
        taps["draws"].append(np.asarray(out).copy())
        return out

    import random
    np.random.seed(seed)
    random.seed(seed)
    np.random.choice = tap_choice
    try:
        inputs, targets, meta = ds[0]
    finally:
        np.random.choice = real_choice
        shutil.rmtree(scratch, ignore_errors=True)
    return inputs, targets, meta, taps
