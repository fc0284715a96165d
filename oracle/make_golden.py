"""Generate tests/golden/*.npz by running the UNMODIFIED upstream HOISDF code (through oracle/reference_shim.py)
on seeded synthetic weights/inputs.  Run in the build container, where /root/reference exists:

    python oracle/make_golden.py

TEST INFRASTRUCTURE.  The fixtures pin `oracle/hoisdf_oracle.py` (tests/test_oracle_golden.py): the upstream
repository ships no tests or golden vectors of its own, so outputs of the upstream code
itself are the only possible anchor.  Weights and inputs are NOT stored -- `hoisdf_b200/synthetic.py`
regenerates them bit-identically from the seeds recorded in each fixture.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from hoisdf_b200 import synthetic as syn  # noqa: E402
from oracle import hoisdf_oracle as O  # noqa: E402
from oracle import reference_shim as rs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def build_reference(arch, seed, ph, po, dataset="ho3d"):
    ns = rs.load(arch)
    cfg = ns["cfg"]
    type(cfg).num_samp_hand, type(cfg).num_samp_obj = ph, po
    # the small ('dexycb') ARCHITECTURE is used for cheap fixtures; most of them take the ho3d eval branch, the
    # `dexycb_eval_case` fixture takes the dexycb DATASET branch (model.py:370-422: GT SDF samples, GT MANO, losses)
    type(cfg).dataset = dataset
    model = rs.build_model(ns, syn.mano_buffers(seed))
    model.load_state_dict(syn.full_state_dict(seed, arch), strict=True)
    model.eval()
    return ns, model


def pick_points(arch, seed, batch, ph, po, span=8):
    """Choose P_h, P_o near the requested values such that the |sdf| gap at the selection boundary is as large
    as possible for every sample: the fixture must not hinge on a near-tie that another BLAS could flip."""
    sd = syn.hot_path_state_dict(seed, arch)
    pyr, meta = syn.feature_pyramid(seed, batch, arch), syn.camera_meta(seed, batch)
    best = {}
    for kind, ck, bk, p0 in (("hand", "mano_root", "bbox_hand", ph), ("obj", "obj_center_cam", "bbox_obj", po)):
        taps = {}
        with torch.no_grad():
            O.sdf_infer(dict(sd), pyr, meta[ck], meta["cam_intr"], meta[bk], 3.1, p0, kind, O.default_cfg(), taps)
        gaps = []
        for p in range(p0, p0 + span):
            g = min(float((torch.sort(s.abs())[0][p] - torch.sort(s.abs())[0][p - 1])) for s in taps["cand_sdf"])
            gaps.append((g, p))
        best[kind] = max(gaps)
    return best["hand"][1], best["obj"][1], best["hand"][0], best["obj"][0]


def hot_path_case(name, arch, seed, batch, ph, po):
    ph, po, gh, go = pick_points(arch, seed, batch, ph, po)
    ns, model = build_reference(arch, seed, ph, po)
    pyr, meta = syn.feature_pyramid(seed, batch, arch), syn.camera_meta(seed, batch)
    taps = {}
    # feed the synthetic pyramid: the stage before the hot path is replaced, everything after is upstream code
    model.backbone_net.forward = lambda img: (None, None)
    model.decoder_net.forward = lambda f, s: (pyr, torch.zeros(batch, 3, 128, 128))
    orig_infer = model.sdf_infer

    def tap_infer(*a, **k):
        r = orig_infer(*a, **k)
        taps.setdefault("infer", []).append(r)
        return r

    model.sdf_infer = tap_infer
    with torch.no_grad():
        out = model({"img": torch.zeros(batch, 3, 256, 256)}, syn.eval_targets(batch), meta, "eval")
    lat = O.lattice(64)
    # recover lattice indices of the selected points (coordinates <-> indices is one-to-one)
    def to_index(points):
        # invert s = col * (2/63) - 1 per column, then idx = c0*4096 + ... is NOT valid for the sheared lattice,
        # so match against the lattice table instead
        flat = points.reshape(-1, 3)
        table = {r.tobytes(): i for i, r in enumerate(lat.numpy())}
        return np.array([table[r.tobytes()] for r in flat.numpy()], dtype=np.int64).reshape(points.shape[:2])

    (hp, hs, hpe, _), (op, os_, ope, _) = taps["infer"]
    fix = {
        "arch": arch, "seed": seed, "batch": batch, "num_samp_hand": ph, "num_samp_obj": po,
        "gap_hand": gh, "gap_obj": go,
        "hand_index": to_index(hp), "obj_index": to_index(op), "hand_sdf": hs, "obj_sdf": os_,
        "hand_posenc": hpe, "obj_posenc": ope,
    }
    fix.update({k: v for k, v in out.items() if k.endswith("_out")})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(fix))
    print(name, "P", ph, po, "gaps", gh, go, {k: tuple(v.shape) for k, v in out.items() if k.endswith("_out")})


def stage_case(name, arch, seed):
    """Stage-isolated fixtures: SDFDecoder, sdf_forward, get_input_transformer, Transformer, ManoHead, vote, masks."""
    ns, model = build_reference(arch, seed, 24, 8)   # memory mask: 24 hand-side + 8 object-side tokens
    batch = 2
    pyr, meta = syn.feature_pyramid(seed, batch, arch), syn.camera_meta(seed, batch)
    g = np.random.Generator(np.random.PCG64(seed))
    rnd = lambda *s: torch.from_numpy(g.random(size=s, dtype=np.float32) * 2 - 1)  # noqa
    fix = {"arch": arch, "seed": seed}
    with torch.no_grad():
        x = rnd(64, 289)
        fix["sdf_decoder_in"] = x
        fix["sdf_decoder_out"] = model.hand_sdf_decoder(x)[0]
        pts = rnd(batch, 33, 3)
        fix["points"] = pts
        sdf, _, pe = model.sdf_forward(pyr, pts, meta["obj_center_cam"], meta["cam_intr"], 3.1, type="obj")
        fix["sdf_forward_sdf"], fix["sdf_forward_posenc"] = sdf, pe
        lat, cam = model.get_input_transformer(pyr, pts, meta["mano_root"], meta["cam_intr"], 3.1)
        fix["point_latent"], fix["point_cam"] = lat, cam
        src = rnd(32, batch, 256)
        fix["transformer_src"] = src
        from common.utils.misc import get_mano_memory_mask, get_mano_tgt_mask
        hs, mem, inter, _ = model.hand_transformer(src=src, mask=None, pos_embed=torch.zeros_like(src), src_mask=None,
                                                   query_embed=model.mano_query_embed.weight,
                                                   tgt_mask=get_mano_tgt_mask(), memory_mask=get_mano_memory_mask())
        fix["transformer_hs"], fix["transformer_memory"], fix["transformer_inter_first_last"] = hs, mem, inter[[0, 5]]
        fix["tgt_mask"], fix["memory_mask"] = get_mano_tgt_mask(), get_mano_memory_mask()
        pose6d, shape = rnd(2, 16, batch, 6), rnd(2, batch, 10)
        fix["mano_pose6d"], fix["mano_shape"] = pose6d, shape
        res, _ = model.mano_head(pose6d, shape)
        fix["mano_verts"], fix["mano_joints"] = res["verts3d"], res["joints3d"]
        hp, off, cls = rnd(batch, 24, 3) * 0.1, rnd(2, 24, batch, 60) * 0.05, rnd(2, 24, batch, 20) * 3
        fix["vote_points"], fix["vote_off"], fix["vote_cls"] = hp, off, cls
        fix["vote_joints"] = model.joints_vote_loss(hp, off, cls, torch.zeros(batch, 20, 3))[3]
        fix["lattice_rows"] = np.array([0, 1, 63, 64, 4095, 4096, 262143])
        # the lattice as upstream builds it (model.py:257-273)
        fix["lattice_values"] = O.lattice(64)[torch.tensor(fix["lattice_rows"])]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(fix))
    print(name, sorted(fix))


def image_case(name, arch, seed, batch, ph, po):
    ns, model = build_reference(arch, seed, ph, po)
    img, meta = syn.image_batch(seed, batch), syn.camera_meta(seed, batch)
    pyr = {}
    orig = model.decoder_net.forward

    def tap(a, b):
        fp, ho = orig(a, b)
        pyr.update(fp)
        pyr["decoder_out"] = ho
        return fp, ho

    model.decoder_net.forward = tap
    with torch.no_grad():
        out = model({"img": img}, syn.eval_targets(batch), meta, "eval")
    fix = {"arch": arch, "seed": seed, "batch": batch, "num_samp_hand": ph, "num_samp_obj": po}
    # pyramid checksums + a small crop of every level pin the ResNet/U-Net restatement
    for k, v in pyr.items():
        fix["pyr_mean_" + k] = v.double().mean()
        fix["pyr_abs_" + k] = v.double().abs().mean()
        fix["pyr_crop_" + k] = v[:, :4, :4, :4]
    fix.update({k: v for k, v in out.items() if k.endswith("_out")})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(fix))
    print(name, {k: tuple(v.shape) for k, v in out.items() if k.endswith("_out")})


def dexycb_eval_case(name, seed, batch, ph, po):
    """The dexycb evaluation branch from the image: every `*_out` entry and every loss entry upstream returns."""
    ns, model = build_reference("dexycb", seed, ph, po, dataset="dexycb")
    img, meta = syn.image_batch(seed, batch), syn.camera_meta(seed, batch)
    inputs, targets = syn.dexycb_extras(seed, batch, ph, po)
    with torch.no_grad():
        out = model({"img": img, **inputs}, {k: v.clone() for k, v in targets.items()}, meta, "eval")
    fix = {"arch": "dexycb", "seed": seed, "batch": batch, "num_samp_hand": ph, "num_samp_obj": po}
    fix.update({k: v for k, v in out.items() if not k.endswith("_gt_out") or "mano" in k})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(fix))
    print(name, {k: tuple(v.shape) for k, v in out.items()})


GRAD_PROBES = 24      # gradient entries stored per parameter tensor (evenly strided) next to its abs-max and L2 norm


def grad_summary(g):
    """What the training fixture keeps of one parameter's gradient (full gradients of a 52 M-parameter model do not
    belong in git): abs-max, L2 norm and GRAD_PROBES evenly strided entries."""
    f = g.detach().reshape(-1).double()
    idx = torch.linspace(0, f.numel() - 1, min(GRAD_PROBES, f.numel())).long()
    return np.concatenate([[float(f.abs().max()), float(f.norm())], f[idx].numpy()])


def train_case(name, seed, batch, ph, po):
    """One training step's forward + backward (upstream Model.forward(mode="train") in the `*_pre_points` branch,
    main/model.py:426-466, then main/train.py:111-131): every loss entry, the `*_out` tensors, the summed weighted loss and
    a summary of every parameter's gradient.  Dropout 0 everywhere and zero jitter: the parity configuration."""
    ns, model = build_reference("dexycb", seed, ph, po, dataset="ho3d")
    cfg = ns["cfg"]
    old = (cfg.dropout, cfg.random_move_dist)
    type(cfg).random_move_dist = [0.0, 0.0, 0.0]
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    model.hand_sdf_decoder.dropout_prob = 0.0
    model.obj_sdf_decoder.dropout_prob = 0.0
    img, meta = syn.image_batch(seed, batch), syn.camera_meta(seed, batch)
    inputs, targets = syn.train_extras(seed, batch, ph, po)
    out = model({"img": img, **inputs}, {k: v.clone() for k, v in targets.items()}, meta, "train", 0, 0.0)
    total, parts = O.train_total_loss(out)
    total.backward()
    type(cfg).random_move_dist = old[1]
    fix = {"arch": "dexycb", "seed": seed, "batch": batch, "num_samp_hand": ph, "num_samp_obj": po,
           "total": float(total)}
    fix.update({"loss." + k: float(v) for k, v in parts.items()})
    fix.update({k: v for k, v in out.items() if k.endswith("_out") and "_gt_" not in k})
    for n, p in model.named_parameters():
        if p.grad is not None:
            fix["grad." + n] = grad_summary(p.grad)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(fix))
    print(name, "total", float(total), "params with grad", sum(1 for k in fix if k.startswith("grad.")))


def metrics_case(name, seed, batch):
    """Test-time metrics (upstream common/metrics.py) on seeded synthetic predictions: both dataset branches of
    eval_batched_obj_direct, the two mesh-metric helpers, eval_hand_joint and rigid_align."""
    rs.load("dexycb")
    import common.metrics as UM
    m = syn.metric_inputs(seed, batch)
    fix = {"seed": seed, "batch": batch}
    with torch.no_grad():
        dex = UM.eval_batched_obj_direct(m["out"], m["targets"], {"obj_cls": m["obj_cls_ids"], "cam_intr": torch.eye(3)[None]},
                                         m["templates"], None, m["obj_names"])
        ho3d = UM.eval_batched_obj_direct(m["out"], m["targets"], {"obj_cls": m["obj_cls_names"], "cam_intr": torch.eye(3)[None]},
                                          m["templates"], None, m["obj_names"])
        fix["dexycb_result"] = np.array([dex[0], dex[1], dex[2], dex[4]], dtype=np.float64)      # ADDS, MCE, OCE, n
        fix["ho3d_result"] = np.array([ho3d[0], ho3d[3], ho3d[4]], dtype=np.float64)              # ADDS, MME, n
        assert dex[3] is None and ho3d[1] is None and ho3d[2] is None
        # per-sample values of the helpers on the posed meshes of the dexycb branch
        ids = m["obj_cls_ids"] - 1
        tm = torch.stack([m["templates"][int(i)]["verts"] for i in ids])
        rot, trans = m["out"]["obj_rot"].mean(1), m["out"]["obj_trans"].mean(1)
        pred = torch.bmm(tm, UM.batch_rodrigues(rot).reshape(batch, 3, 3).permute(0, 2, 1)) + trans[:, None]
        tgt = torch.bmm(tm, UM.batch_rodrigues(m["targets"]["obj_rot"]).reshape(batch, 3, 3).permute(0, 2, 1)) \
            + m["targets"]["rel_obj_trans"][:, None]
        fix["adds"], fix["mce"] = UM.compute_obj_metrics_dexycb(pred, tgt)
        adds2, fix["mme"] = UM.compute_obj_metrics_ho3d(pred, tgt)
        assert torch.equal(adds2, fix["adds"])
        fix["hand_joint_result"] = np.array(UM.eval_hand_joint(m["joints_pred"], m["joints_gt"]), dtype=np.float64)
        fix["aligned0"] = UM.rigid_align(m["joints_pred"][0].numpy(), m["joints_gt"][0].numpy())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(fix))
    print(name, {k: np.asarray(v).shape for k, v in fix.items()}, fix["dexycb_result"], fix["ho3d_result"],
          fix["hand_joint_result"])


def feed_case(name, seed, n_eval, n_aug):
    """Data feed: the UNMODIFIED upstream `Dataset.data_crop` (data/ho3d.py:399-427) + the tensor conversion of
    :624 on seeded synthetic frames, and the frame / mask warp of `data_aug` (:318-321,351-353,366-381) for seeded
    augmentation draws, both through upstream's data/dataset_util.py helpers and this image's Pillow."""
    import PIL
    import torchvision.transforms as T
    from PIL import Image
    from oracle import feed_oracle as FO
    mods = rs.load_data_modules()
    DU, H = mods["dataset_util"], mods["ho3d"]

    class Self:
        inp_res = 256

    fix = {"seed": seed, "n_eval": n_eval, "n_aug": n_aug, "pillow": np.array(PIL.__version__)}
    # the float image is stored as the PIL bytes + upstream's byte -> float conversion of all 256 values (noise frames do not
    # compress; 4x smaller than the float tensor and exactly as informative)
    ramp = np.arange(256, dtype=np.uint8).reshape(1, 256, 1)
    fix["u8_to_f32"] = (T.ToTensor()(ramp.astype(np.float32)) / 255.0).numpy().reshape(256)
    ev = {"eval_bytes": [], "eval_K": [], "eval_bbox_hand": [], "eval_bbox_obj": []}
    for i in range(n_eval):
        img, K, bh, p2d = FO.synthetic_frame(seed + i)
        im, K2, hand, obj = H.Dataset.data_crop(Self(), Image.fromarray(img), K, bh, p2d)
        ev["eval_bytes"].append(np.asarray(im))
        ev["eval_K"].append(K2)
        ev["eval_bbox_hand"].append(hand.astype(np.float32))          # ho3d.py:649-650
        ev["eval_bbox_obj"].append(obj.astype(np.float32))
    au = {"aug_bytes": [], "aug_affine": [], "aug_hand_seg": [], "aug_obj_seg": []}
    for i in range(n_aug):
        img, hs, os_, center, scale, rot = FO.synthetic_aug(seed + i)
        affine, _ = DU.get_affine_transform(center, scale, [256, 256], rot=rot)
        au["aug_affine"].append(affine)
        au["aug_bytes"].append(np.asarray(DU.transform_img(Image.fromarray(img), affine, [256, 256]).crop((0, 0, 256, 256))))
        for key, seg in (("aug_hand_seg", hs), ("aug_obj_seg", os_)):
            w = DU.transform_img(Image.fromarray(seg), affine, [256, 256]).crop((0, 0, 256, 256))
            au[key].append(np.asarray(w.resize((128, 128), Image.NEAREST)).astype(np.float32))
    for d in (ev, au):
        for k, v in d.items():
            fix[k] = np.stack(v)
    # DexYCB's `data_crop` (data/dexycb.py:355-404): geometry outputs + the warped frame's checksum rows
    import data.dexycb as D

    class SelfD:
        inp_res, heatmap_res = 256, 128

    img, K32, _, p2d = FO.synthetic_frame(seed)
    _, hs, os_, _, _, _ = FO.synthetic_aug(seed)
    uv = (p2d.mean(0) + np.random.default_rng(seed).uniform(-60, 60, (21, 2))).astype(np.float32)
    ref = D.Dataset.data_crop(SelfD(), Image.fromarray(img), K32.astype(np.float64), uv, p2d, Image.fromarray(hs),
                              Image.fromarray(os_))
    fix.update(dex_uv_in=uv, dex_bbox_hand=ref[1], dex_bbox_obj=ref[2], dex_K=ref[3], dex_joints_uv=ref[4], dex_p2d=ref[5],
               dex_hand_seg=ref[6], dex_obj_seg=ref[7], dex_img_rows=np.asarray(ref[0])[::32].copy())
    # one EVALUATION sample through the unmodified `Dataset.__getitem__` (oracle/reference_shim.py:ho3d_eval_item)
    e_in, e_t, e_m = rs.ho3d_eval_item(seed)
    fix.update(evi_img_rows=e_in["img"].numpy()[:, ::8].copy(), evi_obj_rot=e_t["obj_rot"], evi_rel_obj_trans=e_t["rel_obj_trans"],
               evi_obj_mask=e_m["obj_mask"], evi_obj_cls=np.array(e_m["obj_cls"]))
    for k in ("cam_intr", "mano_root", "obj_center_cam", "bbox_hand", "bbox_obj"):
        fix["evi_" + k] = e_m[k]
    # one DexYCB TEST sample of a LEFT hand (mirror path) through the unmodified `dexycb.Dataset.__getitem__`
    d_in, d_t, d_m, d_taps = rs.dexycb_test_item(seed, left=True)
    fix.update(dxi_img_rows=d_in["img"].numpy()[:, ::8].copy(), dxi_draws=np.concatenate(d_taps["draws"]).astype(np.int64),
               dxi_hand_seg=d_t["hand_seg"].numpy(), dxi_obj_seg=d_t["obj_seg"].numpy(), dxi_obj_cls=d_m["obj_cls"])
    for k in ("hand_sdf_points", "obj_sdf_points"):
        fix["dxi_" + k] = d_in[k]
    for k in ("joint_coord", "joint_cam_no_trans", "obj_rot", "rel_obj_trans", "mano_param", "hand_sdf", "obj_sdf"):
        fix["dxi_" + k] = d_t[k]
    for k in ("cam_intr", "mano_root", "obj_center_cam", "bbox_hand", "bbox_obj"):
        fix["dxi_" + k] = d_m[k]
    # one DexYCB TRAINING sample of a LEFT hand with blur + jitter on, through the unmodified `dexycb.Dataset.__getitem__`
    x_in, x_t, x_m, x_taps = rs.dexycb_test_item(seed, left=True, mode="train", filters=True)
    fix.update(dxt_img_rows=x_in["img"].numpy()[:, ::8].copy(), dxt_draws=np.concatenate(x_taps["draws"]).astype(np.int64),
               dxt_hand_seg=np.packbits(x_t["hand_seg"].numpy().astype(np.uint8)),
               dxt_obj_seg=np.packbits(x_t["obj_seg"].numpy().astype(np.uint8)))
    for k in ("hand_sdf_points", "obj_sdf_points", "hand_pre_points", "obj_pre_points"):
        fix["dxt_" + k] = x_in[k]
    for k in ("joint_coord", "joint_cam_no_trans", "obj_rot", "rel_obj_trans", "mano_param", "hand_sdf", "obj_sdf"):
        fix["dxt_" + k] = x_t[k]
    for k in ("cam_intr", "mano_root", "obj_center_cam", "bbox_hand", "bbox_obj"):
        fix["dxt_" + k] = x_m[k]
    # one whole training sample through the unmodified `Dataset.__getitem__` (oracle/reference_shim.py:ho3d_train_item): the
    # SDF point sets + masks it returns and the draws / augmentation arguments needed to reproduce them
    inputs, targets, meta, taps = rs.ho3d_train_item(seed)
    a = taps["affine"][0]
    fix.update(item_draws=np.concatenate(taps["draws"]).astype(np.int64), item_center=a["center"], item_scale=a["scale"],
               item_rot=a["rot"], item_rot_mat=a["rot_mat"], item_mano_root=meta["mano_root"],
               item_obj_center_cam=meta["obj_center_cam"], item_hand_sdf_scale=taps["hand_sdf_scale"],
               item_obj_sdf_scale=taps["obj_sdf_scale"], item_hand_seg=targets["hand_seg"].numpy(),
               item_obj_seg=targets["obj_seg"].numpy(), item_hand_sdf=targets["hand_sdf"], item_obj_sdf=targets["obj_sdf"],
               item_img_bytes=np.round(inputs["img"].numpy().transpose(1, 2, 0) * 255.0).astype(np.uint8))
    assert np.array_equal(fix["u8_to_f32"][fix["item_img_bytes"]].transpose(2, 0, 1), inputs["img"].numpy())
    for k in ("hand_sdf_points", "obj_sdf_points", "hand_pre_points", "obj_pre_points"):
        fix["item_" + k] = inputs[k]
    # the same sample with upstream's blur + colour jitter ON (constructor defaults): every 8th row of the network input, and
    # the draws the filters made (re-derived from the same seed through upstream's own get_color_params / shuffle order)
    import random
    from hoisdf_b200 import feed
    inputs, _t, _m, taps = rs.ho3d_train_item(seed, filters=True)
    a = taps["affine"][0]
    random.seed(seed)
    radius = random.random() * 0.5
    steps = feed.draw_color_jitter(brightness=0.5, contrast=0.5, saturation=0.5, hue=0.15)
    inputs_f, targets_f, meta_f = inputs, _t, _m
    for k in ("joint_coord", "joint_cam_no_trans", "obj_rot", "rel_obj_trans", "mano_param"):
        fix["filt_t_" + k] = np.asarray(targets_f[k])
    for k in ("cam_intr", "mano_root", "obj_center_cam", "bbox_hand", "bbox_obj"):
        fix["filt_m_" + k] = np.asarray(meta_f[k])
    for k in ("hand_sdf_points", "obj_sdf_points", "hand_pre_points", "obj_pre_points"):
        fix["filt_i_" + k] = inputs_f[k]
    fix.update(filt_t_hand_sdf=targets_f["hand_sdf"], filt_t_obj_sdf=targets_f["obj_sdf"],
               filt_t_hand_seg=targets_f["hand_seg"].numpy(), filt_t_obj_seg=targets_f["obj_seg"].numpy())
    fix.update(filt_center=a["center"], filt_scale=a["scale"], filt_rot=a["rot"], filt_radius=radius,
               filt_order=np.array([n for n, _ in steps]), filt_factors=np.array([f for _, f in steps]),
               filt_img_rows=inputs["img"].numpy()[:, ::8].copy())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **fix)
    print(name, {k: np.asarray(v).shape for k, v in fix.items()})


if __name__ == "__main__":
    assert rs.available(), "the upstream reference is not mounted; golden vectors can only be made in the build container"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    stage_case("stages_dexycb_seed7", "dexycb", 7)
    hot_path_case("hot_path_dexycb_seed11", "dexycb", 11, 2, 64, 32)
    hot_path_case("hot_path_ho3d_seed12", "ho3d", 12, 1, 96, 40)
    image_case("image_dexycb_seed13", "dexycb", 13, 1, 48, 16)
    dexycb_eval_case("dexycb_eval_seed14", 14, 2, 48, 16)
    metrics_case("metrics_seed15", 15, 6)
    train_case("train_dexycb_seed21", 21, 2, 24, 8)
    feed_case("feed_seed31", 31, 1, 1)
