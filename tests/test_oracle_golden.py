"""Pin the CPU oracle (oracle/hoisdf_oracle.py) to outputs of the UNMODIFIED upstream code (tests/golden/*.npz,
written by oracle/make_golden.py in the build container).  Weights and inputs are regenerated from the recorded
seeds.  Tolerances: 2e-5 relative for floats (different BLAS / ISA between machines), exact for indices and masks."""
import os

import numpy as np
import pytest
import torch

from hoisdf_b200 import synthetic as syn
from oracle import hoisdf_oracle as O
from util import align_selection, aligned

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: z[k] for k in z.files}


def close(a, b, tol=2e-5):
    a, b = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(b)).double()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))
    assert err < tol, err


def test_stage_fixtures():
    g = load("stages_dexycb_seed7")
    arch, seed = str(g["arch"]), int(g["seed"])
    sd = syn.hot_path_state_dict(seed, arch)
    cfg = O.default_cfg(num_samp_hand=24, num_samp_obj=8)
    pyr, meta = syn.feature_pyramid(seed, 2, arch), syn.camera_meta(seed, 2)
    t = lambda k: torch.from_numpy(g[k])  # noqa
    with torch.no_grad():
        close(O.sdf_decoder(sd, "hand_sdf_decoder", t("sdf_decoder_in")), g["sdf_decoder_out"])
        sdf, _, pe = O.sdf_forward(sd, pyr, t("points"), meta["obj_center_cam"], meta["cam_intr"], 3.1, "obj", cfg)
        close(sdf, g["sdf_forward_sdf"]); close(pe, g["sdf_forward_posenc"])
        lat, cam = O.get_input_transformer(sd, pyr, t("points"), meta["mano_root"], meta["cam_intr"], 3.1, cfg)
        close(lat, g["point_latent"]); close(cam, g["point_cam"])
        src = t("transformer_src")
        assert torch.equal(O.mano_tgt_mask(cfg), t("tgt_mask")) and torch.equal(O.mano_memory_mask(cfg), t("memory_mask"))
        hs, mem, inter = O.transformer(sd, "hand_transformer", src, sd["mano_query_embed.weight"], torch.zeros_like(src),
                                       O.mano_tgt_mask(cfg), O.mano_memory_mask(cfg), cfg)
        close(hs, g["transformer_hs"]); close(mem, g["transformer_memory"])
        close(inter[[0, 5]], g["transformer_inter_first_last"])
        verts, joints = O.mano_head(sd, t("mano_pose6d"), t("mano_shape"))
        close(verts, g["mano_verts"]); close(joints, g["mano_joints"])
        close(O.vote_joints(t("vote_points"), t("vote_off"), t("vote_cls")), g["vote_joints"])
        # the sheared lattice, bit for bit
        assert np.array_equal(O.lattice(64)[torch.from_numpy(g["lattice_rows"])].numpy(), g["lattice_values"])


@pytest.mark.parametrize("name", ["hot_path_dexycb_seed11", "hot_path_ho3d_seed12"])
def test_hot_path_fixtures(name):
    g = load(name)
    arch, seed, B = str(g["arch"]), int(g["seed"]), int(g["batch"])
    ph, po = int(g["num_samp_hand"]), int(g["num_samp_obj"])
    sd = syn.hot_path_state_dict(seed, arch)
    pyr, meta = syn.feature_pyramid(seed, B, arch), syn.camera_meta(seed, B)
    taps = {}
    with torch.no_grad():
        out = O.hot_path_eval(sd, pyr, meta, O.default_cfg(num_samp_hand=ph, num_samp_obj=po), taps)
    # selected lattice indices: identical sets; order identical up to swaps of |sdf| near-ties (see util.py)
    ph_perm = align_selection(taps["hand"]["index"], g["hand_index"], g["hand_sdf"])
    po_perm = align_selection(taps["obj"]["index"], g["obj_index"], g["obj_sdf"])
    close(aligned(taps["hand_sdf"], ph_perm), g["hand_sdf"], 1e-3)
    close(aligned(taps["obj_sdf"], po_perm), g["obj_sdf"], 1e-3)
    close(aligned(taps["hand_posenc"], ph_perm), g["hand_posenc"])
    close(aligned(taps["obj_posenc"], po_perm), g["obj_posenc"])
    for k in ("mano_mesh_out", "mano_joints_out", "hand_joints_out"):      # global outputs: order-invariant
        close(out[k], g[k])
    for k in ("obj_rot_out", "obj_trans_out"):                            # per-point outputs: align first
        close(aligned(out[k], po_perm), g[k])


def test_image_fixture():
    g = load("image_dexycb_seed13")
    arch, seed, B = str(g["arch"]), int(g["seed"]), int(g["batch"])
    sd = syn.full_state_dict(seed, arch)
    taps = {}
    with torch.no_grad():
        out = O.model_eval(sd, syn.image_batch(seed, B), syn.camera_meta(seed, B),
                           O.default_cfg(num_samp_hand=int(g["num_samp_hand"]), num_samp_obj=int(g["num_samp_obj"])),
                           arch, taps)
    for lvl in O.LEVELS:
        close(taps["pyramid"][lvl][:, :4, :4, :4], g["pyr_crop_" + lvl], 1e-4)
        assert abs(float(taps["pyramid"][lvl].double().mean()) - float(g["pyr_mean_" + lvl])) < 1e-5 * float(g["pyr_abs_" + lvl])
    close(taps["decoder_out"][:, :4, :4, :4], g["pyr_crop_decoder_out"], 1e-4)
    for k in ("mano_mesh_out", "mano_joints_out", "obj_rot_out", "obj_trans_out", "hand_joints_out"):
        close(out[k], g[k], 1e-3)      # through the conv stack and a top-k: looser across machines


DEX_LOSS_KEYS = ("sdfhand_loss", "sdfobj_loss", "joint_heatmap", "obj_seg", "hand_seg", "loss_joint_3d",
                 "loss_joint_cls", "loss_all_joint_3d", "mano_mesh_loss", "mano_joint_loss", "pose_param_loss",
                 "shape_param_loss", "obj_rot", "obj_trans")
DEX_OUT_KEYS = ("joint_heatmap_out", "hand_seg_pred_out", "obj_seg_pred_out", "mano_mesh_out", "mano_joints_out",
                "mano_joints_gt_out", "mano_mesh_gt_out", "obj_rot_out", "obj_trans_out", "hand_joints_out")


def test_dexycb_eval_fixture():
    """The dexycb DATASET branch of the eval forward (upstream model.py:370-422,606-654): every output and loss."""
    g = load("dexycb_eval_seed14")
    seed, B = int(g["seed"]), int(g["batch"])
    ph, po = int(g["num_samp_hand"]), int(g["num_samp_obj"])
    sd = syn.full_state_dict(seed, "dexycb")
    inputs, targets = syn.dexycb_extras(seed, B, ph, po)
    before = targets["mano_param"].clone()
    with torch.no_grad():
        out = O.model_eval_dexycb(sd, syn.image_batch(seed, B), inputs, targets, syn.camera_meta(seed, B),
                                  O.default_cfg(dataset="dexycb", num_samp_hand=ph, num_samp_obj=po))
    assert torch.equal(before, targets["mano_param"])       # upstream works on a copy of the pose slice
    assert set(DEX_LOSS_KEYS + DEX_OUT_KEYS) <= set(out) and set(g) >= set(DEX_LOSS_KEYS + DEX_OUT_KEYS)
    for k in DEX_LOSS_KEYS + DEX_OUT_KEYS:
        close(out[k], g[k], 1e-3 if k in ("obj_rot_out", "obj_trans_out", "hand_joints_out", "loss_joint_3d",
                                            "loss_all_joint_3d", "loss_joint_cls", "obj_rot", "obj_trans") else 1e-4)


def test_metrics_fixture():
    """Test-time metrics (upstream common/metrics.py:62-248): the oracle's per-sample restatement reproduces the batch
    results of the upstream functions on both dataset branches, and its helpers per sample."""
    g = load("metrics_seed15")
    seed, B = int(g["seed"]), int(g["batch"])
    m = syn.metric_inputs(seed, B)
    templates = torch.stack([t["verts"] for t in m["templates"]])
    ids = m["obj_cls_ids"] - 1
    adds, mme, mce, oce = O.obj_pose_metrics(templates, ids, m["out"]["obj_rot"], m["out"]["obj_trans"],
                                             m["targets"]["obj_rot"], m["targets"]["rel_obj_trans"])
    close(adds, g["adds"]); close(mce, g["mce"]); close(mme, g["mme"])
    close([adds.mean(), mce.mean(), oce.mean(), B], g["dexycb_result"])
    used = [i for i, n in enumerate(m["obj_cls_names"]) if n != "019_pitcher_base"]
    assert 0 < len(used) < B                      # the fixture exercises the HO3D exclusion (metrics.py:129-141)
    close([adds[used].mean(), mme[used].mean(), len(used)], g["ho3d_result"])
    mje, pamje = O.hand_joint_metrics(m["joints_pred"], m["joints_gt"])
    close([mje.mean(), pamje.mean()], g["hand_joint_result"])
    close(O.rigid_align(m["joints_pred"][0].numpy(), m["joints_gt"][0].numpy()), g["aligned0"])


def test_train_step_fixture():
    """One training step (upstream Model.forward(mode="train") + the weighted loss sum of main/train.py:111-131 +
    backward): the oracle's autograd reproduces every loss entry, the outputs and the gradient of EVERY parameter tensor
    of the unmodified upstream model (fixture: abs-max, norm and 24 entries per tensor; tolerance relative to the largest
    gradient of the tensor's sub-network)."""
    from util import grad_summary, group_scales, oracle_train_step, param_group
    g = load("train_dexycb_seed21")
    seed, B = int(g["seed"]), int(g["batch"])
    ph, po = int(g["num_samp_hand"]), int(g["num_samp_obj"])
    out, parts, total, grads = oracle_train_step(seed, "dexycb", B, ph, po)
    total, parts = total.detach(), {k: v.detach() for k, v in parts.items()}
    assert abs(float(total) - float(g["total"])) <= 1e-5 * abs(float(g["total"]))
    for k, v in parts.items():
        assert abs(float(v) - float(g["loss." + k])) <= 1e-4 * max(abs(float(g["loss." + k])), 1e-3), k
    for k in ("joint_heatmap_out", "hand_seg_pred_out", "obj_seg_pred_out", "mano_mesh_out", "mano_joints_out",
              "hand_joints_out"):
        close(out[k].detach(), g[k], 1e-4)
    ref = {k[len("grad."):]: torch.from_numpy(g[k]) for k in g if k.startswith("grad.")}
    # upstream leaves the parameters its graph never reaches without a gradient (norm1, linear_objvote, linear_objcls) and
    # freezes the backbone's BatchNorm affine parameters (model.py:117-121): none of them is in the fixture
    assert not any(n.startswith(("norm1.", "linear_objvote.", "linear_objcls.")) or ("backbone" in n and ".bn" in n)
                   for n in ref)
    scales = group_scales(ref)
    assert set(ref) <= set(grads)
    for n, r in ref.items():
        s = grad_summary(grads[n])
        tol = 2e-4 * scales[param_group(n)]
        assert float((s[2:] - r[2:]).abs().max()) <= tol, (n, float((s[2:] - r[2:]).abs().max()), tol)
        assert abs(float(s[0] - r[0])) <= tol and abs(float(s[1] - r[1])) <= 2e-4 * max(float(r[1]), scales[param_group(n)]), n
