"""Seeded synthetic weights and inputs for the HOISDF hot path (no dataset, no checkpoint, no network).

Everything is drawn from numpy's PCG64 stream (bit-stable across machines and numpy versions), one
independent stream per tensor keyed by `(seed, crc32(name))`, so the golden-vector script (run in the
build container against the upstream reference), the CPU oracle and the B200 path all see bit-identical
weights and inputs without shipping 457 MB of parameters.

Shapes / key names follow the upstream state-dict contract (upstream
`main/model.py:41-90`, `common/nets/sdf_net.py:50-62`, `common/nets/transformer.py:257-366`,
`common/nets/module.py:147-165`, `common/nets/resnet.py:14-40`).  Input distributions follow
SURVEY.md section 8(d).
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import numpy as np
import torch

# (C_l, H_l) per pyramid level stride2..stride32 (upstream common/nets/module.py:172-218, :98-134)
PYRAMID_BIG = ((128, 128), (256, 64), (512, 32), (1024, 16), (2048, 8))
PYRAMID_SMALL = ((32, 128), (64, 64), (128, 32), (256, 16), (512, 8))
LEVEL_NAMES = ("stride2", "stride4", "stride8", "stride16", "stride32")


def pyramid_spec(arch: str):
    if arch == "ho3d":
        return PYRAMID_BIG
    if arch == "dexycb":
        return PYRAMID_SMALL
    raise ValueError("arch must be 'ho3d' or 'dexycb', got %r" % (arch,))


def multiscale_dim(arch: str) -> int:
    return sum(c for c, _ in pyramid_spec(arch))


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([int(seed), zlib.crc32(name.encode())]))


def _uniform(seed, name, shape, lo, hi):
    r = _rng(seed, name).random(size=tuple(shape), dtype=np.float32)
    return torch.from_numpy(r * np.float32(hi - lo) + np.float32(lo))


def _linear(sd, seed, prefix, out_f, in_f, bound=None):
    b = bound if bound is not None else 1.0 / np.sqrt(in_f)
    sd[prefix + ".weight"] = _uniform(seed, prefix + ".weight", (out_f, in_f), -b, b)
    sd[prefix + ".bias"] = _uniform(seed, prefix + ".bias", (out_f,), -b, b)


def _mlp(sd, seed, prefix, dims):
    for i in range(len(dims) - 1):
        _linear(sd, seed, "%s.layers.%d" % (prefix, i), dims[i + 1], dims[i])


def _layernorm(sd, seed, prefix, n):
    sd[prefix + ".weight"] = _uniform(seed, prefix + ".weight", (n,), 0.9, 1.1)
    sd[prefix + ".bias"] = _uniform(seed, prefix + ".bias", (n,), -0.05, 0.05)


def _mha(sd, seed, prefix, d):
    b = np.sqrt(6.0 / (d + 3 * d))  # xavier_uniform on the packed (3d, d) matrix
    sd[prefix + ".in_proj_weight"] = _uniform(seed, prefix + ".in_proj_weight", (3 * d, d), -b, b)
    sd[prefix + ".in_proj_bias"] = _uniform(seed, prefix + ".in_proj_bias", (3 * d,), -0.02, 0.02)
    b = np.sqrt(6.0 / (2 * d))
    sd[prefix + ".out_proj.weight"] = _uniform(seed, prefix + ".out_proj.weight", (d, d), -b, b)
    sd[prefix + ".out_proj.bias"] = _uniform(seed, prefix + ".out_proj.bias", (d,), -0.02, 0.02)


def _enc_layer(sd, seed, prefix, d, ffn):
    _mha(sd, seed, prefix + ".self_attn", d)
    _linear(sd, seed, prefix + ".linear1", ffn, d, np.sqrt(6.0 / (d + ffn)))
    _linear(sd, seed, prefix + ".linear2", d, ffn, np.sqrt(6.0 / (d + ffn)))
    _layernorm(sd, seed, prefix + ".norm1", d)
    _layernorm(sd, seed, prefix + ".norm2", d)


def _dec_layer(sd, seed, prefix, d, ffn):
    _mha(sd, seed, prefix + ".self_attn", d)
    _mha(sd, seed, prefix + ".multihead_attn", d)
    _linear(sd, seed, prefix + ".linear1", ffn, d, np.sqrt(6.0 / (d + ffn)))
    _linear(sd, seed, prefix + ".linear2", d, ffn, np.sqrt(6.0 / (d + ffn)))
    for n in ("norm1", "norm2", "norm3"):
        _layernorm(sd, seed, "%s.%s" % (prefix, n), d)


def _sdf_decoder(sd, seed, prefix, latent=256, point_feat=33):
    d_in = latent + point_feat  # 289
    dims = [(512, d_in), (512 - d_in, 512), (512, 512), (512, 512)]
    for i, (o, k) in enumerate(dims):
        p = "%s.linh%d" % (prefix, i)
        b = 1.0 / np.sqrt(k)
        v = _uniform(seed, p + ".weight_v", (o, k), -b, b)
        # weight_norm initialises g = ||v||_row; perturb it so the fold g*v/||v|| is exercised
        g = v.norm(dim=1, keepdim=True) * _uniform(seed, p + ".weight_g", (o, 1), 0.8, 1.25)
        sd[p + ".bias"] = _uniform(seed, p + ".bias", (o,), -b, b)
        sd[p + ".weight_g"] = g
        sd[p + ".weight_v"] = v
    _linear(sd, seed, prefix + ".linh4", 1, 512)


def _bn(sd, seed, prefix, c, gamma=1.0):
    sd[prefix + ".weight"] = _uniform(seed, prefix + ".weight", (c,), 0.9 * gamma, 1.1 * gamma)
    sd[prefix + ".bias"] = _uniform(seed, prefix + ".bias", (c,), -0.05, 0.05)
    sd[prefix + ".running_mean"] = _uniform(seed, prefix + ".running_mean", (c,), -0.05, 0.05)
    sd[prefix + ".running_var"] = _uniform(seed, prefix + ".running_var", (c,), 0.9, 1.1)
    sd[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def _conv(sd, seed, prefix, out_c, in_c, k, bias=False, gain=2.0, transposed=False):
    fan_in = in_c * k * k
    if transposed:
        # ConvTranspose2d(k=4, s=2, p=1): every output pixel receives 2x2 of the 4x4 taps
        fan_in = in_c * (k // 2) * (k // 2)
    b = np.sqrt(3.0 * gain / fan_in)
    shape = (in_c, out_c, k, k) if transposed else (out_c, in_c, k, k)
    sd[prefix + ".weight"] = _uniform(seed, prefix + ".weight", shape, -b, b)
    if bias:
        sd[prefix + ".bias"] = _uniform(seed, prefix + ".bias", (out_c,), -0.05, 0.05)


def _backbone(sd, seed, prefix="backbone_net.resnet"):
    _conv(sd, seed, prefix + ".conv1", 64, 3, 7)
    _bn(sd, seed, prefix + ".bn1", 64)
    inplanes = 64
    for li, (planes, blocks) in enumerate(((64, 3), (128, 4), (256, 6), (512, 3)), start=1):
        for bi in range(blocks):
            p = "%s.layer%d.%d" % (prefix, li, bi)
            _conv(sd, seed, p + ".conv1", planes, inplanes, 1)
            _bn(sd, seed, p + ".bn1", planes)
            _conv(sd, seed, p + ".conv2", planes, planes, 3)
            _bn(sd, seed, p + ".bn2", planes)
            _conv(sd, seed, p + ".conv3", planes * 4, planes, 1)
            _bn(sd, seed, p + ".bn3", planes * 4, gamma=0.5)
            if bi == 0:
                _conv(sd, seed, p + ".downsample.0", planes * 4, inplanes, 1, gain=1.0)
                _bn(sd, seed, p + ".downsample.1", planes * 4, gamma=0.7)
            inplanes = planes * 4


def _unet_decoder(sd, seed, arch, prefix="decoder_net.resnet_decoder"):
    def conv_bn(name, dims, k, final_bn=True):
        # upstream common/nets/layer.py:23-40 make_conv_layers: [conv, bn, relu] per stage
        idx = 0
        for i in range(len(dims) - 1):
            _conv(sd, seed, "%s.%s.%d" % (prefix, name, idx), dims[i + 1], dims[i], k, bias=True)
            idx += 1
            if i < len(dims) - 2 or final_bn:
                _bn(sd, seed, "%s.%s.%d" % (prefix, name, idx), dims[i + 1])
                idx += 2

    def deconv_bn(name, cin, cout):
        _conv(sd, seed, "%s.%s.0" % (prefix, name), cout, cin, 4, transposed=True)
        _bn(sd, seed, "%s.%s.1" % (prefix, name), cout)

    if arch == "ho3d":  # Decoder_big, upstream common/nets/module.py:147-170
        deconv_bn("deconv1", 2048, 1024)
        conv_bn("conv1", [2048, 1024], 3)
        deconv_bn("deconv2", 1024, 512)
        conv_bn("conv2", [1024, 512], 3)
        deconv_bn("deconv3", 512, 256)
        conv_bn("conv3", [512, 256], 3)
        deconv_bn("deconv4", 256, 128)
        conv_bn("conv4", [64 + 128, 128], 3)
        for n in ("convOut_hm", "convOut_hand_seg", "convOut_obj_seg"):
            conv_bn(n, [128, 128, 64, 1], 1, final_bn=False)
    else:  # Decoder (resnet50 branch), upstream common/nets/module.py:51-96
        conv_bn("conv0d", [2048, 512], 1)
        conv_bn("conv1d", [1024, 256], 1)
        deconv_bn("deconv1", 2048, 256)
        conv_bn("conv1", [512, 256], 3)
        conv_bn("conv2d", [512, 128], 1)
        deconv_bn("deconv2", 256, 128)
        conv_bn("conv2", [256, 128], 3)
        conv_bn("conv3d", [256, 64], 1)
        deconv_bn("deconv3", 128, 64)
        conv_bn("conv3", [128, 64], 3)
        conv_bn("conv4d", [64, 32], 1)
        deconv_bn("deconv4", 64, 64)
        conv_bn("conv4", [64 + 32, 32], 3)
        for n in ("convOut_hm", "convOut_hand_seg", "convOut_obj_seg"):
            conv_bn(n, [32, 32, 1], 1, final_bn=False)


def mano_buffers(seed: int) -> "OrderedDict[str, torch.Tensor]":
    """Synthetic stand-in for the licensed MANO_RIGHT.pkl (shapes: manopth/manopth/manolayer.py:74-100)."""
    b = OrderedDict()
    b["th_betas"] = torch.zeros(1, 10)
    b["th_shapedirs"] = _uniform(seed, "mano.shapedirs", (778, 3, 10), -0.004, 0.004)
    b["th_posedirs"] = _uniform(seed, "mano.posedirs", (778, 3, 135), -0.002, 0.002)
    b["th_v_template"] = _uniform(seed, "mano.v_template", (1, 778, 3), -0.09, 0.09)
    jr = _uniform(seed, "mano.J_regressor", (16, 778), 0.0, 1.0) ** 8
    b["th_J_regressor"] = jr / jr.sum(1, keepdim=True)
    w = _uniform(seed, "mano.weights", (778, 16), 0.0, 1.0) ** 12
    b["th_weights"] = w / w.sum(1, keepdim=True)
    b["th_faces"] = torch.zeros(1538, 3, dtype=torch.long)
    b["th_hands_mean"] = torch.zeros(1, 45)
    b["th_selected_comps"] = torch.eye(45)
    return b


def hot_path_state_dict(seed: int, arch: str = "ho3d") -> "OrderedDict[str, torch.Tensor]":
    """Every parameter/buffer of `Model` except `backbone_net.*` / `decoder_net.*`."""
    C = multiscale_dim(arch)
    d, ffn = 256, 1024
    sd = OrderedDict()
    sd["hand_sigmoid_beta"] = torch.full((1,), 0.1)
    sd["obj_sigmoid_beta"] = torch.full((1,), 0.08)
    _sdf_decoder(sd, seed, "hand_sdf_decoder")
    _sdf_decoder(sd, seed, "obj_sdf_decoder")
    for i in range(6):
        _enc_layer(sd, seed, "hand_transformer.encoder.layers.%d" % i, d, ffn)
    _layernorm(sd, seed, "hand_transformer.encoder.inter_norm", d)
    for i in range(4):
        _dec_layer(sd, seed, "hand_transformer.decoder.layers.%d" % i, d, ffn)
    _layernorm(sd, seed, "hand_transformer.decoder.norm", d)
    for i in range(3):
        _enc_layer(sd, seed, "obj_transformer.encoder.layers.%d" % i, d, ffn)
    _layernorm(sd, seed, "obj_transformer.encoder.inter_norm", d)
    _layernorm(sd, seed, "norm1", C)
    _mlp(sd, seed, "linear_transformerin", [C, 1024, 512, 256, 223])
    _mlp(sd, seed, "linear_sdfin", [C, 512, 256])
    sd["mano_query_embed.weight"] = _uniform(seed, "mano_query_embed.weight", (17, d), -1.7, 1.7)
    sd["mano_head.coord_change_mat"] = torch.tensor(
        [[1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.0, -1.0]])
    for k, v in mano_buffers(seed).items():
        sd["mano_head.mano_layer." + k] = v
    _mlp(sd, seed, "linear_pose", [d, d, d, 6])
    _mlp(sd, seed, "linear_shape", [d, d, d, 10])
    _mlp(sd, seed, "linear_handvote", [d, d, d, d, 60])
    _mlp(sd, seed, "linear_handcls", [d, d, d, 20])
    _mlp(sd, seed, "linear_objvote", [d, d, d, d, 24])
    _mlp(sd, seed, "linear_objcls", [d, d, d, 8])
    _mlp(sd, seed, "linear_obj_rel_trans", [d, d, d, 3])
    _mlp(sd, seed, "linear_obj_rot", [d, d, d, 3])
    return sd


def full_state_dict(seed: int, arch: str = "ho3d") -> "OrderedDict[str, torch.Tensor]":
    """Complete `Model.state_dict()` (upstream key names, loads with strict=True)."""
    sd = OrderedDict()
    _backbone(sd, seed)
    _unet_decoder(sd, seed, arch)
    sd.update(hot_path_state_dict(seed, arch))
    return sd


def camera_meta(seed: int, batch: int):
    """meta_info of SURVEY.md 8(d): intrinsics, hand root, object centre, xyxy boxes (all fp32)."""
    f = _uniform(seed, "meta.f", (batch,), 500.0, 700.0)
    K = torch.zeros(batch, 3, 3)
    K[:, 0, 0] = f
    K[:, 1, 1] = f
    K[:, 0, 2] = 127.5
    K[:, 1, 2] = 127.5
    K[:, 2, 2] = 1.0
    root = torch.stack([
        _uniform(seed, "meta.root.x", (batch,), -0.05, 0.05),
        _uniform(seed, "meta.root.y", (batch,), -0.05, 0.05),
        _uniform(seed, "meta.root.z", (batch,), 0.45, 0.65)], 1)
    obj = root + _uniform(seed, "meta.obj", (batch, 3), -0.05, 0.05)

    def box(center, name):
        uvw = torch.einsum("bij,bj->bi", K, center)
        uv = uvw[:, :2] / uvw[:, 2:3]
        half = _uniform(seed, name, (batch, 1), 80.0, 110.0)
        return torch.cat([uv - half, uv + half], 1).clamp_(0.0, 255.0)

    return {
        "cam_intr": K,
        "mano_root": root,
        "obj_center_cam": obj,
        "bbox_hand": box(root, "meta.hw_hand"),
        "bbox_obj": box(obj, "meta.hw_obj"),
    }


def image_batch(seed: int, batch: int):
    return _uniform(seed, "inputs.img", (batch, 3, 256, 256), 0.0, 1.0)


def feature_pyramid(seed: int, batch: int, arch: str = "ho3d", scale: float = 1.0):
    """A synthetic post-ReLU pyramid (dict of NCHW fp32) for stage-isolated hot-path tests."""
    out = OrderedDict()
    for name, (c, h) in zip(LEVEL_NAMES, pyramid_spec(arch)):
        x = _uniform(seed, "pyramid." + name, (batch, c, h, h), -1.0, 1.0)
        out[name] = torch.relu(x) * scale
    return out


def eval_targets(batch: int):
    return {"obj_rot": torch.zeros(batch, 3), "rel_obj_trans": torch.zeros(batch, 3)}


def dexycb_extras(seed: int, batch: int, ph: int, po: int):
    """The extra `inputs` / `targets` the dexycb evaluation branch reads (upstream main/model.py:370-422,606-629;
    dataset keys data/dexycb.py:627-655), shaped and distributed as SURVEY.md 8(d) states."""
    g = _rng(seed, "dexycb.normal")
    normal = lambda shape, std: torch.from_numpy((g.standard_normal(size=shape) * std).astype(np.float32))  # noqa
    inputs = {
        "hand_sdf_points": _uniform(seed, "dexycb.hand_sdf_points", (batch, ph, 3), -1.0, 1.0),
        "obj_sdf_points": _uniform(seed, "dexycb.obj_sdf_points", (batch, po, 3), -1.0, 1.0),
    }
    targets = {
        "hand_sdf": _uniform(seed, "dexycb.hand_sdf", (batch, ph), -0.1, 0.1),
        "obj_sdf": _uniform(seed, "dexycb.obj_sdf", (batch, po), -0.1, 0.1),
        "hand_seg": (_uniform(seed, "dexycb.hand_seg", (batch, 128, 128), 0.0, 1.0) > 0.5).float(),
        "obj_seg": (_uniform(seed, "dexycb.obj_seg", (batch, 128, 128), 0.0, 1.0) > 0.5).float(),
        "joint_coord": _uniform(seed, "dexycb.joint_coord", (batch, 21, 2), 0.0, 128.0),
        "joint_cam_no_trans": normal((batch, 21, 3), 40.0),
        "mano_param": normal((batch, 58), 0.1),
        "obj_rot": torch.zeros(batch, 3),
        "rel_obj_trans": torch.zeros(batch, 3),
    }
    return inputs, targets


def train_extras(seed: int, batch: int, ph: int, po: int):
    """`inputs` / `targets` of a training step (upstream data/ho3d.py:527-589: P_h + P_o SDF supervision points and as many
    `*_pre_points` per sample, all in normalised coordinates)."""
    inputs, targets = dexycb_extras(seed, batch, ph, po)
    inputs["hand_pre_points"] = _uniform(seed, "train.hand_pre_points", (batch, ph, 3), -0.3, 0.3)
    inputs["obj_pre_points"] = _uniform(seed, "train.obj_pre_points", (batch, po, 3), -0.3, 0.3)
    return inputs, targets


HO3D_OBJECT_NAMES = ("003_cracker_box", "006_mustard_bottle", "010_potted_meat_can", "019_pitcher_base", "021_bleach_cleanser")


def metric_inputs(seed: int, batch: int, votes: int = 40, n_templates: int = 5, n_verts: int = 1000):
    """Synthetic inputs of the test-time metrics (upstream common/metrics.py:110-248), shaped like main/test.py:85-135
    feeds them: a list of object templates ({"verts": (n_verts, 3)} in metres, `prepare_model_template`), the id -> name
    table, per-point pose votes `obj_rot` / `obj_trans` (B, votes, 3) around the ground truth, and joint sets (B, 21, 3)."""
    templates = [{"verts": _uniform(seed, "metrics.template%d" % i, (n_verts, 3), -0.1, 0.1) *
                  torch.tensor([1.0, 0.6 + 0.1 * i, 0.4])} for i in range(n_templates)]
    obj_names = {i + 1: HO3D_OBJECT_NAMES[i % len(HO3D_OBJECT_NAMES)] + ("" if i < len(HO3D_OBJECT_NAMES) else "_%d" % i)
                 for i in range(n_templates)}
    rot_gt = _uniform(seed, "metrics.rot_gt", (batch, 3), -2.0, 2.0)
    trans_gt = _uniform(seed, "metrics.trans_gt", (batch, 3), -0.2, 0.2)
    out = {
        "obj_rot": rot_gt[:, None] + _uniform(seed, "metrics.rot_noise", (batch, votes, 3), -0.3, 0.3),
        "obj_trans": trans_gt[:, None] + _uniform(seed, "metrics.trans_noise", (batch, votes, 3), -0.03, 0.03),
    }
    targets = {"obj_rot": rot_gt, "rel_obj_trans": trans_gt}
    ids = torch.from_numpy(_rng(seed, "metrics.obj_ids").integers(1, n_templates + 1, size=batch)).long()
    joints_gt = _uniform(seed, "metrics.joints_gt", (batch, 21, 3), -0.1, 0.1)
    joints_pred = joints_gt * 1.1 + _uniform(seed, "metrics.joints_noise", (batch, 21, 3), -0.02, 0.02) + 0.01
    return dict(templates=templates, obj_names=obj_names, out=out, targets=targets, obj_cls_ids=ids,
                obj_cls_names=[obj_names[int(i)] for i in ids], joints_pred=joints_pred, joints_gt=joints_gt)
