"""CPU oracle for the HOISDF hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain, functional, fp32 PyTorch-CPU restatement of the upstream algorithm (amathislab/HOISDF @ 666e5b7)
for the path `sdf_infer -> sdf_forward -> get_input_transformer -> transformers -> heads -> MANO/vote`.
Every function cites the upstream file:line it follows.  It takes a flat parameter dict with the upstream
state-dict key names, so the same dict drives the reference, the oracle and the B200 path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this module (plus the developer measurement tools under `scripts/`, which use it as the checker next to a timing);
the product (`hoisdf_b200/`) never does and fails loudly without its CUDA library.

Pinning: the upstream repository has no tests or golden vectors for this path, so the
oracle is pinned against outputs of the upstream code itself, run in the build container through
`oracle/reference_shim.py`: `oracle/make_golden.py` writes `tests/golden/*.npz`,
`tests/test_oracle_golden.py` checks the oracle against them everywhere, and
`tests/test_oracle_vs_reference.py` re-checks live wherever `/root/reference` exists.

The arithmetic the upstream path delegates to PyTorch ATen (`F.grid_sample`, `nn.Linear`, `torch.sort`,
`nn.LayerNorm`, softmax, sin/cos/tanh) is called here as the same ATen CPU ops; what is restated by hand is
everything upstream wrote itself (lattice, projection, filtering, token assembly, weight-norm fold,
multi-head attention wiring, masks, heads, rot6d->axis-angle, MANO LBS, vote aggregation).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F

LEVELS = ("stride2", "stride4", "stride8", "stride16", "stride32")


def default_cfg(**over):
    """The subset of upstream main/config.py:38-151 that the hot path reads at call time."""
    c = SimpleNamespace(
        dataset="ho3d", bins_n=64, PointFeatSize=33, ClampingDistance=0.15,
        hand_sdf_scale=3.1, obj_sdf_scale=3.1, num_samp_hand=600, num_samp_obj=200,
        input_img_shape=(256, 256), hidden_dim=256, nheads=4, dim_feedforward=1024,
        enc_layers=6, dec_layers=4, mano_num_queries=17, mano_shape_indx=16,
        output_hm_shape=(128, 128, 128), sigma=2.5 / 2, hand_cls_dist=0.04,
        lambda_verts3d=1e4, lambda_joints3d=1e4, lambda_manopose=10, lambda_manoshape=0.1,
    )
    for k, v in over.items():
        setattr(c, k, v)
    return c


# ----------------------------------------------------------------------------------------------------
# small operators
# ----------------------------------------------------------------------------------------------------
def linear(p, prefix, x):
    return F.linear(x, p[prefix + ".weight"], p[prefix + ".bias"])


def mlp(p, prefix, x, num_layers, act_last):
    """upstream common/nets/layer.py:192-201 (MLP.forward)."""
    for i in range(num_layers - 1):
        x = F.relu(linear(p, "%s.layers.%d" % (prefix, i), x))
    x = linear(p, "%s.layers.%d" % (prefix, num_layers - 1), x)
    return F.relu(x) if act_last else x


def fold_weight_norm(g, v):
    """nn.utils.weight_norm(dim=0): W = g * v / ||v||_2 per output row (upstream sdf_net.py:57-62)."""
    return v * (g / v.norm(dim=1, keepdim=True))


def sdf_decoder(p, prefix, x):
    """upstream common/nets/sdf_net.py:87-122 in eval mode (dropout off), latent_in=[2], no classifier.

    289 -> 512 -> 223 (cat input 289 -> 512) -> 512 -> 512 -> 1, ReLU after layers 0..3, tanh at the end.
    """
    inp = x
    for layer in range(5):
        if layer == 2:
            x = torch.cat([x, inp], 1)
        pre = "%s.linh%d" % (prefix, layer)
        if layer < 4:
            w = fold_weight_norm(p[pre + ".weight_g"], p[pre + ".weight_v"])
        else:
            w = p[pre + ".weight"]
        x = F.linear(x, w, p[pre + ".bias"])
        if layer < 4:
            x = F.relu(x)
    return torch.tanh(x)[:, 0].unsqueeze(1)


def nerf_embed(x, octaves=5):
    """upstream common/utils/sdf_utils.py:96-141: cat_k [sin(x*2^k), cos(x*2^k)], k=0..4, no identity."""
    out = []
    for k in range(octaves):
        f = float(2.0 ** k)
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, -1)


def lattice(bins_n=64):
    """upstream main/model.py:257-273, reproduced op for op.

    `LongTensor / int` is TRUE division, so columns 0 and 1 carry a fractional shear
    (SURVEY.md section 7 'Bit-exact index masks').
    """
    voxel_origin = [-1, -1, -1]
    voxel_size = 2.0 / (bins_n - 1)
    idx = torch.arange(0, bins_n ** 3, 1, out=torch.LongTensor())
    s = torch.zeros(bins_n ** 3, 3)
    s[:, 2] = idx % bins_n
    s[:, 1] = (idx.long() / bins_n) % bins_n
    s[:, 0] = ((idx.long() / bins_n) / bins_n) % bins_n
    s[:, 0] = (s[:, 0] * voxel_size) + voxel_origin[2]
    s[:, 1] = (s[:, 1] * voxel_size) + voxel_origin[1]
    s[:, 2] = (s[:, 2] * voxel_size) + voxel_origin[0]
    return s


def grid_from_uv(uv, cfg):
    """upstream main/model.py:152-157 / :194-198 / :304-310."""
    normalizer = (torch.tensor([cfg.input_img_shape[1] - 1, cfg.input_img_shape[0] - 1]) / 2).to(uv.device)
    return (uv - normalizer) / normalizer


def gather_pyramid(pyramid, grids, sample=None):
    """5x F.grid_sample(bilinear, border, align_corners=True) + cat + permute (model.py:164-175, 316-328).

    pyramid: dict level -> (B,C_l,H_l,W_l); grids: (b,N,2); `sample` selects one batch item (sdf_infer).
    returns (b, N, C)
    """
    g = grids.unsqueeze(1)
    feats = []
    for name in LEVELS:
        fmap = pyramid[name]
        if sample is not None:
            fmap = fmap[sample].unsqueeze(0)
        feats.append(F.grid_sample(fmap, g, padding_mode="border", align_corners=True))
    return torch.cat(feats, dim=1).squeeze(2).permute(0, 2, 1)


def project(points_cam, K):
    """upstream main/model.py:149-150 (bmm with K^T, divide by z)."""
    uvw = torch.bmm(points_cam, K.transpose(1, 2))
    return uvw[:, :, :2] / uvw[:, :, [2]]


# ----------------------------------------------------------------------------------------------------
# stage (a1) sdf_infer, (a2) sdf_forward, (a3) get_input_transformer
# ----------------------------------------------------------------------------------------------------
def candidate_mask(samples, center, K, bbox, sdf_scale):
    """upstream main/model.py:286-300 for ONE sample: project the lattice, strict bbox test.

    Returns (mask (bins^3,) bool, uv (bins^3, 2)).
    """
    cam = (samples.clone() / sdf_scale) + center.unsqueeze(0)
    uvw = torch.mm(cam, K.transpose(0, 1))
    uv = uvw[:, :2] / uvw[:, [2]]
    m = torch.logical_and(
        torch.logical_and(uv[:, 0] > bbox[0], uv[:, 0] < bbox[2]),
        torch.logical_and(uv[:, 1] > bbox[1], uv[:, 1] < bbox[3]),
    )
    return m, uv


def sdf_infer(p, pyramid, center, K, bbox, sdf_scale, num_points, kind, cfg, taps=None):
    """upstream main/model.py:246-355.  Returns (points, sdf, posenc, None).

    `taps` (optional dict) receives per-sample diagnostics used by the parity tests:
    `index` (B,P) int64 selected lattice indices in selection order, `n_f` (B,) candidate counts,
    `cand_index` / `cand_sdf` lists (per sample: lattice indices that passed the bbox, raw SDF values).
    """
    B = center.shape[0]
    dev = center.device               # cuda only in bench.py's "stock PyTorch eager on the same GPU" baseline leg
    samples = lattice(cfg.bins_n)
    all_idx = torch.arange(cfg.bins_n ** 3)
    pts = torch.zeros(B, num_points, 3, device=dev)
    sdf = torch.zeros(B, num_points, 1, device=dev)
    pe = torch.zeros(B, num_points, cfg.PointFeatSize - 3, device=dev)
    sel_index = torch.zeros(B, num_points, dtype=torch.long, device=dev)
    n_f = torch.zeros(B, dtype=torch.long)
    cand_index, cand_sdf = [], []
    dec = "%s_sdf_decoder" % kind
    for b in range(B):
        # upstream does the projection and the bbox test on the CPU whatever the model's device (model.py:286-302:
        # three .cpu() copies, then the surviving points go back with .cuda())
        m, uv = candidate_mask(samples, center[b].cpu(), K[b].cpu(), bbox[b].cpu(), sdf_scale)
        b_uv = uv[m].unsqueeze(0).to(dev)
        b_samples = samples[m].clone().to(dev)
        b_index = all_idx[m].to(dev)
        feats = gather_pyramid(pyramid, grid_from_uv(b_uv, cfg), sample=b)
        fea = mlp(p, "linear_sdfin", feats, 2, True)
        b_pe = nerf_embed(b_samples, (cfg.PointFeatSize - 3) // 6)
        dec_in = torch.cat([fea.squeeze(0), b_pe, b_samples], 1).contiguous()
        b_sdf = sdf_decoder(p, dec, dec_in).squeeze(1)
        _, order = torch.sort(b_sdf.abs())
        order = order[:num_points]
        pts[b] = b_samples[order]           # raises (shape mismatch) if N_f < num_points, like upstream :348
        sdf[b] = b_sdf[order].unsqueeze(-1)
        pe[b] = b_pe[order]
        sel_index[b] = b_index[order]
        n_f[b] = int(m.sum())
        cand_index.append(b_index)
        cand_sdf.append(b_sdf)
    sdf = torch.clamp(sdf, -cfg.ClampingDistance, cfg.ClampingDistance)
    if taps is not None:
        taps.update(index=sel_index, n_f=n_f, cand_index=cand_index, cand_sdf=cand_sdf)
    return pts, sdf, pe, None


def sdf_forward(p, pyramid, sdf_points, center, K, sdf_scale, kind, cfg):
    """upstream main/model.py:181-244.  Returns (sdf (B,P,1), None, posenc (B,P,30))."""
    B, P, _ = sdf_points.shape
    cam = (sdf_points / sdf_scale) + center[:, None, :]
    uv = project(cam, K)
    feats = gather_pyramid(pyramid, grid_from_uv(uv, cfg)).contiguous()
    fea = mlp(p, "linear_sdfin", feats, 2, True)
    pe = nerf_embed(sdf_points.reshape(-1, 3), (cfg.PointFeatSize - 3) // 6)
    dec_in = torch.cat([fea.reshape(-1, fea.shape[-1]), pe, sdf_points.reshape(-1, 3)], 1).contiguous()
    sdf = sdf_decoder(p, "%s_sdf_decoder" % kind, dec_in).reshape(B, P, 1)
    sdf = torch.clamp(sdf, -cfg.ClampingDistance, cfg.ClampingDistance)
    return sdf, None, pe.reshape(B, P, -1)


def get_input_transformer(p, pyramid, sdf_points, center, K, sdf_scale, cfg):
    """upstream main/model.py:145-179.  Returns (latent (B,P,223), cam_points (B,P,3))."""
    cam = (sdf_points / sdf_scale) + center[:, None, :]
    uv = project(cam, K)
    feats = gather_pyramid(pyramid, grid_from_uv(uv, cfg)).contiguous()
    return mlp(p, "linear_transformerin", feats, 4, True), cam


def sdf_activation(p, name, sdf):
    """upstream main/model.py:123-126 (in-place floor of beta at 2e-3, then sigmoid(sdf/beta)/beta)."""
    beta = p[name]
    # upstream clamps `beta.data` (no tape, no version bump), so a leaf that requires grad stays usable under autograd
    beta.data.copy_(torch.maximum(torch.zeros_like(beta) + 2e-3, beta.data))
    return torch.sigmoid(sdf / beta) / beta


# ----------------------------------------------------------------------------------------------------
# stage (a10) masks, (a11) transformers
# ----------------------------------------------------------------------------------------------------
def mano_tgt_mask(cfg):
    """upstream common/utils/misc.py:11-31: block-diagonal groups {0},{1-3},...,{13-15},{16}; True = blocked."""
    n = cfg.mano_num_queries
    m = torch.zeros(n, n, dtype=torch.bool)
    m[0, :] = True
    m[0, 0] = False
    for i in range(5):
        s, e = 3 * i + 1, 3 * i + 4
        m[s:e, :] = True
        m[s:e, s:e] = False
    m[cfg.mano_shape_indx, :] = True
    m[cfg.mano_shape_indx, cfg.mano_shape_indx] = False
    return m


def mano_memory_mask(cfg):
    """upstream common/utils/misc.py:42-47: queries may not attend to the object-side memory tokens."""
    m = torch.zeros(cfg.mano_num_queries, cfg.num_samp_hand + cfg.num_samp_obj, dtype=torch.bool)
    m[:, cfg.num_samp_hand:] = True
    return m


def multihead_attention(p, prefix, query, key, value, nhead, attn_mask=None):
    """nn.MultiheadAttention forward as used upstream (transformer.py:294,378,383); (L,B,d) layout.

    in_proj rows are [q;k;v]; heads are contiguous 64-wide slices; q is scaled by 1/sqrt(head_dim);
    bool mask True -> -inf before the softmax; dropout off (eval).
    """
    L, B, d = query.shape
    S = key.shape[0]
    hd = d // nhead
    W, bias = p[prefix + ".in_proj_weight"], p[prefix + ".in_proj_bias"]
    q = F.linear(query, W[:d], bias[:d])
    k = F.linear(key, W[d:2 * d], bias[d:2 * d])
    v = F.linear(value, W[2 * d:], bias[2 * d:])
    q = q.reshape(L, B * nhead, hd).transpose(0, 1) * (1.0 / math.sqrt(hd))
    k = k.reshape(S, B * nhead, hd).transpose(0, 1)
    v = v.reshape(S, B * nhead, hd).transpose(0, 1)
    scores = torch.bmm(q, k.transpose(1, 2))
    if attn_mask is not None:
        scores = scores.masked_fill(attn_mask.unsqueeze(0), float("-inf"))
    attn = torch.softmax(scores, dim=-1)
    out = torch.bmm(attn, v).transpose(0, 1).reshape(L, B, d)
    return F.linear(out, p[prefix + ".out_proj.weight"], p[prefix + ".out_proj.bias"])


def layer_norm(p, prefix, x):
    return F.layer_norm(x, (x.shape[-1],), p[prefix + ".weight"], p[prefix + ".bias"], 1e-5)


def encoder_layer(p, prefix, src, pos, nhead):
    """upstream common/nets/transformer.py:279-302 (forward_post)."""
    qk = src + pos
    src2 = multihead_attention(p, prefix + ".self_attn", qk, qk, src, nhead)
    src = layer_norm(p, prefix + ".norm1", src + src2)
    src2 = linear(p, prefix + ".linear2", F.relu(linear(p, prefix + ".linear1", src)))
    return layer_norm(p, prefix + ".norm2", src + src2)


def encoder(p, prefix, src, pos, num_layers, nhead):
    """upstream common/nets/transformer.py:175-202: returns (last, stack(inter_norm(out_l)))."""
    out = src
    inter = []
    for i in range(num_layers):
        out = encoder_layer(p, "%s.layers.%d" % (prefix, i), out, pos, nhead)
        inter.append(layer_norm(p, prefix + ".inter_norm", out))
    return out, torch.stack(inter)


def decoder_layer(p, prefix, tgt, memory, pos, query_pos, tgt_mask, memory_mask, nhead):
    """upstream common/nets/transformer.py:366-395 (forward_post)."""
    qk = tgt + query_pos
    tgt2 = multihead_attention(p, prefix + ".self_attn", qk, qk, tgt, nhead, tgt_mask)
    tgt = layer_norm(p, prefix + ".norm1", tgt + tgt2)
    tgt2 = multihead_attention(p, prefix + ".multihead_attn", tgt + query_pos, memory + pos, memory,
                               nhead, memory_mask)
    tgt = layer_norm(p, prefix + ".norm2", tgt + tgt2)
    tgt2 = linear(p, prefix + ".linear2", F.relu(linear(p, prefix + ".linear1", tgt)))
    return layer_norm(p, prefix + ".norm3", tgt + tgt2)


def decoder(p, prefix, tgt, memory, pos, query_pos, tgt_mask, memory_mask, num_layers, nhead):
    """upstream common/nets/transformer.py:214-252 with return_intermediate=True."""
    out = tgt
    inter = []
    for i in range(num_layers):
        out = decoder_layer(p, "%s.layers.%d" % (prefix, i), out, memory, pos, query_pos,
                            tgt_mask, memory_mask, nhead)
        inter.append(layer_norm(p, prefix + ".norm", out))
    return torch.stack(inter)


def transformer(p, prefix, src, query_embed, pos_embed, tgt_mask, memory_mask, cfg):
    """upstream common/nets/transformer.py:119-155.  Returns (hs, memory, intermediate)."""
    S, B, d = src.shape
    qe = query_embed.unsqueeze(1).repeat(1, B, 1)
    tgt = torch.zeros_like(qe)
    memory, inter = encoder(p, prefix + ".encoder", src + pos_embed, pos_embed, cfg.enc_layers, cfg.nheads)
    hs = decoder(p, prefix + ".decoder", tgt, memory, pos_embed, qe, tgt_mask, memory_mask,
                 cfg.dec_layers, cfg.nheads)
    return hs, memory, inter


def vote_transformer(p, prefix, src, pos_embed, cfg):
    """upstream common/nets/transformer.py:53-65 (3 encoder layers)."""
    return encoder(p, prefix + ".encoder", src + pos_embed, pos_embed, cfg.enc_layers // 2, cfg.nheads)


# ----------------------------------------------------------------------------------------------------
# stage (a13) ManoHead + ManoLayer, (a14) vote aggregation
# ----------------------------------------------------------------------------------------------------
def rot6d_to_mat(x):
    """upstream common/nets/mano_head.py:185-194."""
    a1, a2 = x[:, 0:3], x[:, 3:6]
    b1 = F.normalize(a1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1)
    b3 = torch.cross(b1, b2, dim=1)
    return torch.stack((b1, b2, b3), dim=-1)


def mat_to_quat(R34, eps=1e-6):
    """upstream common/nets/mano_head.py:90-182 (torchgeometry-style branchy conversion)."""
    r = R34.transpose(1, 2)
    mask_d2 = r[:, 2, 2] < eps
    mask_d0_d1 = r[:, 0, 0] > r[:, 1, 1]
    mask_d0_nd1 = r[:, 0, 0] < -r[:, 1, 1]
    t0 = 1 + r[:, 0, 0] - r[:, 1, 1] - r[:, 2, 2]
    q0 = torch.stack([r[:, 1, 2] - r[:, 2, 1], t0, r[:, 0, 1] + r[:, 1, 0], r[:, 2, 0] + r[:, 0, 2]], -1)
    t1 = 1 - r[:, 0, 0] + r[:, 1, 1] - r[:, 2, 2]
    q1 = torch.stack([r[:, 2, 0] - r[:, 0, 2], r[:, 0, 1] + r[:, 1, 0], t1, r[:, 1, 2] + r[:, 2, 1]], -1)
    t2 = 1 - r[:, 0, 0] - r[:, 1, 1] + r[:, 2, 2]
    q2 = torch.stack([r[:, 0, 1] - r[:, 1, 0], r[:, 2, 0] + r[:, 0, 2], r[:, 1, 2] + r[:, 2, 1], t2], -1)
    t3 = 1 + r[:, 0, 0] + r[:, 1, 1] + r[:, 2, 2]
    q3 = torch.stack([t3, r[:, 1, 2] - r[:, 2, 1], r[:, 2, 0] - r[:, 0, 2], r[:, 0, 1] - r[:, 1, 0]], -1)
    c0 = (mask_d2 & mask_d0_d1).view(-1, 1).float()
    c1 = (mask_d2 & ~mask_d0_d1).view(-1, 1).float()
    c2 = (~mask_d2 & mask_d0_nd1).view(-1, 1).float()
    c3 = (~mask_d2 & ~mask_d0_nd1).view(-1, 1).float()
    q = q0 * c0 + q1 * c1 + q2 * c2 + q3 * c3
    q = q / torch.sqrt(t0.view(-1, 1) * c0 + t1.view(-1, 1) * c1 + t2.view(-1, 1) * c2 + t3.view(-1, 1) * c3)
    return q * 0.5


def quat_to_aa(q):
    """upstream common/nets/mano_head.py:54-87."""
    q1, q2, q3 = q[..., 1], q[..., 2], q[..., 3]
    s2 = q1 * q1 + q2 * q2 + q3 * q3
    s = torch.sqrt(s2)
    c = q[..., 0]
    two_theta = 2.0 * torch.where(c < 0.0, torch.atan2(-s, -c), torch.atan2(s, c))
    k = torch.where(s2 > 0.0, two_theta / s, 2.0 * torch.ones_like(s))
    return torch.stack([q1 * k, q2 * k, q3 * k], -1)


def mat_to_aa(R):
    """upstream common/nets/mano_head.py:197-217 (pad to 3x4, quaternion, axis-angle, NaN -> 0)."""
    aa = quat_to_aa(mat_to_quat(F.pad(R, (0, 1), "constant", 1.0)))
    aa[torch.isnan(aa)] = 0.0
    return aa


def rodrigues(aa):
    """upstream manopth/manopth/rodrigues_layer.py:16-56 (via quaternion, +1e-8 inside the norm)."""
    angle = torch.norm(aa + 1e-8, p=2, dim=1).unsqueeze(-1)
    n = aa / angle
    half = angle * 0.5
    quat = torch.cat([torch.cos(half), torch.sin(half) * n], dim=1)
    quat = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = quat[:, 0], quat[:, 1], quat[:, 2], quat[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([
        w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1)


def _with_zeros(t):
    """manopth/manopth/tensutils.py:16-23: append the [0,0,0,1] row."""
    pad = t.new_tensor([0.0, 0.0, 0.0, 1.0]).view(1, 1, 4).repeat(t.shape[0], 1, 1)
    return torch.cat([t, pad], 1)


def mano_layer(p, prefix, pose_aa, betas):
    """upstream manopth/manopth/manolayer.py:111-276 with use_pca=False, axisang root, flat_hand_mean,
    center_idx=0, side='right', th_trans = 0.  pose_aa (N,48), betas (N,10) -> verts (N,778,3), joints (N,21,3) [mm].
    """
    N = pose_aa.shape[0]
    shapedirs, posedirs = p[prefix + ".th_shapedirs"], p[prefix + ".th_posedirs"]
    v_template, J_reg = p[prefix + ".th_v_template"], p[prefix + ".th_J_regressor"]
    weights, hands_mean = p[prefix + ".th_weights"], p[prefix + ".th_hands_mean"]
    full_pose = torch.cat([pose_aa[:, :3], hands_mean + pose_aa[:, 3:48]], 1)
    rot_map = rodrigues(full_pose.contiguous().view(-1, 3)).view(N, 16 * 9)
    eye = torch.eye(3, device=pose_aa.device).view(1, 9).repeat(N, 16)
    pose_map = (rot_map - eye)[:, 9:]
    root_rot = rot_map[:, :9].view(N, 3, 3)
    rot_map = rot_map[:, 9:]
    v_shaped = torch.matmul(shapedirs, betas.transpose(1, 0)).permute(2, 0, 1) + v_template
    J = torch.matmul(J_reg, v_shaped)
    v_posed = v_shaped + torch.matmul(posedirs, pose_map.transpose(0, 1)).permute(2, 0, 1)
    root_j = J[:, 0, :].contiguous().view(N, 3, 1)
    root_trans = _with_zeros(torch.cat([root_rot, root_j], 2))
    all_rots = rot_map.view(N, 15, 3, 3)
    l1, l2, l3 = [1, 4, 7, 10, 13], [2, 5, 8, 11, 14], [3, 6, 9, 12, 15]
    r1, r2, r3 = (all_rots[:, [i - 1 for i in l]] for l in (l1, l2, l3))
    j1, j2, j3 = J[:, l1], J[:, l2], J[:, l3]
    transforms = [root_trans.unsqueeze(1)]
    rel1 = _with_zeros(torch.cat([r1, (j1 - root_j.transpose(1, 2)).unsqueeze(3)], 3).view(-1, 3, 4))
    root_flt = root_trans.unsqueeze(1).repeat(1, 5, 1, 1).view(N * 5, 4, 4)
    t1 = torch.matmul(root_flt, rel1)
    transforms.append(t1.view(N, 5, 4, 4))
    rel2 = _with_zeros(torch.cat([r2, (j2 - j1).unsqueeze(3)], 3).view(-1, 3, 4))
    t2 = torch.matmul(t1, rel2)
    transforms.append(t2.view(N, 5, 4, 4))
    rel3 = _with_zeros(torch.cat([r3, (j3 - j2).unsqueeze(3)], 3).view(-1, 3, 4))
    t3 = torch.matmul(t2, rel3)
    transforms.append(t3.view(N, 5, 4, 4))
    reorder = [0, 1, 6, 11, 2, 7, 12, 3, 8, 13, 4, 9, 14, 5, 10, 15]
    results = torch.cat(transforms, 1)[:, reorder]
    joint_js = torch.cat([J, J.new_zeros(N, 16, 1)], 2)
    tmp2 = torch.matmul(results, joint_js.unsqueeze(3))
    results2 = (results - torch.cat([tmp2.new_zeros(N, 16, 4, 3), tmp2], 3)).permute(0, 2, 3, 1)
    T = torch.matmul(results2, weights.transpose(0, 1))
    rest_h = torch.cat([v_posed.transpose(2, 1), torch.ones(N, 1, v_posed.shape[1], device=v_posed.device)], 1)
    verts = (T * rest_h.unsqueeze(1)).sum(2).transpose(2, 1)[:, :, :3]
    jtr = results[:, :, :3, 3]
    tips = verts[:, [745, 317, 444, 556, 673]]
    jtr = torch.cat([jtr, tips], 1)
    jtr = jtr[:, [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]]
    center = jtr[:, 0].unsqueeze(1)
    return (verts - center) * 1000, (jtr - center) * 1000


def mano_head(p, pose6d, shape):
    """upstream common/nets/mano_head.py:232-256 (prediction branch).  pose6d (L,16,B,6), shape (L,B,10)."""
    L, N, B, C = pose6d.shape
    R = rot6d_to_mat(pose6d.permute(0, 2, 1, 3).reshape(L * B * N, C).contiguous()).contiguous()
    pose = mat_to_aa(R).contiguous().view(-1, 48)
    verts, joints = mano_layer(p, "mano_head.mano_layer", pose, shape.reshape(-1, 10))
    return verts.view(L, B, 778, 3) / 1000, joints.view(L, B, 21, 3) / 1000


def mano_head_gt(p, mano_params):
    """upstream common/nets/mano_head.py:258-276 (ground-truth branch: training and the dexycb evaluation).
    mano_params (B,58) = axis-angle pose (48) | shape (10).  The pose slice is copied (`.contiguous()` of a
    column slice, :260), so the caller's tensor is NOT modified."""
    gt_shape = mano_params[:, 48:]
    gt_pose = mano_params[:, :48].contiguous()
    gt_pose[:, 3:] = gt_pose[:, 3:] - p["mano_head.mano_layer.th_hands_mean"]
    rotmat = rodrigues(gt_pose.view(-1, 3)).view(-1, 16, 3, 3)      # batch_rodrigues, mano_head.py:12-51
    verts, joints = mano_layer(p, "mano_head.mano_layer", gt_pose, gt_shape)
    return {"verts3d": verts / 1000, "joints3d": joints / 1000, "mano_shape": gt_shape, "mano_pose": rotmat}


def vote_joints(hand_points, hand_off, hand_cls):
    """upstream common/nets/loss.py:31-36,54-57: softmax over points of the class logits, weighted vote sum.

    hand_points (B,P,3), hand_off (L,P,B,60), hand_cls (L,P,B,20) -> hand_joints (L,B,20,3)
    """
    l, pn, b, j = hand_cls.shape
    vote = hand_points.unsqueeze(2).unsqueeze(0) + hand_off.reshape(l, pn, b, j, 3).permute(0, 2, 1, 3, 4)
    w = torch.softmax(hand_cls, dim=1).permute(0, 2, 1, 3).unsqueeze(-1)
    return torch.sum(vote * w, dim=2)


# ----------------------------------------------------------------------------------------------------
# ResNet-50 + U-Net (the step BEFORE the hot path; needed for the full-forward CPU baseline)
# ----------------------------------------------------------------------------------------------------
def _bn(p, prefix, x, training=False):
    """nn.BatchNorm2d: running statistics in eval mode; in training mode the batch statistics (upstream trains with
    model.train(), main/train.py:89 -- also for the backbone, whose BN affine parameters alone are frozen,
    main/model.py:117-121).  The running-statistics update of training mode does not touch outputs or gradients and is
    left out."""
    if training:
        return F.batch_norm(x, None, None, p[prefix + ".weight"], p[prefix + ".bias"], True, 0.1, 1e-5)
    return F.batch_norm(x, p[prefix + ".running_mean"], p[prefix + ".running_var"],
                        p[prefix + ".weight"], p[prefix + ".bias"], False, 0.0, 1e-5)


def _bottleneck(p, prefix, x, stride, training=False):
    """torchvision Bottleneck (v1.5: stride on the 3x3) as instantiated by upstream resnet.py:19,49-68."""
    out = F.relu(_bn(p, prefix + ".bn1", F.conv2d(x, p[prefix + ".conv1.weight"]), training))
    out = F.relu(_bn(p, prefix + ".bn2", F.conv2d(out, p[prefix + ".conv2.weight"], stride=stride, padding=1), training))
    out = _bn(p, prefix + ".bn3", F.conv2d(out, p[prefix + ".conv3.weight"]), training)
    if prefix + ".downsample.0.weight" in p:
        x = _bn(p, prefix + ".downsample.1", F.conv2d(x, p[prefix + ".downsample.0.weight"], stride=stride), training)
    return F.relu(out + x)


def backbone(p, img, prefix="backbone_net.resnet", training=False):
    """upstream common/nets/resnet.py:70-87."""
    skips = {}
    x = F.relu(_bn(p, prefix + ".bn1", F.conv2d(img, p[prefix + ".conv1.weight"], stride=2, padding=3), training))
    skips["stride2"] = x
    x = F.max_pool2d(x, 3, 2, 1)
    for li, (blocks, name) in enumerate(((3, "stride4"), (4, "stride8"), (6, "stride16"), (3, "stride32")), 1):
        for bi in range(blocks):
            x = _bottleneck(p, "%s.layer%d.%d" % (prefix, li, bi), x, 2 if (bi == 0 and li > 1) else 1, training)
        skips[name] = x
    return x, skips


def _conv_stack(p, prefix, x, n, k, final_bn=True, training=False):
    idx = 0
    for i in range(n):
        x = F.conv2d(x, p["%s.%d.weight" % (prefix, idx)], p["%s.%d.bias" % (prefix, idx)], padding=k // 2)
        idx += 1
        if i < n - 1 or final_bn:
            x = F.relu(_bn(p, "%s.%d" % (prefix, idx), x, training))
            idx += 2
    return x


def _deconv(p, prefix, x, training=False):
    x = F.conv_transpose2d(x, p[prefix + ".0.weight"], stride=2, padding=1)
    return F.relu(_bn(p, prefix + ".1", x, training))


def unet_decoder(p, feat, skips, arch, prefix="decoder_net.resnet_decoder", training=False):
    """upstream common/nets/module.py:172-218 (Decoder_big, 'ho3d') / :98-144 (Decoder, resnet50)."""
    pyr = {}
    big = arch == "ho3d"
    t = training
    pyr["stride32"] = feat if big else _conv_stack(p, prefix + ".conv0d", feat, 1, 1, training=t)
    x = feat
    for i, name in ((1, "stride16"), (2, "stride8"), (3, "stride4"), (4, "stride2")):
        skip = skips[name] if big else _conv_stack(p, "%s.conv%dd" % (prefix, i), skips[name], 1, 1, training=t)
        up = _deconv(p, "%s.deconv%d" % (prefix, i), x, t)
        x = _conv_stack(p, "%s.conv%d" % (prefix, i), torch.cat((skip, up), 1), 1, 3, training=t)
        pyr[name] = x
    n_out = 3 if big else 2
    hm = _conv_stack(p, prefix + ".convOut_hm", x, n_out, 1, final_bn=False, training=t)
    hs = _conv_stack(p, prefix + ".convOut_hand_seg", x, n_out, 1, final_bn=False, training=t).sigmoid()
    os_ = _conv_stack(p, prefix + ".convOut_obj_seg", x, n_out, 1, final_bn=False, training=t).sigmoid()
    return pyr, torch.cat([hm, hs, os_], dim=1)


# ----------------------------------------------------------------------------------------------------
# (a15) the eval forward from the pyramid on (upstream main/model.py:424-638), and from the image
# ----------------------------------------------------------------------------------------------------
def hot_path_eval(p, pyramid, meta, cfg, taps=None):
    """Everything of upstream Model.forward(mode='eval', dataset='ho3d') after the U-Net decoder.

    Returns the `*_out` dict (model.py:616-638).  `taps` collects intermediate tensors for stage parity.
    """
    root, objc, K = meta["mano_root"], meta["obj_center_cam"], meta["cam_intr"]
    th, to = {}, {}
    hand_points, hand_sdf, hand_pe, _ = sdf_infer(p, pyramid, root, K, meta["bbox_hand"], cfg.hand_sdf_scale,
                                                  cfg.num_samp_hand, "hand", cfg, th)
    obj_points, obj_sdf, obj_pe, _ = sdf_infer(p, pyramid, objc, K, meta["bbox_obj"], cfg.obj_sdf_scale,
                                               cfg.num_samp_obj, "obj", cfg, to)
    if taps is not None:
        taps.update(hand=th, obj=to)
    return pose_from_points(p, pyramid, meta, cfg, hand_points, hand_sdf, hand_pe, obj_points, obj_sdf, obj_pe, taps)


def pose_from_points(p, pyramid, meta, cfg, hand_points, hand_sdf, hand_pe, obj_points, obj_sdf, obj_pe, taps=None):
    """upstream main/model.py:483-638: from the points the pose branch works on (selected by sdf_infer in eval mode,
    jittered `*_pre_points` in the first training epochs) to the `*_out` dict.  The `.detach()` calls are upstream's
    (:483-484,518-519,536,555): they only matter under autograd (`model_train`)."""
    root, objc, K = meta["mano_root"], meta["obj_center_cam"], meta["cam_intr"]
    sigma_hand = sdf_activation(p, "hand_sigmoid_beta", hand_sdf.detach())
    sigma_obj = sdf_activation(p, "obj_sigmoid_beta", obj_sdf.detach())
    hand_fea, hand_cam = get_input_transformer(p, pyramid, hand_points, root, K, cfg.hand_sdf_scale, cfg)
    hand_nt = hand_cam - root[:, None, :]
    obj_fea, obj_cam = get_input_transformer(p, pyramid, obj_points, objc, K, cfg.obj_sdf_scale, cfg)
    obj_nt = obj_cam - objc[:, None, :]
    hand_o_points = (hand_cam - objc[:, None, :]) * cfg.obj_sdf_scale
    hand_o_nt = hand_cam - objc[:, None, :]                       # upstream model.py:498 ("bug", kept)
    hand_o_sdf, _, hand_o_pe = sdf_forward(p, pyramid, hand_o_points, objc, K, cfg.obj_sdf_scale, "obj", cfg)
    obj_h_points = (obj_cam - root[:, None, :]) * cfg.hand_sdf_scale
    obj_h_nt = obj_cam - root[:, None, :]                         # upstream model.py:508 ("bug", kept)
    obj_h_sdf, _, obj_h_pe = sdf_forward(p, pyramid, obj_h_points, root, K, cfg.hand_sdf_scale, "hand", cfg)
    sigma_hand_o = sdf_activation(p, "obj_sigmoid_beta", hand_o_sdf.detach())
    sigma_obj_h = sdf_activation(p, "hand_sigmoid_beta", obj_h_sdf.detach())

    def tok(nt, pe, fea):
        return torch.cat([nt, pe, fea], dim=2).permute(1, 0, 2).contiguous()

    hand_in = torch.cat([tok(hand_nt, hand_pe, hand_fea * sigma_hand),
                         tok(obj_h_nt, obj_h_pe, obj_fea * sigma_obj_h).detach()], dim=0)
    obj_in = torch.cat([tok(obj_nt, obj_pe, obj_fea * sigma_obj),
                        tok(hand_o_nt, hand_o_pe, hand_fea * sigma_hand_o).detach()], dim=0)
    hs, memory, hand_enc = transformer(p, "hand_transformer", hand_in, p["mano_query_embed.weight"],
                                       torch.zeros_like(hand_in), mano_tgt_mask(cfg).to(hand_in.device),
                                       mano_memory_mask(cfg).to(hand_in.device), cfg)
    _, obj_enc = vote_transformer(p, "obj_transformer", obj_in, torch.zeros_like(obj_in), cfg)
    Ph, Po = cfg.num_samp_hand, cfg.num_samp_obj
    hand_off = mlp(p, "linear_handvote", hand_enc[:, :Ph], 4, False)
    hand_cls = mlp(p, "linear_handcls", hand_enc[:, :Ph], 3, False)
    obj_rot = mlp(p, "linear_obj_rot", obj_enc[:, :Po], 3, False)
    obj_trans = mlp(p, "linear_obj_rel_trans", obj_enc[:, :Po], 3, False)
    pose6d = mlp(p, "linear_pose", hs[:, :cfg.mano_shape_indx], 3, False)
    shape = mlp(p, "linear_shape", hs[:, cfg.mano_shape_indx], 3, False)
    verts, joints = mano_head(p, pose6d, shape)
    hand_joints = vote_joints(hand_nt, hand_off, hand_cls)
    if taps is not None:
        taps["hand_points_notrans"] = hand_nt
    out = {
        "mano_mesh_out": verts[-1],
        "mano_joints_out": joints[-1],
        "obj_rot_out": obj_rot[-1].permute(1, 0, 2).contiguous(),
        "obj_trans_out": obj_trans[-1].permute(1, 0, 2).contiguous(),
        "hand_joints_out": hand_joints[-1],
    }
    if taps is not None:
        taps.update(
            hand_points=hand_points, hand_sdf=hand_sdf, hand_posenc=hand_pe,
            obj_points=obj_points, obj_sdf=obj_sdf, obj_posenc=obj_pe, hand_fea=hand_fea, obj_fea=obj_fea,
            hand_o_sdf=hand_o_sdf, obj_h_sdf=obj_h_sdf, hand_transformer_in=hand_in,
            obj_transformer_in=obj_in, hs=hs, memory=memory, hand_encoder_out=hand_enc,
            obj_encoder_out=obj_enc, hand_off=hand_off, hand_cls=hand_cls, obj_rot=obj_rot,
            obj_trans=obj_trans, mano_pose6d=pose6d, mano_shape=shape, hand_joints=hand_joints,
            mano_verts=verts, mano_joints=joints)
    return out


def model_eval(p, img, meta, cfg, arch="ho3d", taps=None):
    """upstream Model.forward(mode='eval') from the image (model.py:357-368 then the hot path)."""
    feat, skips = backbone(p, img)
    pyramid, decoder_out = unet_decoder(p, feat, skips, arch)
    if taps is not None:
        taps["pyramid"] = pyramid
        taps["decoder_out"] = decoder_out
    return hot_path_eval(p, pyramid, meta, cfg, taps)


# ----------------------------------------------------------------------------------------------------
# the dexycb evaluation branch (upstream main/model.py:370-422, 606-654) and the eval-mode loss entries
# ----------------------------------------------------------------------------------------------------
def render_gaussian_heatmap(joint_coord, cfg):
    """upstream main/model.py:128-143."""
    x = torch.arange(cfg.output_hm_shape[2], device=joint_coord.device)
    y = torch.arange(cfg.output_hm_shape[1], device=joint_coord.device)
    yy, xx = torch.meshgrid(y, x, indexing="ij")
    xx, yy = xx[None, None].float(), yy[None, None].float()
    x = joint_coord[:, :, 0, None, None]
    y = joint_coord[:, :, 1, None, None]
    heatmap = torch.exp(-(((xx - x) / cfg.sigma) ** 2) / 2 - (((yy - y) / cfg.sigma) ** 2) / 2)
    return torch.sum(heatmap, 1) * 255


def joint_vote_losses(hand_points, hand_off, hand_cls, hand_joints, joint_gt, cfg):
    """upstream common/nets/loss.py:31-61 (the three loss values; `hand_joints` comes from vote_joints)."""
    l, pn, b, j = hand_cls.shape
    vote = hand_points.unsqueeze(2).unsqueeze(0) + hand_off.reshape(l, pn, b, j, 3).permute(0, 2, 1, 3, 4)
    cls_gt = (torch.norm(hand_points.unsqueeze(2) - joint_gt.unsqueeze(1) / 1000, dim=-1) < cfg.hand_cls_dist).float()
    l3d = F.smooth_l1_loss(vote * 1000, joint_gt.unsqueeze(1).unsqueeze(0).expand(l, b, pn, j, 3), reduction="none")
    l3d = (l3d * cls_gt.unsqueeze(-1).unsqueeze(0).expand(l, b, pn, j, 3)).sum((1, 2, 3)) / cls_gt.sum()
    lcls = F.binary_cross_entropy_with_logits(hand_cls.permute(0, 2, 1, 3).contiguous(),
                                              cls_gt.unsqueeze(0).expand(l, b, pn, j))
    lall = F.smooth_l1_loss(hand_joints * 1000, joint_gt.unsqueeze(0).expand(l, b, j, 3))
    return l3d.mean(), lcls, lall


def mano_losses(pred, gt, cfg):
    """upstream common/nets/loss.py:99-153 with the lambdas of model.py:107-112."""
    exp = lambda k: gt[k].unsqueeze(0).expand(pred[k].shape)  # noqa: E731
    return {"mano_mesh_loss": cfg.lambda_verts3d * F.mse_loss(pred["verts3d"], exp("verts3d")),
            "mano_joint_loss": cfg.lambda_joints3d * F.mse_loss(pred["joints3d"], exp("joints3d")),
            "pose_param_loss": cfg.lambda_manopose * F.mse_loss(pred["mano_pose"], exp("mano_pose")),
            "shape_param_loss": cfg.lambda_manoshape * F.mse_loss(pred["mano_shape"], exp("mano_shape"))}


def model_eval_dexycb(p, img, inputs, targets, meta, cfg, arch="dexycb", taps=None, pyramid=None, decoder_out=None):
    """upstream Model.forward(mode='eval') with cfg.dataset == 'dexycb' (model.py:357-665): the ho3d eval path
    plus SDF supervision queries, heat-map / segmentation heads, the ground-truth MANO forward and every loss
    entry.  Returns the flat dict upstream returns ({**loss, **out})."""
    if pyramid is None:
        feat, skips = backbone(p, img)
        pyramid, decoder_out = unet_decoder(p, feat, skips, arch)
    taps = {} if taps is None else taps
    root, objc, K = meta["mano_root"], meta["obj_center_cam"], meta["cam_intr"]
    c = cfg.ClampingDistance
    loss = {}
    hand_s, _, _ = sdf_forward(p, pyramid, inputs["hand_sdf_points"], root, K, cfg.hand_sdf_scale, "hand", cfg)
    obj_s, _, _ = sdf_forward(p, pyramid, inputs["obj_sdf_points"], objc, K, cfg.obj_sdf_scale, "obj", cfg)
    loss["sdfhand_loss"] = F.l1_loss(hand_s, targets["hand_sdf"].clamp(-c, c).unsqueeze(-1))
    loss["sdfobj_loss"] = F.l1_loss(obj_s, targets["obj_sdf"].clamp(-c, c).unsqueeze(-1))
    out = {"joint_heatmap_out": decoder_out[:, 0], "hand_seg_gt_out": targets["hand_seg"],
           "hand_seg_pred_out": decoder_out[:, 1], "obj_seg_gt_out": targets["obj_seg"],
           "obj_seg_pred_out": decoder_out[:, 2]}
    loss["joint_heatmap"] = (decoder_out[:, 0] - render_gaussian_heatmap(targets["joint_coord"], cfg)) ** 2
    loss["obj_seg"] = F.binary_cross_entropy(decoder_out[:, 2], targets["obj_seg"], reduction="none")
    loss["hand_seg"] = F.binary_cross_entropy(decoder_out[:, 1], targets["hand_seg"], reduction="none")
    out.update(hot_path_eval(p, pyramid, meta, cfg, taps))
    pred = {"verts3d": taps["mano_verts"], "joints3d": taps["mano_joints"], "mano_shape": taps["mano_shape"]}
    L, N, B, _ = taps["mano_pose6d"].shape
    pred["mano_pose"] = rot6d_to_mat(taps["mano_pose6d"].permute(0, 2, 1, 3).reshape(L * B * N, 6)).view(L, B, N, 3, 3)
    gt = mano_head_gt(p, targets["mano_param"])
    out["mano_joints_gt_out"], out["mano_mesh_gt_out"] = gt["joints3d"], gt["verts3d"]
    l3d, lcls, lall = joint_vote_losses(taps["hand_points_notrans"], taps["hand_off"], taps["hand_cls"],
                                        taps["hand_joints"], targets["joint_cam_no_trans"][:, 1:], cfg)
    loss.update(loss_joint_3d=l3d, loss_joint_cls=lcls, loss_all_joint_3d=lall)
    loss.update(mano_losses(pred, gt, cfg))
    obj_rot, obj_trans = taps["obj_rot"], taps["obj_trans"]
    loss["obj_rot"] = F.smooth_l1_loss(obj_rot, targets["obj_rot"][None, None].expand_as(obj_rot))
    loss["obj_trans"] = F.smooth_l1_loss(obj_trans, targets["rel_obj_trans"][None, None].expand_as(obj_trans))
    return {**loss, **out}


def model_train(p, img, inputs, targets, meta, cfg, arch="ho3d", noise=None, taps=None):
    """upstream Model.forward(mode='train') (main/model.py:357-665) in the branch the first `cfg.point_sampling_epoch`
    epochs take (:426-466: the pose branch works on `*_pre_points` + uniform jitter; `noise` = (hand, obj) jitter tensors
    or None for zero jitter), with BatchNorm on batch statistics (model.train()) and every dropout at p = 0 (the parity
    configuration; dropout masks are not reproducible across implementations).  Plain differentiable torch ops: when the
    tensors of `p` require grad, `sum(weighted means).backward()` gives the reference gradients (main/train.py:111-131).
    Returns {**loss, **out} like upstream."""
    feat, skips = backbone(p, img, training=True)
    pyramid, decoder_out = unet_decoder(p, feat, skips, arch, training=True)
    taps = {} if taps is None else taps
    taps["pyramid"], taps["decoder_out"] = pyramid, decoder_out
    root, objc, K = meta["mano_root"], meta["obj_center_cam"], meta["cam_intr"]
    c = cfg.ClampingDistance
    loss = {}
    hand_s, _, _ = sdf_forward(p, pyramid, inputs["hand_sdf_points"], root, K, cfg.hand_sdf_scale, "hand", cfg)
    obj_s, _, _ = sdf_forward(p, pyramid, inputs["obj_sdf_points"], objc, K, cfg.obj_sdf_scale, "obj", cfg)
    loss["sdfhand_loss"] = F.l1_loss(hand_s, targets["hand_sdf"].clamp(-c, c).unsqueeze(-1))
    loss["sdfobj_loss"] = F.l1_loss(obj_s, targets["obj_sdf"].clamp(-c, c).unsqueeze(-1))
    out = {"joint_heatmap_out": decoder_out[:, 0], "hand_seg_gt_out": targets["hand_seg"],
           "hand_seg_pred_out": decoder_out[:, 1], "obj_seg_gt_out": targets["obj_seg"],
           "obj_seg_pred_out": decoder_out[:, 2]}
    loss["joint_heatmap"] = (decoder_out[:, 0] - render_gaussian_heatmap(targets["joint_coord"], cfg)) ** 2
    loss["obj_seg"] = F.binary_cross_entropy(decoder_out[:, 2], targets["obj_seg"], reduction="none")
    loss["hand_seg"] = F.binary_cross_entropy(decoder_out[:, 1], targets["hand_seg"], reduction="none")
    hand_points = inputs["hand_pre_points"] + (0.0 if noise is None else noise[0])
    obj_points = inputs["obj_pre_points"] + (0.0 if noise is None else noise[1])
    hand_sdf, _, hand_pe = sdf_forward(p, pyramid, hand_points, root, K, cfg.hand_sdf_scale, "hand", cfg)
    obj_sdf, _, obj_pe = sdf_forward(p, pyramid, obj_points, objc, K, cfg.obj_sdf_scale, "obj", cfg)
    pose = pose_from_points(p, pyramid, meta, cfg, hand_points, hand_sdf, hand_pe, obj_points, obj_sdf, obj_pe, taps)
    out.update({k: v for k, v in pose.items() if k not in ("obj_rot_out", "obj_trans_out")})     # model.py:618-620
    pred = {"verts3d": taps["mano_verts"], "joints3d": taps["mano_joints"], "mano_shape": taps["mano_shape"]}
    L, N, B, _ = taps["mano_pose6d"].shape
    pred["mano_pose"] = rot6d_to_mat(taps["mano_pose6d"].permute(0, 2, 1, 3).reshape(L * B * N, 6)).view(L, B, N, 3, 3)
    gt = mano_head_gt(p, targets["mano_param"])
    l3d, lcls, lall = joint_vote_losses(taps["hand_points_notrans"], taps["hand_off"], taps["hand_cls"],
                                        taps["hand_joints"], targets["joint_cam_no_trans"][:, 1:], cfg)
    loss.update(loss_joint_3d=l3d, loss_joint_cls=lcls, loss_all_joint_3d=lall)
    loss.update(mano_losses(pred, gt, cfg))
    obj_rot, obj_trans = taps["obj_rot"], taps["obj_trans"]
    loss["obj_rot"] = F.smooth_l1_loss(obj_rot, targets["obj_rot"][None, None].expand_as(obj_rot))
    loss["obj_trans"] = F.smooth_l1_loss(obj_trans, targets["rel_obj_trans"][None, None].expand_as(obj_trans))
    return {**loss, **out}


# loss weights of upstream main/train.py:115-128 with the values of main/config.py:136-145
TRAIN_LOSS_WEIGHTS = {"sdfhand_loss": 50.0, "sdfobj_loss": 25.0, "joint_heatmap": 100.0 / 100000, "obj_seg": 1.0,
                      "hand_seg": 1.0, "obj_rot": 0.7, "obj_trans": 100.0, "loss_joint_3d": 0.1, "loss_joint_cls": 1.0,
                      "loss_all_joint_3d": 0.1}


def train_total_loss(model_out):
    """upstream main/train.py:111-131: mean of every non-`_out` entry, weighted, summed (the scalar `.backward()` runs on)."""
    loss = {k: v.mean() * TRAIN_LOSS_WEIGHTS.get(k, 1.0) for k, v in model_out.items() if "_out" not in k}
    return sum(loss.values()), loss


# ----------------------------------------------------------------------------------------------------
# Test-time metrics (upstream common/metrics.py) -- the consumer of obj_rot_out / obj_trans_out / mano_joints_out
# ----------------------------------------------------------------------------------------------------
def batch_rodrigues(axisang):
    """manopth/manopth/rodrigues_layer.py:15-56: axis-angle (N,3) -> quaternion -> rotation matrices (N,3,3)."""
    angle = torch.norm(axisang + 1e-8, p=2, dim=1, keepdim=True)
    normalized = axisang / angle
    half = angle * 0.5
    quat = torch.cat([torch.cos(half), torch.sin(half) * normalized], dim=1)
    quat = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = quat[:, 0], quat[:, 1], quat[:, 2], quat[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1).view(-1, 3, 3)


def mesh_metrics(pred_meshes, target_meshes):
    """common/metrics.py:62-108 (compute_obj_metrics_dexycb / _ho3d): per sample ADD-S (mean over predicted vertices of
    the distance to the closest target vertex), MME (mean per-vertex distance) and MCE (mean distance between the 8
    corners of the two axis-aligned bounding boxes).  (B,N,3) x2 -> three (B) tensors."""
    dis = (target_meshes[:, None, :, :] - pred_meshes[:, :, None, :]).norm(dim=-1)          # [b, i(pred), j(target)]
    adds = dis.min(dim=2)[0].mean(dim=1)
    mme = (target_meshes - pred_meshes).norm(2, -1).mean(-1)
    sel = torch.tensor([[0, 1, 0, 0, 1, 0, 1, 1], [0, 0, 1, 0, 1, 1, 0, 1], [0, 0, 0, 1, 0, 1, 1, 1]])

    def corners(m):
        mm = torch.stack([m.min(dim=1)[0], m.max(dim=1)[0]], dim=2)                          # (B, 3, 2)
        return torch.stack([mm[:, 0, sel[0]], mm[:, 1, sel[1]], mm[:, 2, sel[2]]], dim=2)   # (B, 8, 3)

    mce = (corners(pred_meshes) - corners(target_meshes)).norm(2, -1).mean(-1)
    return adds, mme, mce


def obj_pose_metrics(templates, obj_ids, rot_votes, trans_votes, rot_gt, trans_gt):
    """common/metrics.py:110-185 (eval_batched_obj_direct) per sample, before the batch means: mean of the pose votes
    (:115-116), posed template meshes (:147-167), then mesh_metrics and OCE = |trans - trans_gt| (:172,179).
    templates (T,N,3), obj_ids (B) -> adds, mme, mce, oce, each (B)."""
    rot, trans = rot_votes.mean(1), trans_votes.mean(1)
    tm = templates[obj_ids]
    target = torch.bmm(tm, batch_rodrigues(rot_gt).permute(0, 2, 1)) + trans_gt[:, None, :]
    pred = torch.bmm(tm, batch_rodrigues(rot).permute(0, 2, 1)) + trans[:, None, :]
    adds, mme, mce = mesh_metrics(pred, target)
    return adds, mme, mce, torch.norm(trans - trans_gt, dim=-1)


def rigid_align(A, B):
    """common/metrics.py:188-213 (rigid_transform_3D + rigid_align), numpy like upstream: A (N,3) mapped onto B by the
    best similarity transform (Umeyama: SVD of the cross-covariance, reflection fixed on the last singular vector)."""
    import numpy as np
    A, B = np.asarray(A), np.asarray(B)
    n = A.shape[0]
    ca, cb = A.mean(axis=0), B.mean(axis=0)
    H = (A - ca).T @ (B - cb) / n
    U, s, V = np.linalg.svd(H)
    R = V.T @ U.T
    if np.linalg.det(R) < 0:
        s[-1] = -s[-1]
        V[2] = -V[2]
        R = V.T @ U.T
    c = 1 / np.var(A, axis=0).sum() * s.sum()
    t = -(c * R) @ ca + cb
    return (c * R @ A.T).T + t


def hand_joint_metrics(pred, gt):
    """common/metrics.py:231-248 (eval_hand_joint) per sample, before the means: (B,J,3) x2 -> mje (B), pamje (B)."""
    import numpy as np
    pred, gt = np.asarray(pred), np.asarray(gt)
    mje = [np.sqrt(((p - g) ** 2).sum(1)).mean() for p, g in zip(pred, gt)]
    pamje = [np.sqrt(((rigid_align(p, g) - g) ** 2).sum(1)).mean() for p, g in zip(pred, gt)]
    return torch.tensor(np.array(mje, dtype=np.float32)), torch.tensor(np.array(pamje, dtype=np.float32))
