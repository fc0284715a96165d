"""The training step (BASELINE configs[3]): `Model.forward(mode="train")` and the optimiser step,
mirroring upstream main/model.py:357-665 (train branch) and main/train.py:104-140 / common/base.py:64-70.

What runs where:
  * image encoder (ResNet-50 + U-Net, BatchNorm with batch statistics): the cuDNN modules of nets/module.py under
    PyTorch autograd ("keep cuDNN" for the step before the hot path; the FP16x3 convolution kernels
    have no dgrad / wgrad form yet);
  * everything after the pyramid -- the hot path -- on the hoisdf_b200 kernels through hoisdf_b200/autograd.py:
    bilinear gathers (forward + scatter-add backward), every Linear of linear_sdfin / SDFDecoder (weight-norm) /
    linear_transformerin / the transformers / the heads on the FP16x3 tcgen05 GEMM in forward AND backward (dX, dW),
    attention (tcgen05 flash forward, batched fp32 backward), LayerNorm, token assembly;
  * MANO forward/backward and the scalar loss formulas: plain torch ops (bookkeeping-sized; nets/mano_torch.py,
    nets/loss.py);
  * AdamW: `hoisdf_adamw_step` over ONE flat parameter / gradient buffer; multi-GPU: one all-reduce of that flat
    gradient buffer per step (data parallel over samples, like upstream's DataParallel).
"""
from __future__ import annotations

import os
import random
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import autograd as A
from . import ops
from ._capi import check, lib
from .config import cfg
from .nets.loss import joint_vote_losses, render_gaussian_heatmap
from .nets.mano_torch import mano_head_train

ACT_NONE, ACT_RELU = ops.ACT_NONE, ops.ACT_RELU
# encoder layout in training: channels_last activations let cuDNN's tensor-op convolutions skip their NCHW <-> NHWC
# conversion kernels and make the pyramid's NHWC view free: measured 280 -> 247 ms per step (HOISDF_TRAIN_CL=0: NCHW).
# (PyTorch's native BatchNorm kernels instead of cuDNN's were measured too: 7-15 ms slower.)
_ENC_CHANNELS_LAST = os.environ.get("HOISDF_TRAIN_CL", "1") != "0"


def _drop(x, p):
    return F.dropout(x, p, True) if p > 0.0 else x


# ----------------------------------------------------------------------------------------------------
# per-point branches
# ----------------------------------------------------------------------------------------------------
def mlp_rows(mlp, x2d):
    """upstream common/nets/layer.py:192-201 on (rows, K)."""
    h = x2d
    n = len(mlp.layers)
    for i, lin in enumerate(mlp.layers):
        act = ACT_RELU if (i < n - 1 or mlp.is_activation_last) else ACT_NONE
        h = A.linear(h, lin.weight, lin.bias, act)
    return h


def sdf_decoder_rows(dec, x):
    """upstream common/nets/sdf_net.py:87-122 on (rows, 289) -> tanh(sdf) (rows, 1); dropout after every ReLU."""
    p = float(dec.dropout_prob)
    w = lambda l: A.WeightNormFn.apply(l.weight_g, l.weight_v)     # noqa: E731
    h = _drop(A.linear(x, w(dec.linh0), dec.linh0.bias, ACT_RELU), p)
    h = _drop(A.linear(h, w(dec.linh1), dec.linh1.bias, ACT_RELU), p)
    h = torch.cat([h, x], 1)
    h = _drop(A.linear(h, w(dec.linh2), dec.linh2.bias, ACT_RELU), p)
    h = _drop(A.linear(h, w(dec.linh3), dec.linh3.bias, ACT_RELU), p)
    return torch.tanh(A.linear(h, dec.linh4.weight, dec.linh4.bias, ACT_NONE))


def gather_rows(maps, pts, center, K, scale, want_cam):
    b, p, _ = pts.shape
    cam, uv = ops.project_points(pts, center, K, scale, want_cam=want_cam)
    return A.GatherFn.apply(uv, b, p, tuple(cfg.input_img_shape), *maps), cam


def sdf_forward(model, maps, sdf_points, center, K, scale, kind):
    """upstream main/model.py:181-244 -> (sdf (B,P,1) clamped, posenc (B,P,30))."""
    pts = sdf_points.detach().to(torch.float32).contiguous()
    b, p, _ = pts.shape
    feats, _ = gather_rows(maps, pts, center, K, scale, False)
    fea = mlp_rows(model.linear_sdfin, feats)
    rows = torch.empty(b * p, ops.ROW_LD, device=pts.device, dtype=torch.float32)
    ops.posenc(rows, points=pts.view(b * p, 3), bins=cfg.bins_n)          # columns 256..288 = posenc (30) | xyz (3)
    dec = model.hand_sdf_decoder if kind == "hand" else model.obj_sdf_decoder
    sdf = sdf_decoder_rows(dec, torch.cat([fea, rows[:, 256:289]], 1))
    c = cfg.ClampingDistance
    return torch.clamp(sdf, -c, c).view(b, p, 1), rows[:, 256:286].reshape(b, p, 30)


def input_transformer(model, maps, sdf_points, center, K, scale):
    """upstream main/model.py:145-179 -> (latent (B,P,223), cam points (B,P,3))."""
    pts = sdf_points.detach().to(torch.float32).contiguous()
    b, p, _ = pts.shape
    feats, cam = gather_rows(maps, pts, center, K, scale, True)
    return mlp_rows(model.linear_transformerin, feats).view(b, p, -1), cam


# ----------------------------------------------------------------------------------------------------
# transformers (batch-major rows: row = b * L + token)
# ----------------------------------------------------------------------------------------------------
def mha(attn, q_in, k_in, v_in, b, lq, lk, mask, kv_valid, p_drop):
    """nn.MultiheadAttention (in_proj [q;k;v], 4 heads of 64, out_proj) on row matrices."""
    d = attn.embed_dim
    W, bias = attn.in_proj_weight, attn.in_proj_bias
    if q_in is k_in and k_in is v_in:
        qkv = A.linear(q_in, W, bias)
        q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    elif q_in is k_in:
        qk = A.linear(q_in, W[:2 * d], bias[:2 * d])
        q, k, v = qk[:, :d], qk[:, d:], A.linear(v_in, W[2 * d:], bias[2 * d:])
    elif k_in is v_in:
        q = A.linear(q_in, W[:d], bias[:d])
        kv = A.linear(k_in, W[d:], bias[d:])
        k, v = kv[:, :d], kv[:, d:]
    else:
        q, k, v = A.linear(q_in, W[:d], bias[:d]), A.linear(k_in, W[d:2 * d], bias[d:2 * d]), \
            A.linear(v_in, W[2 * d:], bias[2 * d:])
    if k.stride(0) != v.stride(0):
        k, v = k.contiguous(), v.contiguous()
    att = A.AttentionFn.apply(q, k, v, b, attn.num_heads, lq, lk, mask, kv_valid, p_drop)
    return A.linear(att, attn.out_proj.weight, attn.out_proj.bias)


def _ln(norm, x, res=None):
    return A.AddLayerNormFn.apply(x, res, norm.weight, norm.bias)


def _ffn(layer, x, p):
    h = _drop(A.linear(x, layer.linear1.weight, layer.linear1.bias, ACT_RELU), p)
    return A.linear(h, layer.linear2.weight, layer.linear2.bias)


def encoder_layer(layer, x, b, s, p):
    """upstream transformer.py:279-302 (forward_post; pos_embed is all zeros upstream, model.py:541-543)."""
    x1 = _ln(layer.norm1, _drop(mha(layer.self_attn, x, x, x, b, s, s, None, None, p), p), x)
    return _ln(layer.norm2, _drop(_ffn(layer, x1, p), p), x1)


def encoder(enc, x, b, s, p):
    """upstream transformer.py:175-202 -> (last output, [inter_norm(out_l)])."""
    inter = []
    for layer in enc.layers:
        x = encoder_layer(layer, x, b, s, p)
        if enc.return_intermediate:
            inter.append(_ln(enc.inter_norm, x))
    if enc.norm is not None:
        x = _ln(enc.norm, x)
    return x, inter


def decoder_layer(layer, t, mem, qpos, b, lq, s, tgt_mask, kv_valid, p):
    """upstream transformer.py:366-395 (forward_post)."""
    qk = t + qpos
    t1 = _ln(layer.norm1, _drop(mha(layer.self_attn, qk, qk, t, b, lq, lq, tgt_mask, None, p), p), t)
    t2 = _ln(layer.norm2, _drop(mha(layer.multihead_attn, t1 + qpos, mem, mem, b, lq, s, None, kv_valid, p), p), t1)
    return _ln(layer.norm3, _drop(_ffn(layer, t2, p), p), t2)


def decoder(dec, mem, query_embed, b, s, tgt_mask, kv_valid, p):
    """upstream transformer.py:214-252 -> [norm(out_l)] per layer, each (B*Lq, d)."""
    lq, d = query_embed.shape
    qpos = query_embed.unsqueeze(0).expand(b, lq, d).reshape(b * lq, d)
    t = torch.zeros(b * lq, d, device=mem.device, dtype=torch.float32)
    hs = []
    for layer in dec.layers:
        t = decoder_layer(layer, t, mem, qpos, b, lq, s, tgt_mask, kv_valid, p)
        hs.append(_ln(dec.norm, t))
    return hs


def vote_joints(points, off, cls):
    """upstream common/nets/loss.py:31-36,54-57, batch-major: (B,P,3), (L,B,P,60), (L,B,P,20) -> (L,B,20,3)."""
    l, b, p, j = cls.shape
    vote = points[None, :, :, None, :] + off.view(l, b, p, j, 3)
    w = torch.softmax(cls, dim=2).unsqueeze(-1)
    return (vote * w).sum(2)


_TAIL_LOSSES = ("loss_joint_3d", "loss_joint_cls", "loss_all_joint_3d", "mano_mesh_loss", "mano_joint_loss", "pose_param_loss",
                "shape_param_loss", "obj_rot", "obj_trans")
_GRAPH_TAIL = os.environ.get("HOISDF_TRAIN_GRAPH_TAIL", "1") != "0"
_tail_graphs = {}


def _tail_fn(mano_head, consts):
    """MANO head on the predicted parameters, joint votes and every pose loss (upstream main/model.py:572-665): plain torch ops
    on bookkeeping-sized tensors.  -> (mano_mesh_out, mano_joints_out, hand_joints of all layers, the 9 loss entries of
    _TAIL_LOSSES)."""
    l_verts, l_joints, l_pose, l_shape = consts

    def fn(pose6d, shape, hand_off, hand_cls, obj_rot, obj_trans, hand_nt, joint_gt, gt_verts, gt_joints, gt_pose, gt_shape,
           rot_gt, trans_gt):
        pred = mano_head_train(mano_head, pose6d, shape)
        hand_joints = vote_joints(hand_nt, hand_off, hand_cls)
        l3d, lcls, lall = joint_vote_losses(hand_nt, hand_off, hand_cls, hand_joints, joint_gt)
        exp = lambda t, like: t.unsqueeze(0).expand(like.shape)     # noqa: E731
        return (pred["verts3d"][-1], pred["joints3d"][-1], hand_joints, l3d, lcls, lall,
                l_verts * F.mse_loss(pred["verts3d"], exp(gt_verts, pred["verts3d"])),
                l_joints * F.mse_loss(pred["joints3d"], exp(gt_joints, pred["joints3d"])),
                l_pose * F.mse_loss(pred["mano_pose"], exp(gt_pose, pred["mano_pose"])),
                l_shape * F.mse_loss(pred["mano_shape"], exp(gt_shape, pred["mano_shape"])),
                F.smooth_l1_loss(obj_rot, rot_gt[None, :, None, :].expand_as(obj_rot)),
                F.smooth_l1_loss(obj_trans, trans_gt[None, :, None, :].expand_as(obj_trans)))
    return fn


def _graphed_tail(model, args):
    """The tail of the training forward, eagerly or -- default on CUDA -- as a pair of CUDA graphs (forward and backward) made
    by torch.cuda.make_graphed_callables, one pair per (shapes, loss constants): static shapes, no host read-back, no random
    numbers in there.  HOISDF_TRAIN_GRAPH_TAIL=0, any input without gradient (a frozen head) or a DataParallel replica takes the
    eager form."""
    consts = (float(cfg.lambda_verts3d), float(cfg.lambda_joints3d), float(cfg.lambda_manopose), float(cfg.lambda_manoshape))
    fn = _tail_fn(model.mano_head, consts)
    grads = tuple(a.requires_grad for a in args)
    # (nn.DataParallel rebuilds its replicas -- new module objects, new buffer copies -- on every forward: nothing to key a
    # captured graph on, so replicas take the eager form)
    if not (_GRAPH_TAIL and args[0].is_cuda and torch.is_grad_enabled() and all(grads[:6])
            and not torch.cuda.is_current_stream_capturing() and not getattr(model, "_is_replica", False)):
        return fn(*args)
    key = (id(model.mano_head), consts, float(cfg.hand_cls_dist), args[0].device, tuple(tuple(a.shape) for a in args))
    graphed = _tail_graphs.get(key)
    if graphed is None:
        sample = tuple(a.detach().clone().requires_grad_(g) for a, g in zip(args, grads))
        import warnings
        with warnings.catch_warnings():
            # (the capture's warm-up runs on a side stream: autograd notes that the sample leaves were made on another one)
            warnings.simplefilter("ignore")
            graphed = torch.cuda.make_graphed_callables(fn, sample)
        _tail_graphs[key] = graphed
    return graphed(*[a.contiguous() for a in args])


# ----------------------------------------------------------------------------------------------------
# the training forward
# ----------------------------------------------------------------------------------------------------
def forward_train(model, inputs, targets, meta_info, epoch_cnt=1e8, batch_ratio=0) -> Dict[str, torch.Tensor]:
    """upstream Model.forward(mode="train") (main/model.py:357-665): returns {**loss, **out} -- unreduced loss entries
    (main/train.py:111-113 takes their means, weights and sums them) and the `*_out` tensors."""
    if not ops.use_h3():
        raise RuntimeError("the training step runs on the FP16x3 tensor-core kernels (HOISDF_TC=1, TC_MODE 'h3')")
    img = inputs["img"]
    root = meta_info["mano_root"].to(torch.float32).contiguous()
    objc = meta_info["obj_center_cam"].to(torch.float32).contiguous()
    K = meta_info["cam_intr"].to(torch.float32).contiguous()
    b = img.shape[0]
    Ph, Po = int(cfg.num_samp_hand), int(cfg.num_samp_obj)
    hs_scale, os_scale = cfg.hand_sdf_scale, cfg.obj_sdf_scale
    c = cfg.ClampingDistance
    loss, out = {}, {}

    if _ENC_CHANNELS_LAST:
        img = img.contiguous(memory_format=torch.channels_last)
    img_feat, skips = model.backbone_net(img)
    pyramid, decoder_out = model.decoder_net(img_feat, skips)
    maps = [pyramid[name].permute(0, 2, 3, 1).contiguous() for name in cfg.mutliscale_layers]     # NHWC, once

    # SDF supervision (model.py:370-401)
    hand_s, _ = sdf_forward(model, maps, inputs["hand_sdf_points"], root, K, hs_scale, "hand")
    obj_s, _ = sdf_forward(model, maps, inputs["obj_sdf_points"], objc, K, os_scale, "obj")
    loss["sdfhand_loss"] = F.l1_loss(hand_s, targets["hand_sdf"].clamp(-c, c).unsqueeze(-1))
    loss["sdfobj_loss"] = F.l1_loss(obj_s, targets["obj_sdf"].clamp(-c, c).unsqueeze(-1))
    out["joint_heatmap_out"] = decoder_out[:, 0]
    out["hand_seg_gt_out"] = targets["hand_seg"]
    out["hand_seg_pred_out"] = decoder_out[:, 1]
    out["obj_seg_gt_out"] = targets["obj_seg"]
    out["obj_seg_pred_out"] = decoder_out[:, 2]
    loss["joint_heatmap"] = (decoder_out[:, 0] - render_gaussian_heatmap(targets["joint_coord"])) ** 2
    loss["obj_seg"] = F.binary_cross_entropy(decoder_out[:, 2], targets["obj_seg"], reduction="none")
    loss["hand_seg"] = F.binary_cross_entropy(decoder_out[:, 1], targets["hand_seg"], reduction="none")

    # the points the pose branch works on (model.py:424-481)
    if random.uniform(0, 1) < 0.4 or epoch_cnt < cfg.point_sampling_epoch:
        dist = cfg.random_move_dist[len([a for a in cfg.random_ratio if batch_ratio > a])]
        hand_points = inputs["hand_pre_points"] + torch.empty_like(inputs["hand_pre_points"]).uniform_(-dist, dist)
        obj_points = inputs["obj_pre_points"] + torch.empty_like(inputs["obj_pre_points"]).uniform_(-dist, dist)
        # upstream runs these two queries with the tape on, but consumes their results only detached (`hand_sdf.detach()`,
        # model.py:483-484) or as constants (the positional encoding): no gradient ever flows through them
        with torch.no_grad():
            hand_sdf, hand_pe = sdf_forward(model, maps, hand_points, root, K, hs_scale, "hand")
            obj_sdf, obj_pe = sdf_forward(model, maps, obj_points, objc, K, os_scale, "obj")
    else:
        with torch.no_grad():       # the inference selection on the tensor-core cascade, from a detached pyramid
            dpyr = {k: v.detach() for k, v in pyramid.items()}
            plans = model._plans(meta_info)
            ctx = model._ctx(dpyr)
            hand_points, hand_sdf, hand_pe, _ = model.sdf_infer(ctx, root, K, None, hs_scale, Ph, "hand", plans[0])
            obj_points, obj_sdf, obj_pe, _ = model.sdf_infer(ctx, objc, K, None, os_scale, Po, "obj", plans[1])
    hand_points, obj_points = hand_points.detach(), obj_points.detach()
    Ph, Po = hand_points.shape[1], obj_points.shape[1]
    S = Ph + Po

    model.hand_sigmoid_beta.data.clamp_(min=2e-3)          # model.py:124
    model.obj_sigmoid_beta.data.clamp_(min=2e-3)
    hand_fea, hand_cam = input_transformer(model, maps, hand_points, root, K, hs_scale)
    obj_fea, obj_cam = input_transformer(model, maps, obj_points, objc, K, os_scale)
    hand_nt = hand_cam - root[:, None, :]
    obj_nt = obj_cam - objc[:, None, :]
    hand_o_nt = hand_cam - objc[:, None, :]                # model.py:498 ("bug": unscaled coords, kept)
    obj_h_nt = obj_cam - root[:, None, :]                  # model.py:508
    with torch.no_grad():                                  # cross-field tokens are detached upstream (model.py:536,555)
        hand_o_sdf, hand_o_pe = sdf_forward(model, maps, hand_o_nt * os_scale, objc, K, os_scale, "obj")
        obj_h_sdf, obj_h_pe = sdf_forward(model, maps, obj_h_nt * hs_scale, root, K, hs_scale, "hand")
        cross_h = torch.empty(b, Po, 256, device=img.device, dtype=torch.float32)
        ops.tokens(obj_h_nt.contiguous(), obj_h_pe.contiguous(), obj_fea.detach().contiguous(), obj_h_sdf.contiguous(),
                   model.hand_sigmoid_beta.data, cross_h, 0)
        cross_o = torch.empty(b, Ph, 256, device=img.device, dtype=torch.float32)
        ops.tokens(hand_o_nt.contiguous(), hand_o_pe.contiguous(), hand_fea.detach().contiguous(), hand_o_sdf.contiguous(),
                   model.obj_sigmoid_beta.data, cross_o, 0)
    own_h = A.TokensFn.apply(hand_fea, model.hand_sigmoid_beta, hand_nt, hand_pe.detach(), hand_sdf.detach())
    own_o = A.TokensFn.apply(obj_fea, model.obj_sigmoid_beta, obj_nt, obj_pe.detach(), obj_sdf.detach())
    hand_in = torch.cat([own_h, cross_h], 1).view(b * S, 256)
    obj_in = torch.cat([own_o, cross_o], 1).view(b * S, 256)

    p = float(cfg.dropout)
    tgt_mask, _ = model._masks(img.device)
    ht, ot = model.hand_transformer, model.obj_transformer
    memory, hand_inter = encoder(ht.encoder, hand_in, b, S, p)
    hs = decoder(ht.decoder, memory, model.mano_query_embed.weight, b, S, tgt_mask.to(torch.uint8).contiguous(), Ph, p)
    _, obj_inter = encoder(ot.encoder, obj_in, b, S, p)

    Le, Lo, Ld = len(hand_inter), len(obj_inter), len(hs)
    nq = cfg.mano_num_queries
    hand_enc = torch.stack(hand_inter).view(Le, b, S, 256)[:, :, :Ph].reshape(Le * b * Ph, 256)
    obj_enc = torch.stack(obj_inter).view(Lo, b, S, 256)[:, :, :Po].reshape(Lo * b * Po, 256)
    hs_all = torch.stack(hs).view(Ld, b, nq, 256)
    hand_off = mlp_rows(model.linear_handvote, hand_enc).view(Le, b, Ph, 60)
    hand_cls = mlp_rows(model.linear_handcls, hand_enc).view(Le, b, Ph, 20)
    obj_rot = mlp_rows(model.linear_obj_rot, obj_enc).view(Lo, b, Po, 3)
    obj_trans = mlp_rows(model.linear_obj_rel_trans, obj_enc).view(Lo, b, Po, 3)
    pose6d = mlp_rows(model.linear_pose, hs_all[:, :, :cfg.mano_shape_indx].reshape(-1, 256)).view(
        Ld, b, cfg.mano_shape_indx, 6)
    shape = mlp_rows(model.linear_shape, hs_all[:, :, cfg.mano_shape_indx].reshape(-1, 256)).view(Ld, b, 10)

    with torch.no_grad():
        gt_mano = model.mano_head.forward_gt(targets["mano_param"])
    joint_gt = targets["joint_cam_no_trans"][:, 1:]
    # MANO + vote aggregation + the scalar loss formulas: ~700 bookkeeping-sized torch kernels forward and as many backward,
    # bound by the host's launch rate -> replayed as two CUDA graphs (_graphed_tail)
    tail = _graphed_tail(model, (pose6d, shape, hand_off, hand_cls, obj_rot, obj_trans, hand_nt.detach(),
                                 joint_gt.contiguous(), gt_mano["verts3d"], gt_mano["joints3d"], gt_mano["mano_pose"],
                                 gt_mano["mano_shape"], targets["obj_rot"].contiguous(),
                                 targets["rel_obj_trans"].contiguous()))
    # (graph outputs live in the graph's static buffers, which the next step overwrites: the returned tensors are copies)
    hand_joints = tail[2]
    out["mano_mesh_out"], out["mano_joints_out"], out["hand_joints_out"] = tail[0].clone(), tail[1].clone(), \
        hand_joints[-1].clone()
    for key, v in zip(_TAIL_LOSSES, tail[3:12]):
        loss[key] = v
    dbg = getattr(model, "_train_debug", None)
    if dbg is not None:            # developer hook (scripts/train_debug.py): live tensors whose gradients are compared
        dbg.update(hand_cls=hand_cls, hand_off=hand_off, hand_fea=hand_fea, obj_fea=obj_fea, hand_in=hand_in, obj_in=obj_in,
                   memory=memory, hand_enc=hand_enc, pose6d=pose6d, shape=shape, hand_joints=hand_joints,
                   pyramid=pyramid, hand_s=hand_s, hand_sdf=hand_sdf)
        for v in dbg.values():
            if torch.is_tensor(v) and v.requires_grad:
                v.retain_grad()
            elif isinstance(v, dict):
                for t in v.values():
                    t.retain_grad()
    model.last_taps = dict(hand_points=hand_points, obj_points=obj_points, hand_sdf=hand_sdf.detach(),
                           obj_sdf=obj_sdf.detach(), hand_fea=hand_fea.detach(), hand_transformer_in=hand_in.detach(),
                           obj_transformer_in=obj_in.detach(), hand_off=hand_off.detach(), hand_cls=hand_cls.detach(),
                           obj_rot=obj_rot.detach(), obj_trans=obj_trans.detach(), mano_pose6d=pose6d.detach(),
                           mano_shape=shape.detach())
    return {**loss, **out}


# loss weights of upstream main/train.py:115-128 (read from the config at call time like upstream; defaults = config.py:136-145)
LOSS_WEIGHTS = {"sdfhand_loss": ("sdf_hand_weight", 50.0), "sdfobj_loss": ("sdf_obj_weight", 25.0),
                "joint_heatmap": ("hm_weight", 100.0 / 100000), "obj_seg": ("obj_hm_weight", 1.0),
                "hand_seg": ("obj_hm_weight", 1.0), "obj_rot": ("obj_rot_weight", 0.7),
                "obj_trans": ("obj_trans_weight", 100.0), "loss_joint_3d": ("joint_weight", 0.1),
                "loss_joint_cls": ("cls_weight", 1.0), "loss_all_joint_3d": ("joint_weight", 0.1)}


def total_loss(model_out: Dict[str, torch.Tensor]):
    """upstream main/train.py:111-131: means of the non-`_out` entries, weighted, summed.  Returns (sum, dict of the
    weighted scalar entries)."""
    loss = {k: v.mean() for k, v in model_out.items() if "_out" not in k}
    for k, (attr, default) in LOSS_WEIGHTS.items():
        if k in loss:
            loss[k] = loss[k] * float(getattr(cfg, attr, default))
    return sum(loss.values()), loss


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """Data parallel over samples (upstream: DataParallel, common/base.py:103): every rank holds the gradient of ITS
    batch-mean loss; the global-batch gradient is their average -- ONE all-reduce of the flat gradient buffer per step.
    No-op without an initialised process group / with a single rank."""
    dist = torch.distributed
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, group=group)
        flat.mul_(1.0 / dist.get_world_size(group))
    return flat


class Trainer:
    """zero_grad -> forward("train") -> weighted loss sum -> backward -> AdamW step (upstream main/train.py:104-140 with
    common/base.py:64-70: AdamW(lr=cfg.lr) over all parameters, StepLR(cfg.lr_drop, cfg.lr_decay_gamma) per epoch).

    Parameters are re-homed into ONE flat fp32 buffer (their `.data` become views of it, state-dict keys unchanged), with a
    flat gradient buffer beside it: the optimiser is a single `hoisdf_adamw_step` launch and the data-parallel gradient
    exchange a single all-reduce.  Parameters that receive no gradient in a step (upstream: norm1, linear_objvote,
    linear_objcls, ...) are skipped by torch.optim.AdamW; here their gradient slice is zero, which moves them only by
    the decoupled weight decay -- set `skip_unused=True` (default) to restore such slices after the step."""

    def __init__(self, model, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01,
                 lr_drop: int = 20, lr_decay_gamma: float = 0.7, process_group=None, skip_unused: bool = True):
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        al = 64                                             # every tensor starts on a 256-byte boundary (TMA / float4 loads)
        n = sum((p.numel() + al - 1) // al * al for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        self.slices = []
        o = 0
        for p in self.params:
            k = p.numel()
            self.flat[o:o + k].copy_(p.data.reshape(-1))
            p.data = self.flat[o:o + k].view(p.shape)
            p.grad = self.grad[o:o + k].view(p.shape)
            self.slices.append((o, k))
            o += (k + al - 1) // al * al
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), betas, float(eps), float(weight_decay)
        self.lr_drop, self.gamma = int(lr_drop), float(lr_decay_gamma)
        self.step_count, self.epoch = 0, 0
        self.group = process_group
        self.skip_unused = skip_unused
        self._unused = None

    def zero_grad(self):
        self.grad.zero_()
        for p, (o, k) in zip(self.params, self.slices):      # autograd accumulates into these views in place
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + k].view(p.shape)

    def epoch_end(self):
        """StepLR.step() (upstream main/train.py:92 via adjust_learning_rate / base.py:66-68)."""
        self.epoch += 1
        if self.epoch % self.lr_drop == 0:
            self.lr *= self.gamma

    # ------------------------------------------------------------------ checkpoints
    def _torch_optimizer(self):
        """A stock torch.optim.AdamW / StepLR pair over ALL named parameters, as upstream builds them (common/base.py:64-
        75): used only to read and write optimizer state in torch's own format, never to step."""
        opt = torch.optim.AdamW([{"params": [p for _, p in self.model.named_parameters()]}], lr=self.lr, betas=self.betas,
                                eps=self.eps, weight_decay=self.weight_decay)
        sched = torch.optim.lr_scheduler.StepLR(opt, self.lr_drop, gamma=self.gamma)
        return opt, sched

    def state_dict(self, epoch: Optional[int] = None) -> Dict[str, object]:
        """The snapshot upstream's trainer writes (main/train.py:559-568 -> common/base.py:111-116): {"epoch", "network",
        "optimizer", "lr_scheduler"}, the network keys prefixed `module.` like the DataParallel wrapper's, the optimizer /
        scheduler entries in torch.optim's own state-dict format (per-parameter `step`, `exp_avg`, `exp_avg_sq`; parameters
        that never received a gradient have no entry, as with torch.optim) -- loadable by upstream's `load_model`
        (base.py:145-150) and by `Trainer.load_state_dict`."""
        opt, sched = self._torch_optimizer()
        unused = set(self._unused or [])
        with torch.no_grad():
            for i, (p, (o, k)) in enumerate(zip(self.params, self.slices)):
                if self.step_count == 0 or i in unused:
                    continue
                opt.state[p] = {"step": torch.tensor(float(self.step_count)),
                                "exp_avg": self.exp_avg[o:o + k].view(p.shape).clone(),
                                "exp_avg_sq": self.exp_avg_sq[o:o + k].view(p.shape).clone()}
        base_lr = self.lr / (self.gamma ** (self.epoch // self.lr_drop)) if self.lr_drop > 0 else self.lr
        sd_s = sched.state_dict()
        sd_s.update({"last_epoch": self.epoch, "_step_count": self.epoch + 1, "base_lrs": [base_lr], "_last_lr": [self.lr]})
        sd_o = opt.state_dict()
        for g in sd_o["param_groups"]:
            g["lr"], g["initial_lr"] = self.lr, base_lr
        net = {"module." + k: v.detach().clone() for k, v in self.model.state_dict().items()}
        return {"epoch": self.epoch - 1 if epoch is None else int(epoch), "network": net, "optimizer": sd_o,
                "lr_scheduler": sd_s}

    def load_state_dict(self, ckpt: Dict[str, object]) -> int:
        """Resume from a snapshot in upstream's layout (see state_dict); returns the epoch to continue with
        (upstream: `start_epoch = ckpt["epoch"] + 1`, base.py:146)."""
        from .model import load_checkpoint
        load_checkpoint(self.model, {"network": ckpt["network"]})          # strict, `module.` prefix handled; writes into
        opt, _ = self._torch_optimizer()                                      # the flat buffer (p.data are views of it)
        opt.load_state_dict(ckpt["optimizer"])                               # torch validates the layout
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = [0]
        with torch.no_grad():
            for p, (o, k) in zip(self.params, self.slices):
                st = opt.state.get(p)
                if not st:
                    continue
                self.exp_avg[o:o + k].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[o:o + k].copy_(st["exp_avg_sq"].reshape(-1))
                steps.append(int(st["step"]))
        self.step_count = max(steps)
        g = opt.param_groups[0]
        self.lr, self.betas, self.eps, self.weight_decay = float(g["lr"]), tuple(g["betas"]), float(g["eps"]), \
            float(g["weight_decay"])
        sd_s = ckpt.get("lr_scheduler")
        self.epoch = int(sd_s["last_epoch"]) if sd_s else int(ckpt["epoch"]) + 1
        if sd_s:
            self.lr_drop, self.gamma = int(sd_s.get("step_size", self.lr_drop)), float(sd_s.get("gamma", self.gamma))
        torch.autograd.graph.increment_version(self.params)
        return int(ckpt["epoch"]) + 1

    def step(self, inputs, targets, meta_info, epoch_cnt=0, batch_ratio=0.0):
        model = self.model
        model.train()
        self.zero_grad()
        out = model(inputs, targets, meta_info, "train", epoch_cnt, batch_ratio)
        total, parts = total_loss(out)
        if self.skip_unused and self._unused is None:
            # parameters the graph never reaches (upstream: norm1, linear_objvote, linear_objcls -- model.py:55,86-87) keep
            # grad None there and torch.optim.AdamW skips them.  Found once, from the tape itself (not from gradient VALUES:
            # a reached parameter may well have an all-zero gradient): the first backward runs with unset .grad fields
            for p in self.params:
                p.grad = None
            total.backward()
            self._unused = [i for i, p in enumerate(self.params) if p.grad is None]
            for p, (o, k) in zip(self.params, self.slices):
                if p.grad is not None:
                    self.grad[o:o + k].copy_(p.grad.reshape(-1))
                p.grad = self.grad[o:o + k].view(p.shape)
        else:
            total.backward()
        allreduce_mean_(self.grad, self.group)
        self.step_count += 1
        # their slices stay exactly zero; their values are restored after the step (undoes the decoupled weight decay)
        saved = [(o, k, self.flat[o:o + k].clone()) for o, k in (self.slices[i] for i in (self._unused or []))]
        check(lib.hoisdf_adamw_step(self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                    self.exp_avg_sq.data_ptr(), self.flat.numel(), self.lr, self.betas[0], self.betas[1],
                                    self.eps, self.weight_decay, self.step_count, ops._stream()), "hoisdf_adamw_step")
        for o, k, v in saved:
            self.flat[o:o + k].copy_(v)
        # the kernel wrote through raw pointers: tell PyTorch (and the packed-weight caches keyed on `_version`)
        torch.autograd.graph.increment_version(self.params)
        return total.detach(), {k: v.detach() for k, v in parts.items()}, out
