"""ResNet-50 encoder (upstream common/nets/resnet.py:14-98 = torchvision Bottleneck ResNet, stride on the 3x3) on the
FP16x3 tensor-core kernels -- the step in front of the U-Net decoder.

cuDNN runs these 53 convolutions on the fp32 FMA pipe (12.5 ms of a 59 ms step at batch 32).  Here:

  * the 7x7 stride-2 stem is an im2col pass (`hoisdf_stem_im2col_split`, K = 147 -> 160) + one FP16x3 Linear;
  * max-pool 3x3/2 runs on the split-half NHWC map (`hoisdf_maxpool3x3s2_split`);
  * every bottleneck is  1x1 Linear -> 3x3 implicit-GEMM convolution (stride 1 or 2 through the TMA element stride)
    -> 1x1 Linear whose epilogue adds the shortcut (split-half residual) and applies the ReLU;
    the projection shortcut is a 1x1 (stride-2) implicit-GEMM convolution;
  * BatchNorm (eval) is folded into weights and bias; activations stay NHWC split-half end to end, and each stage's
    output is written straight into the channel window of the U-Net's concat buffer that will consume it.

Parameters stay in the `ResNetBackbone` module (upstream names, strict checkpoint loading); this file only packs them
(cached until a parameter or buffer changes) and launches kernels.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .. import ops
from ..config import cfg
from .unet_h3 import TAPS_1X1, TAPS_3X3, SplitMap, _fold_bn, _pack_conv


class ResNetH3:
    STAGES = (("layer1", "stride4"), ("layer2", "stride8"), ("layer3", "stride16"), ("layer4", "stride32"))

    def __init__(self, resnet: torch.nn.Module):
        self.net = resnet
        self._packed = None
        self._key = None

    def _pack(self):
        key = tuple((p.data_ptr(), p._version) for p in list(self.net.parameters()) + list(self.net.buffers())) + \
            (int(cfg.backbone_chunk_kb),)
        if self._packed is not None and self._key == key:
            return self._packed
        n, pk = self.net, {}
        w, b = _fold_bn(n.conv1.weight, None, n.bn1, 0)
        # K index = (ky * 7 + kx) * 3 + c -- the column order hoisdf_stem_im2col_split writes
        ck = int(cfg.backbone_chunk_kb)
        pk["stem"] = ops.PackedLinearH3.pack(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).float().contiguous(), b,
                                             chunk_kb=ck)
        for lname, _ in self.STAGES:
            blocks = []
            for blk in getattr(n, lname):
                d = {"conv1": _pack_conv(blk.conv1, blk.bn1, ck), "conv2": _pack_conv(blk.conv2, blk.bn2, ck),
                     "conv3": _pack_conv(blk.conv3, blk.bn3, ck), "stride": int(blk.conv2.stride[0]), "down": None}
                if blk.downsample is not None:
                    d["down"] = _pack_conv(blk.downsample[0], blk.downsample[1], ck)
                blocks.append(d)
            pk[lname] = blocks
        self._packed, self._key = pk, key
        return pk

    @staticmethod
    def _bottleneck(x: SplitMap, d: dict, out: Optional[ops.SplitRows]) -> SplitMap:
        b, h, w, cin = x.b, x.h, x.w, x.c
        s = d["stride"]
        ho, wo = h // s, w // s
        dev = x.rows.buf.device
        planes, cout = d["conv1"].n, d["conv3"].n
        h1 = ops.linear_h3(x.rows, d["conv1"], ops.ACT_RELU, split_out=True)
        h2 = ops.conv_h3(h1, b, h, w, planes, d["conv2"], TAPS_3X3, ho, wo, stride=s, act=ops.ACT_RELU,
                         out=ops.SplitRows.empty(b * ho * wo, planes, dev))
        if d["down"] is None:
            idt = x.rows
        elif s == 1:
            idt = ops.linear_h3(x.rows, d["down"], ops.ACT_NONE, split_out=True)
        else:
            idt = ops.conv_h3(x.rows, b, h, w, cin, d["down"], TAPS_1X1, ho, wo, stride=s, act=ops.ACT_NONE,
                              out=ops.SplitRows.empty(b * ho * wo, cout, dev))
        if out is None:
            out = ops.SplitRows.empty(b * ho * wo, cout, dev)
        ops.linear_h3(h2, d["conv3"], ops.ACT_RELU, out=out, residual_split=idt)
        return SplitMap(out, b, ho, wo, cout)

    def __call__(self, img: torch.Tensor, slots: Optional[Dict[str, ops.SplitRows]] = None):
        """img (B, 3, H, W) fp32 -> (img_feat SplitMap (B, H/32, W/32, 2048), skips {stride2..stride16: SplitMap}).
        `slots[name]`, when given, is the split-half window the map `name` must be written to."""
        pk = self._pack()
        slots = slots or {}
        b, _, hh, ww = img.shape
        h, w = hh // 2, ww // 2
        dev = img.device
        cols = ops.stem_im2col(img.to(torch.float32))
        s2 = slots.get("stride2") or ops.SplitRows.empty(b * h * w, pk["stem"].n, dev)
        ops.linear_h3(cols, pk["stem"], ops.ACT_RELU, out=s2)
        skips = {"stride2": SplitMap(s2, b, h, w, pk["stem"].n)}
        x = SplitMap(ops.maxpool3x3s2(s2, b, h, w, pk["stem"].n), b, h // 2, w // 2, pk["stem"].n)
        for lname, sname in self.STAGES:
            blocks = pk[lname]
            for i, d in enumerate(blocks):
                x = self._bottleneck(x, d, slots.get(sname) if i == len(blocks) - 1 else None)
            skips[sname] = x
        return skips.pop("stride32"), skips
