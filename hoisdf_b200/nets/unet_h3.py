"""U-Net decoder (upstream common/nets/module.py:98-218) on the FP16x3 tensor-core kernels.

The decoder is 84 % of the convolution FLOPs in front of the hot path (55.8 of 66.5
GFLOP/sample for the 'ho3d' setting) and cuDNN runs it on the fp32 FMA pipe.  Here every layer is an implicit GEMM
on `hoisdf_conv_h3_fwd` / `hoisdf_linear_h3_fwd` (fp32-grade accuracy at fp16 tensor-core rate):

  * BatchNorm (eval) is folded into the weights and bias, ReLU runs in the GEMM epilogue;
  * activations are NHWC in split-half format; the skip concatenation costs no copy: the encoder feature map is
    transposed into the left channel window of the concat buffer, the transposed convolution writes the right one;
  * ConvTranspose2d(k=4, s=2, p=1) = four 2x2-tap convolutions, one per output-pixel parity class, each writing its
    interleaved quarter of the output through a strided 4-D TMA store;
  * each level's 3x3 convolution writes the pyramid level in fp32 NHWC (what the bilinear gather and the
    `linear_sdfin` projection read) -- returned as logical-NCHW views, so callers see upstream's shapes.

The parameters stay in the `Decoder` / `Decoder_big` modules (upstream names, strict checkpoint loading); this file
only packs them (cached until a parameter changes) and launches kernels.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn as nn

from .. import ops

TAPS_3X3 = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
TAPS_1X1 = [(0, 0)]
# ConvTranspose2d(k=4, s=2, p=1): output row 2y + py reads input rows y + d through kernel row ky
_DECONV_TAPS = {0: [(1, 0), (3, -1)], 1: [(0, 1), (2, 0)]}     # parity -> [(k index, input offset)]


class Pyramid(dict):
    """Feature pyramid {name: logical-NCHW fp32 tensor}; `split[name]` holds the same level as NHWC rows in split-half
    format when the producer had it anyway (saves the consumer a conversion pass)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.split = {}


class SplitMap:
    """An NHWC feature map in split-half format: `rows` holds b*h*w pixel rows of (at least) c channels."""

    def __init__(self, rows: "ops.SplitRows", b: int, h: int, w: int, c: int):
        assert rows.rows == b * h * w and rows.cols >= c
        self.rows, self.b, self.h, self.w, self.c = rows, b, h, w, c

    def nchw(self) -> torch.Tensor:
        """fp32 copy as a logical-NCHW view of NHWC memory."""
        win = self.rows if self.rows.cols == self.c else ops.SplitRows(self.rows.buf, self.c, self.rows.col0)
        return win.float().view(self.b, self.h, self.w, self.c).permute(0, 3, 1, 2)


def _fold_bn(weight: torch.Tensor, bias, bn: nn.BatchNorm2d, out_dim: int):
    """conv/deconv weight with BatchNorm(eval) folded in: returns (scale-multiplied weight, bias)."""
    w = weight.detach().double()
    cout = w.shape[out_dim]
    b = torch.zeros(cout, dtype=torch.float64, device=w.device) if bias is None else bias.detach().double()
    if bn is not None:
        scale = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
        shape = [1] * w.dim()
        shape[out_dim] = cout
        w = w * scale.view(shape)
        b = (b - bn.running_mean.detach().double()) * scale + bn.bias.detach().double()
    return w, b.float()


def _pack_conv(conv: nn.Conv2d, bn, chunk_kb: int = 0) -> ops.PackedLinearH3:
    """(Cout, Cin, kh, kw) -> planes of (Cout, kh*kw*Cin), K index = tap * Cin + channel."""
    w, b = _fold_bn(conv.weight, conv.bias, bn, 0)
    cout = w.shape[0]
    mat = w.permute(0, 2, 3, 1).reshape(cout, -1).float().contiguous()
    return ops.PackedLinearH3.pack(mat, b, chunk_kb=chunk_kb)


def _pack_deconv(deconv: nn.ConvTranspose2d, bn) -> Dict[Tuple[int, int], Tuple[ops.PackedLinearH3, list]]:
    """(Cin, Cout, 4, 4) -> one (Cout, 4*Cin) matrix + tap list per output parity class."""
    w, b = _fold_bn(deconv.weight, deconv.bias, bn, 1)
    out = {}
    for py in (0, 1):
        for px in (0, 1):
            taps, mats = [], []
            for ky, dy in _DECONV_TAPS[py]:
                for kx, dx in _DECONV_TAPS[px]:
                    taps.append((dy, dx))
                    mats.append(w[:, :, ky, kx].t())            # (Cout, Cin)
            mat = torch.stack(mats, 1).reshape(w.shape[1], -1).float().contiguous()
            out[(py, px)] = (ops.PackedLinearH3.pack(mat, b), taps)
    return out


def _seq_conv_bn(seq: nn.Sequential) -> List[Tuple[nn.Conv2d, object, bool]]:
    """[(conv, bn or None, relu?)] of a Conv(-BN-ReLU)* stack."""
    mods, out, i = list(seq), [], 0
    while i < len(mods):
        conv = mods[i]
        assert isinstance(conv, (nn.Conv2d, nn.ConvTranspose2d))
        bn = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm2d) else None
        relu = bn is not None and i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
        out.append((conv, bn, relu))
        i += 3 if bn is not None else 1
    return out


class UNetH3:
    """Runs a `Decoder` / `Decoder_big` module's forward on the FP16x3 kernels."""

    def __init__(self, decoder: nn.Module):
        self.dec = decoder
        self.big = not hasattr(decoder, "conv0d")
        self._packed = None
        self._key = None

    def _pack(self):
        key = tuple((p.data_ptr(), p._version) for p in list(self.dec.parameters()) + list(self.dec.buffers()))
        if self._packed is not None and self._key == key:
            return self._packed
        d, pk = self.dec, {}
        for i in (1, 2, 3, 4):
            (dc, dbn, _), = _seq_conv_bn(getattr(d, "deconv%d" % i))
            pk["deconv%d" % i] = _pack_deconv(dc, dbn)
            (cv, cbn, _), = _seq_conv_bn(getattr(d, "conv%d" % i))
            pk["conv%d" % i] = _pack_conv(cv, cbn)
            if not self.big:
                (sv, sbn, _), = _seq_conv_bn(getattr(d, "conv%dd" % i))
                pk["conv%dd" % i] = _pack_conv(sv, sbn)
        if not self.big:
            (sv, sbn, _), = _seq_conv_bn(d.conv0d)
            pk["conv0d"] = _pack_conv(sv, sbn)
        firsts = []
        for name in self.HEADS:
            layers = _seq_conv_bn(getattr(d, name))
            assert len(layers) >= 2 and layers[0][2]
            firsts.append(_fold_bn(layers[0][0].weight, layers[0][0].bias, layers[0][1], 0))
            pk[name] = [(_pack_conv(cv, bn), relu) for cv, bn, relu in layers[1:-1]]
            # the last 1x1 convolution has ONE output channel: fp32 weights for the narrow (HBM-bound) Linear kernel
            w, b = _fold_bn(layers[-1][0].weight, layers[-1][0].bias, layers[-1][1], 0)
            pk[name + ".last"] = (w.reshape(w.shape[0], -1).float().contiguous(), b)
        # the three heads' first 1x1 convolutions read the same 524 288-pixel map: ONE GEMM with their weights stacked
        pk["heads.first"] = ops.PackedLinearH3.pack(
            torch.cat([w.reshape(w.shape[0], -1).float() for w, _ in firsts], 0).contiguous(),
            torch.cat([b for _, b in firsts], 0).contiguous())
        pk["heads.width"] = firsts[0][0].shape[0]
        self._packed, self._key = pk, key
        return pk

    LEVELS = ((1, "stride16"), (2, "stride8"), (3, "stride4"), (4, "stride2"))
    HEADS = ("convOut_hm", "convOut_hand_seg", "convOut_obj_seg")

    def concat_slots(self, b: int, h: int, w: int, dev):
        """'ho3d' decoder: allocate the four skip-concat buffers up front and return (cats, slots): the encoder writes
        its stage outputs straight into `slots[name]` = the left channel window of `cats[name]`.  (h, w) = spatial
        size of the stride-32 map.  The small decoder passes its skips through a 1x1 convolution first: no slots."""
        if not self.big:
            return {}, {}
        pk = self._pack()
        cats, slots = {}, {}
        for i, name in self.LEVELS:
            h, w = 2 * h, 2 * w
            cu = pk["deconv%d" % i][(0, 0)][0].n
            cs = pk["conv%d" % i].k // 9 - cu
            cats[name] = ops.SplitRows.empty(b * h * w, cs + cu, dev)
            slots[name] = cats[name].window(0, cs)
        return cats, slots

    def __call__(self, img_feat: torch.Tensor, skips: Dict[str, torch.Tensor]):
        """img_feat (B, 2048, 8, 8), skips {stride2..stride16} (logical NCHW fp32) -> (feature pyramid dict of
        logical-NCHW fp32 tensors backed by NHWC memory, decoder_out (B, 3, 128, 128))."""
        dev = img_feat.device
        b, c0, h, w = img_feat.shape
        x = SplitMap(ops.nchw_to_split(img_feat, ops.SplitRows.empty(b * h * w, c0, dev)), b, h, w, c0)
        cats, slots = self.concat_slots(b, h, w, dev)
        smaps = {}
        for _, name in self.LEVELS:
            skip = skips[name]
            cs, ho, wo = skip.shape[1], skip.shape[2], skip.shape[3]
            dst = slots[name] if self.big else ops.SplitRows.empty(b * ho * wo, cs, dev)
            assert dst.rows == b * ho * wo and dst.cols == cs
            smaps[name] = SplitMap(ops.nchw_to_split(skip, dst), b, ho, wo, cs)
        return self.run(x, smaps, cats, img_feat)

    def run(self, feat: SplitMap, skips: Dict[str, SplitMap], cats: Dict[str, "ops.SplitRows"],
            img_feat: torch.Tensor = None):
        """The decoder on split-half inputs (what nets/resnet_h3.py produces).  For the 'ho3d' decoder `cats` are
        the buffers of `concat_slots`, whose left windows already hold the skips."""
        pk = self._pack()
        x, b, h, w, cx = feat.rows, feat.b, feat.h, feat.w, feat.c
        dev = x.buf.device
        pyr = Pyramid()
        if self.big:
            pyr["stride32"] = img_feat if img_feat is not None else feat.nchw()
            pyr.split["stride32"] = x if x.cols == cx else ops.SplitRows(x.buf, cx, x.col0)
        else:
            p0 = pk["conv0d"]
            lvl = torch.empty(b, h, w, p0.n, device=dev, dtype=torch.float32)
            ops.linear_h3(x, p0, ops.ACT_RELU, out=lvl.view(-1, p0.n))
            pyr["stride32"] = lvl.permute(0, 3, 1, 2)
        for i, name in self.LEVELS:
            skip = skips[name]
            parts = pk["deconv%d" % i]
            cu = parts[(0, 0)][0].n
            ho, wo = 2 * h, 2 * w
            assert skip.h == ho and skip.w == wo and skip.b == b
            if self.big:
                cs = skip.c
                cat = cats[name]
                assert cat.rows == b * ho * wo and cat.cols == cs + cu
            else:       # 1x1 conv + BN + ReLU on the skip, written straight into the concat buffer
                sp = pk["conv%dd" % i]
                cs = sp.n
                cat = ops.SplitRows.empty(b * ho * wo, cs + cu, dev)
                ops.linear_h3(skip.rows, sp, ops.ACT_RELU, out=cat.window(0, cs))
            up = cat.window(cs, cu)
            pitch = cat.ld
            for (py, px), (pw, taps) in parts.items():
                ops.conv_h3(x, b, h, w, cx, pw, taps, h, w, act=ops.ACT_RELU, out=up,
                            out_strides=(2 * pitch, 2 * wo * pitch, ho * wo * pitch),
                            out_offset=(py * wo + px) * pitch)
            pc = pk["conv%d" % i]
            lvl = torch.empty(b, ho, wo, pc.n, device=dev, dtype=torch.float32)
            ops.conv_h3(cat, b, ho, wo, cs + cu, pc, TAPS_3X3, ho, wo, act=ops.ACT_RELU, out=lvl.view(-1, pc.n))
            pyr[name] = lvl.permute(0, 3, 1, 2)
            x = ops.split_rows(lvl.view(-1, pc.n))
            pyr.split[name] = x
            h, w, cx = ho, wo, pc.n
        # 1x1 heads on the stride-2 level: heat map, hand / object segmentation (sigmoid)
        outs = []
        h1 = ops.linear_h3(x, pk["heads.first"], ops.ACT_RELU, split_out=True)
        n1 = pk["heads.width"]
        for j, name in enumerate(self.HEADS):
            hcur = h1.window(j * n1, n1)
            for pw, relu in pk[name]:
                hcur = ops.linear_h3(hcur, pw, ops.ACT_RELU if relu else ops.ACT_NONE, split_out=True)
            wl, bl = pk[name + ".last"]
            act = ops.ACT_NONE if name == "convOut_hm" else ops.ACT_SIGMOID      # upstream module.py:211,215
            outs.append(ops.linear_narrow(hcur, wl, bl, act).reshape(b, 1, h, w))
        return pyr, torch.cat(outs, 1)
