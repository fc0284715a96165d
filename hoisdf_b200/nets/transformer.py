"""`Transformer` / `VoteTransformer` with the upstream constructor, forward signature and state-dict keys
(common/nets/transformer.py:15-459), running on the hoisdf_b200 kernels: fused QKV projection, streaming
softmax attention, out-projection with fused residual, fused (add +) LayerNorm (+ the shared inter_norm).

Activations are kept BATCH-MAJOR (B, S, d) internally so that one sample's tokens are contiguous; the public
`forward` accepts and returns the upstream sequence-major (S, B, d) layout.  Post-norm only
(`normalize_before=False`, upstream main/config.py:122), ReLU FFN, eval mode (dropout off).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from .layer import _require_inference


class MultiheadAttentionParams(nn.Module):
    """Parameter container with nn.MultiheadAttention's names: in_proj_weight [q;k;v], in_proj_bias, out_proj."""

    def __init__(self, d_model, nhead):
        super().__init__()
        if d_model // nhead != 64 or d_model % nhead:
            raise NotImplementedError("hoisdf_b200 attention kernels are built for head_dim 64")
        self.embed_dim = d_model
        self.num_heads = nhead
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d_model, d_model))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d_model))
        self.out_proj = nn.Linear(d_model, d_model)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)
        self._packed = None
        self._key = None

    def packed(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is None or self._key != key:
            d = self.embed_dim
            full = ops.PackedLinear.pack(self.in_proj_weight, self.in_proj_bias)
            self._packed = {
                "qkv": full, "q": full.rows(0, d), "k": full.rows(d, 2 * d), "v": full.rows(2 * d, 3 * d),
                "qk": full.rows(0, 2 * d), "kv": full.rows(d, 3 * d),
                "out": ops.PackedLinear.pack(self.out_proj.weight, self.out_proj.bias),
            }
            self._key = key
        return self._packed


class _LayerBase(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward, dropout, activation, normalize_before):
        super().__init__()
        if normalize_before:
            raise NotImplementedError("pre-norm transformer layers are not used upstream (config.py:122)")
        if activation != "relu":
            raise NotImplementedError("only the ReLU feed-forward of upstream config is built")
        self.d_model = d_model
        self.nhead = nhead
        self.normalize_before = normalize_before
        self._ffn = None
        self._ffn_key = None

    def ffn_packed(self):
        ps = (self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias)
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._ffn is None or self._ffn_key != key:
            self._ffn = (ops.PackedLinear.pack(self.linear1.weight, self.linear1.bias),
                         ops.PackedLinear.pack(self.linear2.weight, self.linear2.bias))
            self._ffn_key = key
        return self._ffn

    def _ffn_block(self, x2d, norm, norm2=None, out2=None, xs=None, out_split=None, out2_split=None):
        """FFN + residual + LayerNorm (+ the shared second norm).  The residual is added by the LayerNorm kernel
        (coalesced row reads) rather than in the GEMM epilogue; `xs` = split-half copy of x2d when the caller has
        one, `out_split` / `out2_split` receive split-half copies of the results (FP16x3 path)."""
        l1, l2 = self.ffn_packed()
        h = ops.linear(xs if xs is not None else x2d, l1, ops.ACT_RELU, split_out=ops.use_h3())
        y = ops.linear(h, l2, ops.ACT_NONE)
        if norm2 is None:
            return ops.add_layernorm(y, x2d, norm.weight, norm.bias, out_split=out_split)
        return ops.add_layernorm(y, x2d, norm.weight, norm.bias, gamma2=norm2.weight, beta2=norm2.bias, out2=out2,
                                 out_split=out_split, out2_split=out2_split)


def _split_like(x2d):
    """Split-half buffer for a (rows, d) activation on the FP16x3 path, else None."""
    return ops.SplitRows.empty(x2d.shape[0], x2d.shape[1], x2d.device) if ops.use_h3() else None


class TransformerEncoderLayer(_LayerBase):
    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False):
        super().__init__(d_model, nhead, dim_feedforward, dropout, activation, normalize_before)
        self.self_attn = MultiheadAttentionParams(d_model, nhead)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)

    def forward_bm(self, x: torch.Tensor, pos: Optional[torch.Tensor], inter_norm=None, inter_out=None, xs=None,
                   out_split=None, inter_split=None):
        """x (B, S, d) batch-major contiguous -> same shape; upstream transformer.py:279-302 (forward_post).
        xs / out_split / inter_split: split-half copies of x / the result / the inter_norm output (FP16x3 path)."""
        b, s, d = x.shape
        x2d = x.view(b * s, d)
        pk = self.self_attn.packed()
        if pos is None:
            qkv = ops.linear(xs if xs is not None else x2d, pk["qkv"])   # (rows, 3d): q | k | v
            q, k, v, ld = qkv, qkv[:, d:], qkv[:, 2 * d:], 3 * d
            ldq = ld
        else:
            qk_in = (x + pos).view(b * s, d)
            qk = ops.linear(qk_in, pk["qk"])
            vv = ops.linear(xs if xs is not None else x2d, pk["v"])
            # the attention entry point takes one pitch for k and v: copy v next to k
            kv = torch.empty(b * s, 2 * d, device=x.device, dtype=torch.float32)
            kv[:, :d] = qk[:, d:]
            kv[:, d:] = vv
            q, ldq, k, v, ld = qk, 2 * d, kv, kv[:, d:], 2 * d
        # (tensor-core attention writes its result in the split-half format the out-projection GEMM reads)
        att = ops.SplitRows.empty(b * s, d, x.device) if (ops.use_h3() and s > 32) else \
            torch.empty(b * s, d, device=x.device, dtype=torch.float32)
        ops.attention(q, ldq, k, v, ld, att, d, b, self.nhead, s, s)
        y = ops.linear(att, pk["out"], ops.ACT_NONE)
        x1s = _split_like(x2d)
        x1 = ops.add_layernorm(y, x2d, self.norm1.weight, self.norm1.bias, out_split=x1s)
        out = self._ffn_block(x1, self.norm2, inter_norm, inter_out, x1s, out_split, inter_split)
        return out.view(b, s, d)


class TransformerDecoderLayer(_LayerBase):
    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False):
        super().__init__(d_model, nhead, dim_feedforward, dropout, activation, normalize_before)
        self.self_attn = MultiheadAttentionParams(d_model, nhead)
        self.multihead_attn = MultiheadAttentionParams(d_model, nhead)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)

    def forward_bm(self, tgt, memory, pos, query_pos, tgt_mask, memory_mask, final_norm, final_out,
                   memory_kv_valid=None, memory_split=None):
        """tgt (B, Lq, d), memory (B, S, d), query_pos (B, Lq, d) batch-major; masks uint8 (1 = blocked).
        upstream transformer.py:366-395 (forward_post)."""
        b, lq, d = tgt.shape
        s = memory.shape[1]
        dev = tgt.device
        t2d = tgt.view(b * lq, d)
        # masked self-attention over the queries
        sa = self.self_attn.packed()
        qk_in = (tgt + query_pos).view(b * lq, d)
        qk = ops.linear(qk_in, sa["qk"])
        kv = torch.empty(b * lq, 2 * d, device=dev, dtype=torch.float32)
        ops.linear(t2d, sa["v"], out=kv[:, d:])
        kv[:, :d] = qk[:, d:]
        att = torch.empty(b * lq, d, device=dev, dtype=torch.float32)
        ops.attention(qk, 2 * d, kv, kv[:, d:], 2 * d, att, d, b, self.nhead, lq, lq, mask=tgt_mask)
        y = ops.linear(att, sa["out"], ops.ACT_NONE)
        t1 = ops.add_layernorm(y, t2d, self.norm1.weight, self.norm1.bias)
        # cross-attention to the encoder memory
        ca = self.multihead_attn.packed()
        q = ops.linear((t1.view(b, lq, d) + query_pos).view(b * lq, d), ca["q"])
        m2d = memory.view(b * s, d)
        if pos is None:
            mkv = ops.linear(memory_split if memory_split is not None else m2d, ca["kv"])   # (B*S, 2d): k | v
        else:
            mkv = torch.empty(b * s, 2 * d, device=dev, dtype=torch.float32)
            ops.linear((memory + pos).view(b * s, d), ca["k"], out=mkv[:, :d])
            ops.linear(memory_split if memory_split is not None else m2d, ca["v"], out=mkv[:, d:])
        att2 = torch.empty(b * lq, d, device=dev, dtype=torch.float32)
        if memory_kv_valid is not None:
            # memory_mask == "keys >= memory_kv_valid are blocked for every query" (common/utils/misc.py:42-47):
            # expressed as a key-range limit, which the tensor-core kernel supports (a dense mask does not)
            ops.attention(q, d, mkv, mkv[:, d:], 2 * d, att2, d, b, self.nhead, lq, s, kv_valid=memory_kv_valid,
                          tensor_cores=True)
        else:
            ops.attention(q, d, mkv, mkv[:, d:], 2 * d, att2, d, b, self.nhead, lq, s, mask=memory_mask)
        y2 = ops.linear(att2, ca["out"], ops.ACT_NONE)
        t2s = _split_like(t1)
        t2 = ops.add_layernorm(y2, t1, self.norm2.weight, self.norm2.bias, out_split=t2s)
        out = self._ffn_block(t2, self.norm3, final_norm, final_out, t2s)
        return out.view(b, lq, d)


def _clones(make, n):
    return nn.ModuleList([make() for _ in range(n)])


class TransformerEncoder(nn.Module):
    def __init__(self, make_layer, num_layers, norm=None, inter_norm=None, return_intermediate=False):
        super().__init__()
        self.layers = _clones(make_layer, num_layers)
        self.num_layers = num_layers
        self.norm = norm
        self.inter_norm = inter_norm
        self.return_intermediate = return_intermediate

    def forward_bm(self, x, pos):
        """(B,S,d) -> (last (B,S,d), intermediate (L,B,S,d) = inter_norm of every layer output);
        upstream transformer.py:175-202."""
        b, s, d = x.shape
        n = b * s
        from ..config import cfg
        if (pos is None and ops.use_h3() and s > 32 and cfg.native_encoder and self.norm is None and ops.PROFILE is None
                and (not self.return_intermediate or self.inter_norm is not None)):
            res = self._forward_native(x)
            if res is not None:
                return res
        inter = torch.empty(self.num_layers, b, s, d, device=x.device, dtype=torch.float32) \
            if self.return_intermediate else None
        # FP16x3 path: every layer also emits its output (and the inter_norm output) in split-half format from the
        # LayerNorm kernel, so neither the next layer's projections nor the heads need a separate split pass
        h3 = ops.use_h3()
        inter_split = ops.SplitRows.empty(self.num_layers * n, d, x.device) if (h3 and inter is not None) else None
        xs = None
        for i, layer in enumerate(self.layers):
            out_split = ops.SplitRows.empty(n, d, x.device) if h3 else None
            if inter is not None:
                isp = ops.SplitRows(inter_split.buf[i * n:(i + 1) * n], d) if inter_split is not None else None
                x = layer.forward_bm(x, pos, self.inter_norm, inter[i], xs, out_split, isp)
            else:
                x = layer.forward_bm(x, pos, None, None, xs, out_split)
            xs = out_split
        self.last_out_split, self.last_inter_split = xs, inter_split
        return x, inter


    def _forward_native(self, x):
        """The whole stack through ONE C call (csrc/transformer.cu: hoisdf_encoder_fwd) with a caller-owned workspace;
        same kernels, same order as the per-layer Python path below (bit-identical).  None = not applicable."""
        import ctypes as C
        from .. import _capi
        b, s, d = x.shape
        n, dev = b * s, x.device
        L = self.num_layers
        layers = (_capi.EncoderLayer * L)()
        keep = []
        d_ff = None

        def fill(dst, pw):
            h3 = pw.h3
            if h3 is None:
                return False
            dst.a, dst.b, dst.c, dst.ld = h3.plane_ptr(0), h3.plane_ptr(1), h3.plane_ptr(2), h3.ld
            dst.bias, dst.scale = ops._ptr(h3.b), float(h3.scale)
            keep.append(h3)
            return True

        for i, layer in enumerate(self.layers):
            pk = layer.self_attn.packed()
            l1, l2 = layer.ffn_packed()
            if not (fill(layers[i].qkv, pk["qkv"]) and fill(layers[i].out, pk["out"]) and fill(layers[i].lin1, l1)
                    and fill(layers[i].lin2, l2)):
                return None
            d_ff = l1.n if d_ff is None else d_ff
            if l1.n != d_ff or layer.nhead * 64 != d:
                return None
            layers[i].norm1_g, layers[i].norm1_b = layer.norm1.weight.data_ptr(), layer.norm1.bias.data_ptr()
            layers[i].norm2_g, layers[i].norm2_b = layer.norm2.weight.data_ptr(), layer.norm2.bias.data_ptr()
        heads = self.layers[0].nhead
        nbytes = int(_capi.lib.hoisdf_encoder_workspace_bytes(b, s, d_ff, heads))
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        out = torch.empty(b, s, d, device=dev, dtype=torch.float32)
        out_split = ops.SplitRows.empty(n, d, dev)
        inter = inter_split = None
        a = _capi.EncoderArgs()
        a.layers, a.num_layers, a.heads, a.d_ff = layers, L, heads, d_ff
        a.batch, a.seq, a.x = b, s, x.data_ptr()
        a.out, a.out_hi, a.out_lo, a.ld_out = out.data_ptr(), out_split.hi_ptr, out_split.lo_ptr, out_split.ld
        if self.return_intermediate:
            inter = torch.empty(L, b, s, d, device=dev, dtype=torch.float32)
            inter_split = ops.SplitRows.empty(L * n, d, dev)
            a.inter_g, a.inter_b = self.inter_norm.weight.data_ptr(), self.inter_norm.bias.data_ptr()
            a.inter, a.inter_hi, a.inter_lo, a.ld_inter = inter.data_ptr(), inter_split.hi_ptr, inter_split.lo_ptr, \
                inter_split.ld
        a.workspace, a.workspace_bytes = ws.data_ptr(), nbytes
        ops._count(L * 12 + 1)
        st = _capi.lib.hoisdf_encoder_fwd(C.byref(a), ops._stream())
        if st == _capi.E_UNSUPPORTED:
            return None
        _capi.check(st, "hoisdf_encoder_fwd")
        del keep
        self.last_out_split, self.last_inter_split = out_split, inter_split
        return out, inter


class TransformerDecoder(nn.Module):
    def __init__(self, make_layer, num_layers, norm=None, return_intermediate=False):
        super().__init__()
        self.layers = _clones(make_layer, num_layers)
        self.num_layers = num_layers
        self.norm = norm
        self.return_intermediate = return_intermediate
        self._tgt_is_zero = False          # set by Transformer.forward_bm, which builds tgt = zeros itself

    def forward_bm(self, tgt, memory, pos, query_pos, tgt_mask, memory_mask, memory_kv_valid=None,
                   memory_split=None):
        """-> hs (L, B, Lq, d) = norm(out_l) for every layer (upstream transformer.py:214-252)."""
        b, lq, d = tgt.shape
        hs = torch.empty(self.num_layers, b, lq, d, device=tgt.device, dtype=torch.float32)
        from ..config import cfg
        if (pos is None and memory_kv_valid is not None and memory_split is not None and ops.use_h3() and cfg.native_decoder
                and ops.PROFILE is None and self.norm is not None):
            # (the C entry starts from tgt = 0, which is what Transformer.forward_bm passes -- upstream transformer.py:150)
            if self._tgt_is_zero and self._forward_native(hs, memory, query_pos, tgt_mask, memory_kv_valid, memory_split):
                return hs
        x = tgt
        for i, layer in enumerate(self.layers):
            x = layer.forward_bm(x, memory, pos, query_pos, tgt_mask, memory_mask, self.norm, hs[i], memory_kv_valid,
                                 memory_split)
        return hs


def _fill_h3(dst, pw, keep):
    h3 = pw.h3
    if h3 is None:
        return False
    dst.a, dst.b, dst.c, dst.ld = h3.plane_ptr(0), h3.plane_ptr(1), h3.plane_ptr(2), h3.ld
    dst.bias, dst.scale = ops._ptr(h3.b), float(h3.scale)
    keep.append(h3)
    return True


def _decoder_forward_native(self, hs, memory, query_pos, tgt_mask, kv_valid, memory_split):
    """The whole decoder stack through ONE C call (csrc/transformer.cu: hoisdf_decoder_fwd; tgt = 0 as upstream
    transformer.py:150): same kernels, same order as the per-layer Python path (bit-identical).  False = not applicable."""
    import ctypes as C
    from .. import _capi
    L, b, lq, d = hs.shape
    s = memory.shape[1]
    layers = (_capi.DecoderLayer * L)()
    keep, d_ff = [], None
    for i, layer in enumerate(self.layers):
        sa, ca = layer.self_attn.packed(), layer.multihead_attn.packed()
        l1, l2 = layer.ffn_packed()
        ok = (_fill_h3(layers[i].sa_qk, sa["qk"], keep) and _fill_h3(layers[i].sa_v, sa["v"], keep)
              and _fill_h3(layers[i].sa_out, sa["out"], keep) and _fill_h3(layers[i].ca_q, ca["q"], keep)
              and _fill_h3(layers[i].ca_kv, ca["kv"], keep) and _fill_h3(layers[i].ca_out, ca["out"], keep)
              and _fill_h3(layers[i].lin1, l1, keep) and _fill_h3(layers[i].lin2, l2, keep))
        d_ff = l1.n if d_ff is None else d_ff
        if not ok or l1.n != d_ff or layer.nhead * 64 != d:
            return False
        for j, nm in enumerate((layer.norm1, layer.norm2, layer.norm3), 1):
            setattr(layers[i], "norm%d_g" % j, nm.weight.data_ptr())
            setattr(layers[i], "norm%d_b" % j, nm.bias.data_ptr())
    heads = self.layers[0].nhead
    nbytes = int(_capi.lib.hoisdf_decoder_workspace_bytes(b, lq, s, d_ff, heads))
    ws = torch.empty(nbytes, device=hs.device, dtype=torch.uint8)
    qp = query_pos.contiguous()
    a = _capi.DecoderArgs()
    a.layers, a.num_layers, a.heads, a.d_ff = layers, L, heads, d_ff
    a.norm_g, a.norm_b = self.norm.weight.data_ptr(), self.norm.bias.data_ptr()
    a.batch, a.queries, a.seq, a.kv_valid = b, lq, s, int(kv_valid)
    a.query_pos, a.tgt_mask = qp.data_ptr(), ops._ptr(tgt_mask)
    a.memory_hi, a.memory_lo, a.ld_memory = memory_split.hi_ptr, memory_split.lo_ptr, memory_split.ld
    a.hs, a.workspace, a.workspace_bytes = hs.data_ptr(), ws.data_ptr(), nbytes
    ops._count(L * 22)
    _capi.check(_capi.lib.hoisdf_decoder_fwd(C.byref(a), ops._stream()), "hoisdf_decoder_fwd")
    del keep
    return True


TransformerDecoder._forward_native = _decoder_forward_native


def _mask_u8(mask: Optional[torch.Tensor], device):
    if mask is None:
        return None
    return mask.to(device=device, dtype=torch.uint8).contiguous()


def _suffix_mask_limit(mask: Optional[torch.Tensor]):
    """If `mask` (Lq, S) bool blocks exactly the keys [n, S) for every query, return n; else None.  Only evaluated for
    host tensors (the model builds its masks on the CPU, like upstream model.py:568-569), so no device sync."""
    if mask is None or mask.is_cuda or mask.dtype != torch.bool or mask.dim() != 2:
        return None
    col = mask[0]
    if not bool((mask == col).all()):
        return None
    n = int((~col).sum())
    if n == 0 or bool(col[:n].any()) or not bool(col[n:].all()):
        return None
    return n


def _xavier(module):
    for p in module.parameters():
        if p.dim() > 1:
            nn.init.xavier_uniform_(p)


class VoteTransformer(nn.Module):
    """Encoder-only transformer (upstream transformer.py:15-65)."""

    def __init__(self, d_model=512, nhead=8, num_encoder_layers=6, dim_feedforward=2048, dropout=0.1,
                 activation="relu", normalize_before=False, return_intermediate_dec=False):
        super().__init__()
        self.encoder = TransformerEncoder(
            lambda: TransformerEncoderLayer(d_model, nhead, dim_feedforward, dropout, activation, normalize_before),
            num_encoder_layers, None, nn.LayerNorm(d_model), return_intermediate_dec)
        _xavier(self)
        self.d_model = d_model
        self.nhead = nhead

    def forward_bm(self, src_bm, pos_bm=None):
        x = src_bm if pos_bm is None else src_bm + pos_bm
        return self.encoder.forward_bm(x.contiguous(), pos_bm)

    def forward(self, src, mask, pos_embed, src_mask=None):
        """src (S,B,d), pos_embed (S,B,d) -> (memory (S,B,d), intermediate (L,S,B,d))."""
        _require_inference(self, src)
        if mask is not None or src_mask is not None:
            raise NotImplementedError("key-padding / source masks are never passed upstream (model.py:582-584)")
        pos = None if pos_embed is None else pos_embed.transpose(0, 1).contiguous()
        mem, inter = self.forward_bm(src.transpose(0, 1).contiguous(), pos)
        return mem.transpose(0, 1).contiguous(), (inter.transpose(1, 2).contiguous() if inter is not None else [])


class Transformer(nn.Module):
    """Encoder-decoder transformer (upstream transformer.py:68-155)."""

    def __init__(self, d_model=512, nhead=8, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=2048,
                 dropout=0.1, activation="relu", normalize_before=False, return_intermediate_dec=False):
        super().__init__()
        self.encoder = TransformerEncoder(
            lambda: TransformerEncoderLayer(d_model, nhead, dim_feedforward, dropout, activation, normalize_before),
            num_encoder_layers, None, nn.LayerNorm(d_model), return_intermediate_dec)
        self.decoder = TransformerDecoder(
            lambda: TransformerDecoderLayer(d_model, nhead, dim_feedforward, dropout, activation, normalize_before),
            num_decoder_layers, nn.LayerNorm(d_model), return_intermediate=return_intermediate_dec)
        _xavier(self)
        self.d_model = d_model
        self.nhead = nhead

    def forward_bm(self, src_bm, query_embed, pos_bm, tgt_mask, memory_mask):
        """src (B,S,d), query_embed (Lq,d) -> hs (L,B,Lq,d), memory (B,S,d), intermediate (Le,B,S,d).
        A memory_mask of the form "all queries blocked from keys >= n" (the only form upstream builds) is
        recognised on the host and turned into a key-range limit."""
        b = src_bm.shape[0]
        x = src_bm if pos_bm is None else src_bm + pos_bm
        memory, inter = self.encoder.forward_bm(x.contiguous(), pos_bm)
        qpos = query_embed.detach().unsqueeze(0).expand(b, -1, -1).contiguous()
        tgt = torch.zeros_like(qpos)
        dev = src_bm.device
        kv_valid = _suffix_mask_limit(memory_mask)
        self.decoder._tgt_is_zero = True       # tgt was built as zeros right above: the C entry may start from its own zeros
        try:
            hs = self.decoder.forward_bm(tgt, memory, pos_bm, qpos, _mask_u8(tgt_mask, dev),
                                         None if kv_valid is not None else _mask_u8(memory_mask, dev), kv_valid,
                                         self.encoder.last_out_split)
        finally:
            self.decoder._tgt_is_zero = False
        return hs, memory, inter

    def forward(self, src, mask, query_embed, pos_embed, tgt_mask=None, src_mask=None, memory_mask=None):
        """Upstream layout: src/pos_embed (S,B,d) -> (hs (L,Lq,B,d), memory (S,B,d), intermediate (Le,S,B,d), None).

        The fourth return value (per-layer averaged attention maps upstream) is never consumed by the model
        (main/model.py:571) and is returned as None instead of materialising Lq x S maps.
        """
        _require_inference(self, src)
        if mask is not None or src_mask is not None:
            raise NotImplementedError("key-padding / source masks are never passed upstream (model.py:571-581)")
        pos = None if pos_embed is None else pos_embed.transpose(0, 1).contiguous()
        hs, memory, inter = self.forward_bm(src.transpose(0, 1).contiguous(), query_embed, pos, tgt_mask, memory_mask)
        inter_sm = inter.transpose(1, 2).contiguous() if inter is not None else []
        return hs.transpose(1, 2).contiguous(), memory.transpose(0, 1).contiguous(), inter_sm, None
