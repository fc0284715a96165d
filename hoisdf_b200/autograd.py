"""torch.autograd wrappers of the hoisdf_b200 kernels -- what `Model.forward(mode="train")` is assembled from
(upstream main/train.py:104-140 back-propagates through main/model.py:357-665).

Every Function calls the C ABI for the forward AND the backward; PyTorch only provides the tape.

  LinearFn        Y = act(X W^T + b)            forward : FP16x3 tcgen05 GEMM (hoisdf_linear_h3_fwd)
                                                 backward: hoisdf_absmax + hoisdf_linear_bwd_prep (one pass over dY: dZ = dY *
                                                 relu', db, and both GEMM operands in their tensor-core formats), then the
                                                 SAME tensor-core GEMM on transposed operands: dX = dZ . W = linear_h3(dZ,
                                                 pack(W^T)), dW^T = X^T . dZ = linear_h3(X^T, pack(dZ^T)).  Gradients are brought
                                                 into the fp16 planes' range by a power-of-two scale computed ON THE DEVICE
                                                 (exact, undone after the product): no host read-back.  Layers with <= 16 rows,
                                                 inputs or outputs take hoisdf_act_bias_bwd + the fp32 FMA GEMM
  WeightNormFn    W = g v / |v|                 hoisdf_fold_weight_norm / hoisdf_weight_norm_bwd
  GatherFn        5-level bilinear gather       hoisdf_gather_fwd (CONCAT) / hoisdf_gather_bwd (scatter-add into the pyramid grad)
  AddLayerNormFn  LayerNorm(x + res)            hoisdf_add_layernorm_fwd / hoisdf_layernorm_bwd
  AttentionFn     softmax(q k^T / 8 [mask]) v   forward: tcgen05 flash kernel (hoisdf_attention_fwd; hoisdf_attention_dropout_fwd with
                                                 dropout on the probabilities), the materialised form for masked / short
                                                 sequences with dropout; backward: hoisdf_gemm_f32_batched +
                                                 hoisdf_softmax_rows_fwd / _bwd per (sample, head) -- fp32 SIMT, the S x S
                                                 probabilities are recomputed, not stored; dropout on the probabilities through
                                                 hoisdf_softmax_dropout_rows_fwd / _bwd (hashed keep decisions: only a seed is kept)
  TokensFn        token assembly + sdf_activation  hoisdf_tokens_fwd / hoisdf_tokens_bwd (own-field blocks only: upstream detaches
                                                 the SDF values and the cross-field tokens, model.py:483-484,536,555)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch
from torch.autograd import Function

from . import ops
from ._capi import check, lib
from .ops import _count, _ptr, _stream

ACT_NONE, ACT_RELU = ops.ACT_NONE, ops.ACT_RELU
TRAIN_CHUNK_KB = 4          # TMEM accumulation chunk of the training GEMMs (K blocks of 32 per drain)
_DEBUG_TORCH_MATMUL = os.environ.get("HOISDF_DEBUG_TORCH_MATMUL", "0") == "1"
_ATTN_BWD_TC = os.environ.get("HOISDF_ATTN_BWD_TC", "1") != "0"     # 0: batched fp32 FMA attention backward everywhere


# ----------------------------------------------------------------------------------------------------
# tensor-core A . B^T with on-device power-of-two scaling
# ----------------------------------------------------------------------------------------------------
def _pow2_scale(t: torch.Tensor, log2_target: int) -> torch.Tensor:
    """0-dim device tensor s = 2^e with max|t| / s in (2^(log2_target-1), 2^log2_target]; 1 for an all-zero tensor."""
    amax = t.abs().amax()
    e = torch.ceil(torch.log2(amax.clamp_min(1e-30))) - log2_target
    return torch.where(amax > 0, torch.exp2(e), torch.ones_like(e))


def _c2d(t: torch.Tensor) -> torch.Tensor:
    return t if (t.stride(-1) == 1 and t.dim() == 2) else t.contiguous()


def matmul_nt(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
              chunk_kb: int = TRAIN_CHUNK_KB) -> torch.Tensor:
    """act(a (M,K) . b (N,K)^T + bias) -> (M,N) fp32 on the FP16x3 kernel.  `b` is the "weight" operand: |b| < 16 (its hi
    plane is stored times 2^11 in fp16); `a` any fp16-range values.  No host synchronisation.  Products with a handful of
    output columns or a tiny contraction (the N = 1 / 3 / 6 / 10 head layers and their gradients) are not tensor-core
    shapes: they take the fp32 FMA GEMM."""
    a, b = _c2d(a), _c2d(b)
    m, k = a.shape
    n = b.shape[0]
    if _DEBUG_TORCH_MATMUL:        # developer switch (scripts/train_debug.py): isolates the GEMM kernels from the tape logic
        y = a @ b.t()
        y = y if bias is None else y + bias
        return torch.relu(y) if act == ACT_RELU else y
    if n <= 16 and m > 16:
        # a handful of output features over many rows: one streaming pass over `a` (csrc/backward.cu: thin_linear_fwd)
        y = torch.empty(m, n, device=a.device, dtype=torch.float32)
        _count(1)
        check(lib.hoisdf_thin_linear_fwd(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0) if n > 1 else k, _ptr(bias), m, k,
                                         n, act, y.data_ptr(), n, _stream()), "hoisdf_thin_linear_fwd")
        return y
    if n <= 16 or k <= 16:
        y = torch.empty(m, n, device=a.device, dtype=torch.float32)
        _count(1)
        lda = a.stride(0) if m > 1 else k        # a one-row tensor may carry any row stride
        ldb = b.stride(0) if n > 1 else k
        check(lib.hoisdf_gemm_f32(a.data_ptr(), lda, 0, b.data_ptr(), ldb, 1, y.data_ptr(), n, m, n, k, 0,
                                  _stream()), "hoisdf_gemm_f32")
        if bias is not None:
            y += bias
        return torch.relu_(y) if act == ACT_RELU else y
    xs = ops.split_rows(a)
    pw = ops.PackedLinearH3.pack(b, None if bias is None else bias.detach(), assume_max=1.0)
    return ops.linear_h3(xs, pw, act, chunk_kb=chunk_kb)


_lin_ws = {}


def _linear_workspace(device, nbytes: int) -> torch.Tensor:
    """Grow-only scratch buffer per device for hoisdf_linear_train_fwd / _bwd (operand copies in the tensor-core formats; used
    inside one call only, calls are ordered on the stream)."""
    buf = _lin_ws.get(device)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), device=device, dtype=torch.uint8)
        _lin_ws[device] = buf
    return buf


def _fused_linear(m: int, n: int, k: int) -> bool:
    """One C call per direction (hoisdf_linear_train_fwd / _bwd) for the tensor-core shapes; the per-kernel Python path stays
    for bench.py's per-launch profile and the developer switch."""
    return min(m, n, k) > 16 and not _DEBUG_TORCH_MATMUL and ops.PROFILE is None


class LinearFn(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act):
        x = _c2d(x)
        m, k = x.shape
        n = weight.shape[0]
        if _fused_linear(m, n, k):
            w = _c2d(weight.detach())
            y = torch.empty(m, ops.round_up(n, 4), device=x.device, dtype=torch.float32)[:, :n]
            nbytes = lib.hoisdf_linear_train_workspace_bytes(m, n, k)
            ws = _linear_workspace(x.device, nbytes)
            _count(3)
            check(lib.hoisdf_linear_train_fwd(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0),
                                              _ptr(None if bias is None else bias.detach()), m, n, k, act, y.data_ptr(),
                                              y.stride(0), ws.data_ptr(), nbytes, _stream()), "hoisdf_linear_train_fwd")
        else:
            y = matmul_nt(x, weight, bias, act)
        ctx.act = act
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight, y if act == ACT_RELU else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        m, n = dy.shape
        k = x.shape[1]
        if min(m, n, k) <= 16 or _DEBUG_TORCH_MATMUL:
            return LinearFn._backward_small(ctx, dy, x, weight, y)
        if _fused_linear(m, n, k):
            dy = _c2d(dy)
            dev = dy.device
            w = _c2d(weight.detach())
            want_dx, want_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
            dx = torch.empty(m, ops.round_up(k, 4), device=dev, dtype=torch.float32)[:, :k] if want_dx else None
            dwt = torch.empty(k, ops.round_up(n, 4), device=dev, dtype=torch.float32)[:, :n] if want_dw else None
            db = torch.empty(n, device=dev, dtype=torch.float32) if ctx.has_bias else None
            nbytes = lib.hoisdf_linear_train_workspace_bytes(m, n, k)
            ws = _linear_workspace(dev, nbytes)
            _count(2 + 2 * int(want_dx) + 2 * int(want_dw))
            check(lib.hoisdf_linear_train_bwd(dy.data_ptr(), dy.stride(0), _ptr(y), 0 if y is None else y.stride(0),
                                              x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), m, n, k, ctx.act,
                                              _ptr(dx), 0 if dx is None else dx.stride(0), _ptr(dwt),
                                              0 if dwt is None else dwt.stride(0), _ptr(db), ws.data_ptr(), nbytes, _stream()),
                  "hoisdf_linear_train_bwd")
            return dx, (None if dwt is None else dwt.t()), db, None
        # tensor-core backward: operands prepared in two passes over dY (csrc/train_prep.cu), no host read-back
        dy = _c2d(dy)
        dev = dy.device
        amax = torch.empty(1, device=dev, dtype=torch.float32)
        scale = torch.empty(1, device=dev, dtype=torch.float32)
        dz = ops.SplitRows.empty(m, n, dev)
        ldt = ops.round_up(m, 8)
        dzt = torch.empty(3, n, ldt, device=dev, dtype=torch.float16)
        db = torch.empty(n, device=dev, dtype=torch.float32) if ctx.has_bias else None
        _count(2)
        check(lib.hoisdf_absmax(dy.data_ptr(), m, n, dy.stride(0), amax.data_ptr(), _stream()), "hoisdf_absmax")
        check(lib.hoisdf_linear_bwd_prep(dy.data_ptr(), dy.stride(0), _ptr(y), 0 if y is None else y.stride(0), m, n, ctx.act,
                                         amax.data_ptr(), dz.hi_ptr, dz.lo_ptr, dz.ld, dzt[0].data_ptr(), dzt[1].data_ptr(),
                                         dzt[2].data_ptr(), ldt, _ptr(db), scale.data_ptr(), _stream()),
              "hoisdf_linear_bwd_prep")
        dx = dw = None
        if ctx.needs_input_grad[0]:
            pwt = ops.PackedLinearH3.pack(weight.detach().t().contiguous(), None, assume_max=1.0)      # W^T (K, N)
            dx = ops.linear_h3(dz, pwt, ACT_NONE, chunk_kb=TRAIN_CHUNK_KB, y_scale=scale)                # (dZ/s . W) * s
        if ctx.needs_input_grad[1]:
            xt = ops.SplitRows.empty(k, m, dev)
            _count(1)
            check(lib.hoisdf_split_rows_t(x.data_ptr(), m, k, x.stride(0) if m > 1 else max(x.stride(0), k), xt.hi_ptr,
                                          xt.lo_ptr, xt.ld, _stream()), "hoisdf_split_rows_t")
            pdz = ops.PackedLinearH3(dzt, None, n, m)
            dw = ops.linear_h3(xt, pdz, ACT_NONE, chunk_kb=TRAIN_CHUNK_KB, y_scale=scale,
                               split_k=True).t()                                              # (X^T . dZ/s) * s = dW^T
        return dx, dw, db, None

    @staticmethod
    def _backward_small(ctx, dy, x, weight, y):
        """Heads with a handful of outputs (N = 1 / 3 / 6 / 10), tiny batches: torch glue + the fp32 FMA GEMM."""
        m, n = dy.shape
        dz = dy.contiguous().clone() if ctx.act == ACT_RELU else dy.contiguous()
        db = torch.empty(n, device=dy.device, dtype=torch.float32) if ctx.has_bias else None
        if ctx.act == ACT_RELU or db is not None:
            _count(1)
            check(lib.hoisdf_act_bias_bwd(dz.data_ptr(), dz.stride(0), _ptr(y), 0 if y is None else y.stride(0), m, n,
                                          ctx.act, _ptr(db), 0, _stream()), "hoisdf_act_bias_bwd")
        s = _pow2_scale(dz, 3)                       # |dz / s| <= 8: fits both operand formats
        dzs = dz * (1.0 / s)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = matmul_nt(dzs, weight.detach().t().contiguous())            # (M,N) . (K,N)^T
            dx = dx.mul_(s) if dx.is_contiguous() else dx * s
        if ctx.needs_input_grad[1]:
            if n <= 16 and m > 16:
                # dW = dZ^T X as one pass over X: no transposed copies, no 64-wide tiles for 1 / 3 / 6 / 10 output features
                dw = torch.empty(n, x.shape[1], device=dy.device, dtype=torch.float32)
                _count(1)
                check(lib.hoisdf_thin_linear_dw(x.data_ptr(), x.stride(0), dz.data_ptr(), dz.stride(0), m, x.shape[1], n,
                                                dw.data_ptr(), dw.stride(0), 0, _stream()), "hoisdf_thin_linear_dw")
            else:
                dwt = matmul_nt(x.t().contiguous(), dzs.t().contiguous())    # (K,M) . (N,M)^T = dW^T / s
                dw = dwt.t() * s
        return dx, dw, db, None


def linear(x2d: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], act: int = ACT_NONE) -> torch.Tensor:
    return LinearFn.apply(x2d, weight, bias, act)


class WeightNormFn(Function):
    """nn.utils.weight_norm (dim 0): W[r] = g[r] * v[r] / |v[r]| (upstream common/nets/sdf_net.py:57-62)."""

    @staticmethod
    def forward(ctx, g, v):
        w = ops.fold_weight_norm(g, v)[:, : v.shape[1]]
        ctx.save_for_backward(g, v)
        return w

    @staticmethod
    def backward(ctx, dw):
        g, v = ctx.saved_tensors
        dw = _c2d(dw)
        rows, cols = v.shape
        dg = torch.empty(rows, device=v.device, dtype=torch.float32)
        dv = torch.empty(rows, cols, device=v.device, dtype=torch.float32)
        vc, gc = v.detach().contiguous(), g.detach().reshape(-1).contiguous()
        _count(1)
        check(lib.hoisdf_weight_norm_bwd(gc.data_ptr(), vc.data_ptr(), dw.data_ptr(), dw.stride(0), rows, cols,
                                         dg.data_ptr(), dv.data_ptr(), 0, _stream()), "hoisdf_weight_norm_bwd")
        return dg.view_as(g), dv


class GatherFn(Function):
    """CONCAT-mode bilinear gather of the 5 NHWC pyramid levels at `uv` (rows, 2) pixels (upstream F.grid_sample calls,
    main/model.py:166-171,206-211; the grid is detached there, so `uv` gets no gradient)."""

    @staticmethod
    def forward(ctx, uv, batch, rows_per_sample, img_hw, *maps):
        maps = [m.contiguous() for m in maps]
        rows = uv.shape[0]
        out = torch.empty(rows, sum(m.shape[3] for m in maps), device=uv.device, dtype=torch.float32)
        ops.gather(maps, uv, batch, mode=ops.GATHER_CONCAT, out=out, rows_per_sample=rows_per_sample, img_hw=img_hw)
        ctx.save_for_backward(uv)
        ctx.geom = (batch, rows_per_sample, img_hw, [tuple(m.shape) for m in maps])
        return out

    @staticmethod
    def backward(ctx, dout):
        (uv,) = ctx.saved_tensors
        batch, rps, img_hw, shapes = ctx.geom
        dout = _c2d(dout)
        grads = [torch.zeros(s, device=dout.device, dtype=torch.float32) for s in shapes]
        pyr = ops.make_pyramid(grads, img_hw)
        _count(1)
        check(lib.hoisdf_gather_bwd(C.byref(pyr), uv.data_ptr(), uv.shape[0], None, batch, rps, dout.data_ptr(),
                                    dout.stride(0), _stream()), "hoisdf_gather_bwd")
        return (None, None, None, None, *grads)


class AddLayerNormFn(Function):
    """LayerNorm(x + res) over the last dim (256), eps 1e-5 (upstream transformer.py:296-301)."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta):
        h = x if res is None else x + res
        h = h.contiguous()
        y = ops.add_layernorm(h, None, gamma.detach(), beta.detach())
        ctx.save_for_backward(h, gamma)
        ctx.has_res = res is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        h, gamma = ctx.saved_tensors
        d = h.shape[-1]
        rows = h.numel() // d
        dy = dy.contiguous()
        dh = torch.empty_like(h)
        dg = torch.empty(d, device=h.device, dtype=torch.float32)
        db = torch.empty(d, device=h.device, dtype=torch.float32)
        stats = torch.empty(rows * 2, device=h.device, dtype=torch.float32)
        gc = gamma.detach().contiguous()
        _count(2)
        check(lib.hoisdf_layernorm_bwd(h.data_ptr(), gc.data_ptr(), dy.data_ptr(), rows, d, dh.data_ptr(), dg.data_ptr(),
                                       db.data_ptr(), stats.data_ptr(), 0, _stream()), "hoisdf_layernorm_bwd")
        return dh, (dh if ctx.has_res else None), dg, db


def _gemm_batched(a, lda, ta, a_o, a_i, b, ldb, tb, b_o, b_i, c, ldc, c_o, c_i, m, n, k, alpha, bo, bi):
    _count(1)
    check(lib.hoisdf_gemm_f32_batched(a, lda, ta, a_o, a_i, b, ldb, tb, b_o, b_i, c, ldc, c_o, c_i, m, n, k, alpha, 0,
                                      bo, bi, _stream()), "hoisdf_gemm_f32_batched")


class AttentionFn(Function):
    """Multi-head attention core on (B*L, d) row matrices (head h = columns [64h, 64h+64)): softmax(q k^T / 8 + mask) v.
    q: (B*Lq, d) view with pitch ldq, k / v: (B*Lk, d) views; `mask` uint8 (Lq, Lk), non-zero = blocked; keys >= kv_valid are
    blocked for every query; `p_drop` = dropout on the probabilities (nn.MultiheadAttention's), fused into the softmax kernels."""

    @staticmethod
    def forward(ctx, q, k, v, batch, heads, lq, lk, mask, kv_valid, p_drop):
        d = heads * 64
        for t in (q, k, v):
            assert t.stride(1) == 1 and t.shape[1] == d
        dev = q.device
        out = torch.empty(batch * lq, d, device=dev, dtype=torch.float32)
        seed, lse = 0, None
        if p_drop > 0.0:
            # The keep decisions are a hash of (seed, position): nothing but the seed is kept for the backward (CPU
            # generator: no device sync, reproducible under torch.manual_seed)
            seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64))
        kend = lk if kv_valid is None else min(lk, kv_valid)
        if ops.USE_TENSOR_CORES and mask is None and lq > 32 and kend >= 128 and k.stride(0) == v.stride(0):
            # encoder self-attention: tensor-core flash kernel (dropout applied to P on its way to P.V); it also leaves the
            # log-sum-exp rows the tensor-core backward recomputes the probabilities from
            nbytes = lib.hoisdf_attention_workspace_bytes(batch, heads, lq, lk)
            ws = ops._attention_workspace(dev, nbytes)
            lse = torch.empty(batch * heads * lq, device=dev, dtype=torch.float32) if _ATTN_BWD_TC else None
            _count(4)
            check(lib.hoisdf_attention_train_fwd(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0),
                                                 out.data_ptr(), d, _ptr(lse), batch, heads, lq, lk, kend, float(p_drop), seed,
                                                 _ptr(ws), nbytes, _stream()), "hoisdf_attention_train_fwd")
        elif p_drop > 0.0:
            # materialised form (masked / short decoder attention): the dropped probabilities Pd (B, H, Lq, Lk) times V
            pd = AttentionFn._probs(q, k, batch, heads, lq, lk, mask, kv_valid, p_drop, seed)[1]
            _gemm_batched(pd.data_ptr(), lk, 0, heads * lq * lk, lq * lk, v.data_ptr(), v.stride(0), 0, lk * v.stride(0),
                          64, out.data_ptr(), d, lq * d, 64, lq, 64, lk, 1.0, batch, heads)
        else:
            if k.stride(0) != v.stride(0):
                raise ValueError("attention: k and v must share their row pitch")
            ops.attention(q, q.stride(0), k, v, k.stride(0), out, d, batch, heads, lq, lk, kv_valid=kv_valid, mask=mask,
                          tensor_cores=(mask is None))
        ctx.save_for_backward(q, k, v, mask, lse, out if lse is not None else None)
        ctx.geom = (batch, heads, lq, lk, kv_valid, p_drop, seed)
        return out

    @staticmethod
    def _probs(q, k, batch, heads, lq, lk, mask, kv_valid, p_drop=0.0, seed=0, want_p=False):
        """-> (P or None, Pd): softmax(q k^T / 8 [mask]) per (sample, head) and, with dropout, its dropped / rescaled copy.
        Without dropout Pd is P.  The scores are overwritten in place."""
        s = torch.empty(batch, heads, lq, lk, device=q.device, dtype=torch.float32)
        _gemm_batched(q.data_ptr(), q.stride(0), 0, lq * q.stride(0), 64, k.data_ptr(), k.stride(0), 1, lk * k.stride(0), 64,
                      s.data_ptr(), lk, heads * lq * lk, lq * lk, lq, lk, 64, 0.125, batch, heads)
        rows, valid = batch * heads * lq, lk if kv_valid is None else kv_valid
        _count(1)
        if p_drop <= 0.0:
            check(lib.hoisdf_softmax_rows_fwd(s.data_ptr(), lk, rows, lk, valid, _ptr(mask), 0 if mask is None else lq,
                                              s.data_ptr(), lk, _stream()), "hoisdf_softmax_rows_fwd")
            return s, s
        pd = torch.empty_like(s) if want_p else s
        check(lib.hoisdf_softmax_dropout_rows_fwd(s.data_ptr(), lk, rows, lk, valid, _ptr(mask), 0 if mask is None else lq,
                                                  s.data_ptr() if want_p else None, lk, pd.data_ptr(), lk, float(p_drop),
                                                  int(seed), _stream()), "hoisdf_softmax_dropout_rows_fwd")
        return (s if want_p else None), pd

    @staticmethod
    def backward(ctx, dout):
        q, k, v, mask, lse, out = ctx.saved_tensors
        batch, heads, lq, lk, kv_valid, p_drop, seed = ctx.geom
        d = heads * 64
        dev = q.device
        dout = dout.contiguous()
        if lse is not None:
            # tensor-core backward: P recomputed tile by tile from the forward's log-sum-exp rows, never materialised
            dq = torch.empty(batch * lq, d, device=dev, dtype=torch.float32)
            dk = torch.empty(batch * lk, d, device=dev, dtype=torch.float32)
            dv = torch.empty(batch * lk, d, device=dev, dtype=torch.float32)
            nbytes = lib.hoisdf_attention_bwd_workspace_bytes(batch, heads, lq, lk)
            ws = ops._attention_workspace(dev, nbytes)
            _count(10)
            check(lib.hoisdf_attention_bwd(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), out.data_ptr(),
                                           dout.data_ptr(), d, lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), d,
                                           batch, heads, lq, lk, lk if kv_valid is None else min(lk, kv_valid),
                                           float(p_drop), seed, _ptr(ws), nbytes, _stream()), "hoisdf_attention_bwd")
            return dq, dk, dv, None, None, None, None, None, None, None
        p, pd = AttentionFn._probs(q, k, batch, heads, lq, lk, mask, kv_valid, p_drop, seed, want_p=True)
        dq = torch.empty(batch * lq, d, device=dev, dtype=torch.float32)
        dk = torch.empty(batch * lk, d, device=dev, dtype=torch.float32)
        dv = torch.empty(batch * lk, d, device=dev, dtype=torch.float32)
        hh = heads * lq * lk
        # dV = Pd^T dO
        _gemm_batched(pd.data_ptr(), lk, 1, hh, lq * lk, dout.data_ptr(), d, 0, lq * d, 64, dv.data_ptr(), d, lk * d, 64,
                      lk, 64, lq, 1.0, batch, heads)
        # dPd = dO V^T (over the Pd buffer when it is a separate one)
        dp = pd if pd is not p else torch.empty(batch, heads, lq, lk, device=dev, dtype=torch.float32)
        _gemm_batched(dout.data_ptr(), d, 0, lq * d, 64, v.data_ptr(), v.stride(0), 1, lk * v.stride(0), 64, dp.data_ptr(),
                      lk, hh, lq * lk, lq, lk, 64, 1.0, batch, heads)
        # dS = P * (g - sum_j g_j P_j), g = dPd with the dropout mask folded in
        _count(1)
        if p_drop > 0.0:
            check(lib.hoisdf_softmax_dropout_rows_bwd(p.data_ptr(), lk, dp.data_ptr(), lk, batch * heads * lq, lk, dp.data_ptr(),
                                                      lk, float(p_drop), int(seed), _stream()), "hoisdf_softmax_dropout_rows_bwd")
        else:
            check(lib.hoisdf_softmax_rows_bwd(p.data_ptr(), lk, dp.data_ptr(), lk, batch * heads * lq, lk, dp.data_ptr(), lk,
                                              _stream()), "hoisdf_softmax_rows_bwd")
        # dQ = dS K / 8, dK = dS^T Q / 8
        _gemm_batched(dp.data_ptr(), lk, 0, hh, lq * lk, k.data_ptr(), k.stride(0), 0, lk * k.stride(0), 64, dq.data_ptr(),
                      d, lq * d, 64, lq, 64, lk, 0.125, batch, heads)
        _gemm_batched(dp.data_ptr(), lk, 1, hh, lq * lk, q.data_ptr(), q.stride(0), 0, lq * q.stride(0), 64, dk.data_ptr(),
                      d, lk * d, 64, lk, 64, lq, 0.125, batch, heads)
        return dq, dk, dv, None, None, None, None, None, None, None


class TokensFn(Function):
    """One own-field token block (upstream main/model.py:520-531 with sdf_activation :123-126):
    tokens (B, P, 256) = [xyz (3) | posenc (30) | fea (223) * sigmoid(sdf / beta) / beta].  Gradients: fea and beta (the SDF
    values are detached upstream, :483-484)."""

    @staticmethod
    def forward(ctx, fea, beta, xyz, pe, sdf):
        b, p, _ = xyz.shape
        fea = fea.contiguous()
        out = torch.empty(b, p, 256, device=fea.device, dtype=torch.float32)
        ops.tokens(xyz.contiguous(), pe.contiguous(), fea, sdf.contiguous(), beta.detach(), out, 0)
        ctx.save_for_backward(fea, beta, sdf.contiguous())
        return out

    @staticmethod
    def backward(ctx, dtok):
        fea, beta, sdf = ctx.saved_tensors
        b, p, f = fea.shape
        dtok = dtok.contiguous()
        dfea = torch.empty(b * p, f, device=fea.device, dtype=torch.float32)
        dbeta = torch.empty(1, device=fea.device, dtype=torch.float32)
        nbytes = lib.hoisdf_tokens_bwd_workspace_bytes(b, p)
        ws = torch.empty(max(int(nbytes), 4), device=fea.device, dtype=torch.uint8)
        _count(2)
        check(lib.hoisdf_tokens_bwd(dtok.data_ptr(), p, 0, fea.data_ptr(), f, sdf.data_ptr(), beta.detach().data_ptr(), b, p,
                                    dfea.data_ptr(), f, None, dbeta.data_ptr(), 0, ws.data_ptr(), nbytes, _stream()),
              "hoisdf_tokens_bwd")
        return dfea.view(b, p, f), dbeta.view_as(beta), None, None, None
