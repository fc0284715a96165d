"""Thin torch-tensor wrappers over the C ABI (include/hoisdf_b200.h).

PyTorch is plumbing here: it owns device memory and the current stream; every kernel that runs is one of
ours, launched through `libhoisdf_b200.so`.  Nothing in this file computes with torch ops.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _capi
from ._capi import ACT_NONE, ACT_RELU, ACT_SIGMOID, GATHER_CONCAT, GATHER_SUM, check, lib

ROW_LD = 516          # padded SDF row-buffer pitch (see csrc/sdf.cu)
ROWH_LD = 520         # pitch (halfs per plane) of the split-half row buffer of the FP16x3 path
SKIP_OFF_H = 296      # where relu(linh1) lands in it
DEC_IN = 289
DEC_IN_PAD = 292
SKIP_OFF = 292
H1 = 223
# TMEM accumulation chunk (K blocks of 32) of the FP16x3 GEMMs that only SCREEN candidates for the exact re-ranking:
# one drain per tile, the fastest setting -- the re-ranking's device-side check covers its ~1e-6 error
SCREEN_CHUNK_KB = 1 << 20
# algorithmic FLOPs of one SDFDecoder row (2*(289*512+512*223+512*512+512*512+512))
SDF_DECODER_FLOPS = 2.0 * (289 * 512 + 512 * 223 + 512 * 512 + 512 * 512 + 512)


# Launch accounting (bench.py reports it as `gpu_launches`) and optional per-launch CUDA-event profiling of the
# Linear kernel (bench.py's roofline leg): PROFILE is None or a list receiving (tag, flops, start_evt, end_evt).
STATS = {"launches": 0}
PROFILE = None


def _count(n: int = 1):
    STATS["launches"] += n


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (what, t.dtype))
    if not t.is_cuda:
        raise RuntimeError("%s must live on a CUDA device: hoisdf_b200 has no CPU path" % what)
    return t if t.is_contiguous() else t.contiguous()


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# ----------------------------------------------------------------------------------------------------
# Linear
# ----------------------------------------------------------------------------------------------------
# Linear implementation switch: True -> tcgen05 3xTF32 tensor-core kernel (fp32-grade accuracy) wherever the rows are
# not batched; False -> fp32 FMA kernel everywhere.  HOISDF_TC=0 in the environment disables the tensor path.
USE_TENSOR_CORES = os.environ.get("HOISDF_TC", "1") != "0"
# Which tensor-core Linear: "h3" = FP16x3 on split-half activations (csrc/linear_h3.cu, default),
# "tf32" = the 3xTF32 kernel on fp32 activations (csrc/linear_tc.cu).
TC_MODE = os.environ.get("HOISDF_TC_MODE", "h3")


if os.environ.get("HOISDF_H3_CHUNK"):          # developer knob: K blocks (of 32) per TMEM accumulation chunk
    lib.hoisdf_debug_h3_chunk(int(os.environ["HOISDF_H3_CHUNK"]))


def use_h3() -> bool:
    return USE_TENSOR_CORES and TC_MODE == "h3"


def split_tf32(w: torch.Tensor):
    """(w_hi, w_lo): w_hi = round-to-nearest TF32 of w, w_lo = w - w_hi (exact), same shape/pitch as w."""
    assert w.is_contiguous() and w.dtype == torch.float32 and w.is_cuda
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    _count(1)
    check(lib.hoisdf_split_tf32(w.data_ptr(), w.numel(), hi.data_ptr(), lo.data_ptr(), _stream()), "hoisdf_split_tf32")
    return hi, lo


@dataclass
class PackedLinear:
    """(N, ldw) fp32 weight with K zero-padded to a multiple of 4, plus bias and the TF32 hi/lo split."""
    w: torch.Tensor
    b: Optional[torch.Tensor]
    n: int
    k: int        # padded K the kernel contracts over (input rows must expose this many columns)
    ldw: int
    w_hi: Optional[torch.Tensor] = None
    w_lo: Optional[torch.Tensor] = None
    h3: Optional["PackedLinearH3"] = None      # FP16x3 planes of the same weights

    @staticmethod
    def pack(weight: torch.Tensor, bias: Optional[torch.Tensor], tensor_cores: bool = True) -> "PackedLinear":
        weight = weight.detach()
        n, k = weight.shape
        kp = round_up(k, 4)
        if kp == k and weight.is_contiguous() and weight.dtype == torch.float32 and weight.data_ptr() % 16 == 0:
            w = weight
        else:
            w = torch.zeros(n, kp, device=weight.device, dtype=torch.float32)
            w[:, :k] = weight
        b = None if bias is None else bias.detach().to(torch.float32).contiguous()
        hi = lo = h3 = None
        if tensor_cores and w.is_cuda:
            hi, lo = split_tf32(w)
            h3 = PackedLinearH3.pack(weight, b, k)
        return PackedLinear(w, b, n, kp, kp, hi, lo, h3)

    @staticmethod
    def from_packed(w: torch.Tensor, bias: Optional[torch.Tensor], tensor_cores: bool = True) -> "PackedLinear":
        """w already (N, K4) contiguous."""
        return PackedLinear.pack(w, bias, tensor_cores)

    def _sl(self, f):
        return (f(self.w), None if self.w_hi is None else f(self.w_hi), None if self.w_lo is None else f(self.w_lo))

    def cols(self, start: int, stop: int) -> "PackedLinear":
        """A K-slice W[:, start:stop] (no copy; start must be a multiple of 4). Bias dropped."""
        assert start % 4 == 0 and (stop - start) % 4 == 0
        w, hi, lo = self._sl(lambda t: t[:, start:stop])
        h3 = self.h3.cols(start, min(stop, self.h3.k)) if (self.h3 is not None and start % 8 == 0) else None
        return PackedLinear(w, None, self.n, stop - start, self.ldw, hi, lo, h3)

    def rows(self, start: int, stop: int) -> "PackedLinear":
        b = None if self.b is None else self.b[start:stop]
        w, hi, lo = self._sl(lambda t: t[start:stop])
        return PackedLinear(w, b, stop - start, self.k, self.ldw, hi, lo, None if self.h3 is None else self.h3.rows(start, stop))


def linear_raw(x_ptr: int, ldx: int, m: int, pw: PackedLinear, y_ptr: int, ldy: int, act: int = ACT_NONE,
               residual_ptr: Optional[int] = None, x_batch=(0, 0), y_batch=(0, 0), passes: int = 3):
    if m == 0:
        return
    tc = (USE_TENSOR_CORES and pw.w_lo is not None and y_batch[0] <= 0
          and (x_batch[0] <= 0 or m % x_batch[0] == 0))
    a = _capi.LinearArgs(
        x_ptr, ldx, x_batch[0], x_batch[1], (pw.w_hi if tc else pw.w).data_ptr(), pw.ldw, _ptr(pw.b), residual_ptr,
        y_ptr, ldy, y_batch[0], y_batch[1], m, pw.n, pw.k, act, pw.w_lo.data_ptr() if tc else None, passes)
    _count()
    if PROFILE is None:
        check(lib.hoisdf_linear_fwd(C.byref(a), _stream()), "hoisdf_linear_fwd")
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.hoisdf_linear_fwd(C.byref(a), _stream()), "hoisdf_linear_fwd")
        e1.record()
        PROFILE.append(("linear_tc" if tc else "linear", 2.0 * m * pw.n * pw.k, e0, e1))


def fma_only(pw: PackedLinear) -> PackedLinear:
    """The same weights, forced onto the fp32 FMA kernel."""
    return PackedLinear(pw.w, pw.b, pw.n, pw.k, pw.ldw, None, None, None)


def linear(x, pw: PackedLinear, act: int = ACT_NONE, out=None, residual: Optional[torch.Tensor] = None,
           out_ld: Optional[int] = None, passes: int = 3, split_out: bool = False, chunk_kb: Optional[int] = None,
           single: bool = False):
    """x: (M, >=K) 2-D fp32 (unit inner stride) or SplitRows.  Returns (M, N) fp32 (a view of a (M, out_ld) buffer if
    padded) -- or, on the FP16x3 path with split_out / a SplitRows `out`, the result in split-half format."""
    if isinstance(x, SplitRows) or isinstance(out, SplitRows) or (use_h3() and pw.h3 is not None):
        if pw.h3 is None or not use_h3():
            raise RuntimeError("split-half activations need the FP16x3 Linear (tensor cores enabled, TC_MODE='h3')")
        xs = x if isinstance(x, SplitRows) else split_rows(x[:, :pw.h3.k] if x.shape[1] > pw.h3.k else x)
        if out is None and not split_out:
            ld = out_ld or round_up(pw.n, 4)
            alloc = torch.empty if ld == pw.n else torch.zeros
            out = alloc(xs.rows, ld, device=xs.buf.device, dtype=torch.float32)[:, :pw.n]
        return linear_h3(xs, pw.h3, act, out=out, residual=residual, split_out=split_out, chunk_kb=chunk_kb,
                         single=single)
    assert x.dim() == 2 and x.stride(1) == 1 and x.shape[1] >= pw.k, (x.shape, x.stride(), pw.k)
    m = x.shape[0]
    if out is None:
        ld = out_ld or round_up(pw.n, 4)
        # padded tail columns are zeroed so a following layer can contract over them (times zero weights)
        alloc = torch.empty if ld == pw.n else torch.zeros
        buf = alloc(m, ld, device=x.device, dtype=torch.float32)
        out = buf[:, :pw.n]
    assert out.stride(1) == 1
    if residual is not None:
        assert residual.shape == out.shape and residual.stride() == out.stride()
    linear_raw(x.data_ptr(), x.stride(0), m, pw, out.data_ptr(), out.stride(0), act, _ptr(residual), passes=passes)
    return out


# ----------------------------------------------------------------------------------------------------
# FP16x3 tensor-core Linear on split-half activations (csrc/linear_h3.cu)
# ----------------------------------------------------------------------------------------------------
class SplitRows:
    """A (rows, cols) fp32-valued matrix in split-half format: `buf` is a (rows, 2, ld) fp16 tensor whose
    [:, 0] plane holds hi = fp16(x) and [:, 1] plane lo = fp16((x - hi) * 2^11); `col0` selects a column window."""

    def __init__(self, buf: torch.Tensor, cols: int, col0: int = 0):
        assert buf.dtype == torch.float16 and buf.dim() == 3 and buf.shape[1] == 2 and buf.stride(2) == 1
        assert buf.stride(1) % 8 == 0 and buf.stride(0) % 8 == 0 and col0 % 8 == 0 and buf.data_ptr() % 16 == 0
        self.buf, self.cols, self.col0 = buf, cols, col0

    @staticmethod
    def empty(rows: int, cols: int, device, ld: Optional[int] = None) -> "SplitRows":
        ld = ld or round_up(cols, 8)
        return SplitRows(torch.empty(rows, 2, ld, device=device, dtype=torch.float16), cols)

    @property
    def rows(self) -> int:
        return self.buf.shape[0]

    @property
    def ld(self) -> int:           # row pitch in halfs
        return self.buf.stride(0)

    @property
    def hi_ptr(self) -> int:
        return self.buf.data_ptr() + 2 * self.col0

    @property
    def lo_ptr(self) -> int:
        return self.buf.data_ptr() + 2 * (self.buf.stride(1) + self.col0)

    def window(self, col0: int, cols: int) -> "SplitRows":
        return SplitRows(self.buf, cols, self.col0 + col0)

    def head(self, rows: int) -> "SplitRows":
        return SplitRows(self.buf[:rows], self.cols, self.col0)

    def float(self) -> torch.Tensor:
        """fp32 copy (tests / diagnostics)."""
        out = torch.empty(self.rows, self.cols, device=self.buf.device, dtype=torch.float32)
        _count(1)
        check(lib.hoisdf_join_rows(self.hi_ptr, self.lo_ptr, self.ld, self.rows, self.cols, out.data_ptr(), self.cols,
                                   _stream()), "hoisdf_join_rows")
        return out


def split_rows(x: torch.Tensor, out: Optional[SplitRows] = None, kpad: Optional[int] = None) -> SplitRows:
    """fp32 (M, K) rows (unit inner stride) -> split-half; columns [K, kpad) are zeroed."""
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.float32
    m, k = x.shape
    kpad = kpad or round_up(k, 4)
    if out is None:
        out = SplitRows.empty(m, k, x.device, round_up(kpad, 8))
    _count(1)
    check(lib.hoisdf_split_rows(x.data_ptr(), m, k, x.stride(0) if m > 1 else max(x.stride(0), k), kpad, out.hi_ptr,
                                out.lo_ptr, out.ld, _stream()), "hoisdf_split_rows")
    return out


@dataclass
class PackedLinearH3:
    """Three fp16 weight planes (A = w_hi * 2^11, B = w_hi, C = w_lo * 2^11), each (N, ld) halfs, + fp32 bias."""
    planes: torch.Tensor          # (3, N, ld) fp16
    b: Optional[torch.Tensor]
    n: int
    k: int
    row0: int = 0                 # row window (N slice)
    col0: int = 0                 # column window (K slice), multiple of 8
    chunk_kb: int = 0             # TMEM accumulation chunk of the launches using these weights (0 = kernel default)
    scale: float = 1.0            # power of two the weights were divided by before packing (multiplied back in the epilogue)

    @staticmethod
    def pack(weight: torch.Tensor, bias: Optional[torch.Tensor], k: Optional[int] = None,
             chunk_kb: int = 0, assume_max: Optional[float] = None) -> "PackedLinearH3":
        """`assume_max`: the caller's bound on max |w| -- skips the host read-back of the real maximum (the training
        path packs operands every step and scales them on the device, hoisdf_b200/autograd.py)."""
        w = weight.detach().to(torch.float32)
        if w.stride(1) != 1:
            w = w.contiguous()
        n = w.shape[0]
        k = k or w.shape[1]
        # plane A = w_hi * 2^11 must fit fp16: layers with |w| >= 16 (a BatchNorm fold with a tiny running variance can do
        # that) are divided by a power of two -- exact -- which the GEMM epilogue multiplies back (`w_scale`)
        wmax = float(assume_max) if assume_max is not None else (float(w.abs().max()) if w.numel() else 0.0)
        if not math.isfinite(wmax):
            raise ValueError("FP16x3 Linear: non-finite weight")
        scale = 1.0
        if wmax >= 16.0:
            scale = float(2.0 ** math.ceil(math.log2(wmax / 8.0)))
            w = w * (1.0 / scale)
        ld = round_up(k, 8)
        planes = torch.empty(3, n, ld, device=w.device, dtype=torch.float16)
        _count(1)
        check(lib.hoisdf_pack_h3(w.data_ptr(), n, k, w.stride(0), planes[0].data_ptr(), planes[1].data_ptr(),
                                 planes[2].data_ptr(), ld, _stream()), "hoisdf_pack_h3")
        b = None if bias is None else bias.detach().to(torch.float32).contiguous()
        return PackedLinearH3(planes, b, n, k, chunk_kb=chunk_kb, scale=scale)

    @property
    def ld(self) -> int:
        return self.planes.stride(1)

    def plane_ptr(self, i: int) -> int:
        return self.planes[i].data_ptr() + 2 * (self.row0 * self.ld + self.col0)

    def cols(self, start: int, stop: int) -> "PackedLinearH3":
        assert start % 8 == 0
        return PackedLinearH3(self.planes, None, self.n, stop - start, self.row0, self.col0 + start, self.chunk_kb,
                              self.scale)

    def rows(self, start: int, stop: int) -> "PackedLinearH3":
        b = None if self.b is None else self.b[start:stop]
        return PackedLinearH3(self.planes, b, stop - start, self.k, self.row0 + start, self.col0, self.chunk_kb,
                              self.scale)


def linear_h3(x: SplitRows, pw: PackedLinearH3, act: int = ACT_NONE, out=None, residual: Optional[torch.Tensor] = None,
              split_out: bool = False, x_batch=(0, 0), m: Optional[int] = None, chunk_kb: Optional[int] = None,
              residual_split: Optional[SplitRows] = None, single: bool = False, y_scale: Optional[torch.Tensor] = None,
              split_k: bool = False):
    """Y = act(X . W^T + b) (+ residual) on the FP16x3 tensor-core kernel.  `out` is an fp32 (M, N) tensor view (unit
    inner stride) or a SplitRows window; allocated when None (fp32, or split-half if split_out).
    x_batch = (rows_per_batch, batch_stride in halfs) walks strided row groups of `x` (m rows in total).
    y_scale: one fp32 value ON THE DEVICE the product is multiplied by in the epilogue (training backward).
    split_k: allow the few-tile / long-contraction launches to split K over the SMs (summation order then varies)."""
    m = x.rows if m is None else m
    assert x.cols >= pw.k, (x.cols, pw.k)
    if out is None:
        out = (SplitRows.empty(m, pw.n, x.buf.device) if split_out
               else torch.empty(m, round_up(pw.n, 4), device=x.buf.device, dtype=torch.float32)[:, :pw.n])
    if m == 0:
        return out
    a = _capi.LinearH3Args()
    a.x_hi, a.x_lo, a.ldx, a.x_rows_per_batch, a.x_batch_stride = x.hi_ptr, x.lo_ptr, x.ld, x_batch[0], x_batch[1]
    a.w_a, a.w_b, a.w_c, a.ldw = pw.plane_ptr(0), pw.plane_ptr(1), pw.plane_ptr(2), pw.ld
    a.bias, a.residual = _ptr(pw.b), _ptr(residual)
    if isinstance(out, SplitRows):
        assert residual is None and out.cols >= pw.n
        a.y, a.ldy, a.y_hi, a.y_lo, a.ldyh = None, 0, out.hi_ptr, out.lo_ptr, out.ld
    else:
        assert out.stride(1) == 1 and out.dtype == torch.float32
        if residual is not None:
            assert residual.stride() == out.stride()
        a.y, a.ldy, a.y_hi, a.y_lo, a.ldyh = out.data_ptr(), out.stride(0), None, None, 0
    a.m, a.n, a.k, a.act = m, pw.n, pw.k, act
    a.chunk_kb = int(pw.chunk_kb if chunk_kb is None else chunk_kb)
    a.single_pass = int(single)
    a.w_scale = float(pw.scale)
    a.y_scale = _ptr(y_scale)
    a.split_k = int(split_k)
    if residual_split is not None:
        assert residual_split.cols >= pw.n and residual_split.rows >= m
        a.res_hi, a.res_lo, a.ldr = residual_split.hi_ptr, residual_split.lo_ptr, residual_split.ld
    _count()
    if PROFILE is None:
        check(lib.hoisdf_linear_h3_fwd(C.byref(a), _stream()), "hoisdf_linear_h3_fwd")
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.hoisdf_linear_h3_fwd(C.byref(a), _stream()), "hoisdf_linear_h3_fwd")
        e1.record()
        PROFILE.append(("linear_h3", 2.0 * m * pw.n * pw.k, e0, e1, "M=%d N=%d K=%d chunk=%d %s%s" % (
            m, pw.n, pw.k, a.chunk_kb, "split" if isinstance(out, SplitRows) else "f32", " single" if single else "")))
    return out


NARROW_MAX_N = 12          # measured: at N = 20 (M = 295k) the tensor-core kernel is 2x faster than the narrow one


def linear_narrow(x: SplitRows, w: torch.Tensor, bias: Optional[torch.Tensor], act: int = ACT_NONE,
                  out: Optional[torch.Tensor] = None, m: Optional[int] = None) -> torch.Tensor:
    """Y (M, N <= 24) = act(X . W^T + b): split-half X, fp32 W (N, >= K, unit inner stride, K even) -> fp32."""
    assert w.dim() == 2 and w.stride(1) == 1 and w.dtype == torch.float32 and w.is_cuda
    n = w.shape[0]
    k = w.shape[1] if w.shape[1] <= x.cols else x.cols
    m = x.rows if m is None else m
    if out is None:
        out = torch.empty(m, n, device=w.device, dtype=torch.float32)
    assert out.stride(1) == 1 and out.shape[0] >= m
    _count(1)
    check(lib.hoisdf_linear_narrow_split_fwd(x.hi_ptr, x.lo_ptr, x.ld, m, w.data_ptr(), w.stride(0), _ptr(bias), n, k, act,
                                             out.data_ptr(), out.stride(0), _stream()), "hoisdf_linear_narrow_split_fwd")
    return out


def conv_h3(x: SplitRows, batch: int, in_h: int, in_w: int, cin: int, pw: PackedLinearH3, taps, out_h: int, out_w: int,
            *, stride: int = 1, act: int = ACT_NONE, out=None, out_strides=None, out_offset: int = 0,
            chunk_kb: Optional[int] = None, residual_split: Optional[SplitRows] = None):
    """Implicit-GEMM convolution on the FP16x3 kernel.  x: NHWC pixels (batch*in_h*in_w rows of >= cin columns) in
    split-half format; pw: planes of the (cout, len(taps)*cin) weight matrix; taps: [(dy, dx), ...].
    out: fp32 (rows, >= cout) tensor or SplitRows window; out_strides = (sx, sy, sb) in elements (default: dense
    NHWC rows of `out`), out_offset = element offset of output pixel (0, 0, 0) from the start of `out`."""
    assert x.rows == batch * in_h * in_w and x.cols >= cin and pw.k == len(taps) * cin
    a = _capi.ConvH3Args()
    a.x_hi, a.x_lo, a.batch, a.in_h, a.in_w, a.cin, a.ldx = x.hi_ptr, x.lo_ptr, batch, in_h, in_w, cin, x.ld
    a.w_a, a.w_b, a.w_c, a.ldw, a.bias = pw.plane_ptr(0), pw.plane_ptr(1), pw.plane_ptr(2), pw.ld, _ptr(pw.b)
    a.taps, a.stride = len(taps), stride
    for i, (dy, dx) in enumerate(taps):
        a.tap_dy[i], a.tap_dx[i] = dy, dx
    a.out_h, a.out_w, a.cout = out_h, out_w, pw.n
    if isinstance(out, SplitRows):
        pitch = out.ld
        a.y, a.y_hi, a.y_lo = None, out.hi_ptr + 2 * out_offset, out.lo_ptr + 2 * out_offset
    else:
        assert out.dtype == torch.float32 and out.stride(-1) == 1
        pitch = out.stride(-2)
        a.y, a.y_hi, a.y_lo = out.data_ptr() + 4 * out_offset, None, None
    sx, sy, sb = out_strides if out_strides is not None else (pitch, out_w * pitch, out_h * out_w * pitch)
    a.y_sx, a.y_sy, a.y_sb = sx, sy, sb
    a.act = act
    a.chunk_kb = int(pw.chunk_kb if chunk_kb is None else chunk_kb)
    a.w_scale = float(pw.scale)
    if residual_split is not None:
        assert residual_split.cols >= pw.n and residual_split.rows >= batch * out_h * out_w
        a.res_hi, a.res_lo, a.ldr = residual_split.hi_ptr, residual_split.lo_ptr, residual_split.ld
    _count()
    flops = 2.0 * batch * out_h * out_w * pw.n * pw.k
    if PROFILE is None:
        check(lib.hoisdf_conv_h3_fwd(C.byref(a), _stream()), "hoisdf_conv_h3_fwd")
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.hoisdf_conv_h3_fwd(C.byref(a), _stream()), "hoisdf_conv_h3_fwd")
        e1.record()
        PROFILE.append(("conv_h3", flops, e0, e1, "conv B=%d %dx%d->%dx%d Cin=%d Cout=%d taps=%d s=%d chunk=%d" % (
            batch, in_h, in_w, out_h, out_w, cin, pw.n, len(taps), stride, a.chunk_kb)))
    return out


def nchw_to_split(x: torch.Tensor, out: SplitRows):
    """(B, C, H, W) fp32 feature map -> NHWC pixels in split-half format, written at `out`'s column window."""
    assert x.dim() == 4 and x.dtype == torch.float32 and x.is_cuda
    b, c, h, w = x.shape
    assert out.rows == b * h * w and out.cols >= c
    nhwc = x.permute(0, 2, 3, 1)
    if nhwc.is_contiguous():          # channels_last memory: a plain row-wise split
        return split_rows(nhwc.reshape(b * h * w, c), out=out, kpad=c if c % 4 == 0 else None)
    x = x.contiguous()
    _count(1)
    check(lib.hoisdf_nchw_to_nhwc_split(x.data_ptr(), out.hi_ptr, out.lo_ptr, b, c, h, w, out.ld, _stream()),
          "hoisdf_nchw_to_nhwc_split")
    return out


STEM_K = 147          # 7 x 7 x 3 im2col columns of the ResNet stem ...
STEM_K_PAD = 160      # ... padded to 5 K blocks of 32 halfs


def stem_im2col(img: torch.Tensor, out: Optional[SplitRows] = None) -> SplitRows:
    """(B, 3, H, W) fp32 NCHW image -> (B*H/2*W/2, 160) split-half im2col rows of the 7x7 stride-2 pad-3 stem."""
    assert img.dim() == 4 and img.shape[1] == 3 and img.dtype == torch.float32 and img.is_cuda
    img = img.contiguous()
    b, _, h, w = img.shape
    if out is None:
        out = SplitRows.empty(b * (h // 2) * (w // 2), STEM_K_PAD, img.device)
    assert out.rows == b * (h // 2) * (w // 2) and out.cols >= STEM_K_PAD
    _count(1)
    check(lib.hoisdf_stem_im2col_split(img.data_ptr(), b, h, w, out.hi_ptr, out.lo_ptr, out.ld, _stream()),
          "hoisdf_stem_im2col_split")
    return out


def maxpool3x3s2(x: SplitRows, batch: int, h: int, w: int, c: int, out: Optional[SplitRows] = None) -> SplitRows:
    """NHWC split-half (batch, h, w, c) -> (batch, h/2, w/2, c): nn.MaxPool2d(3, 2, 1)."""
    assert x.rows == batch * h * w and x.cols >= c
    if out is None:
        out = SplitRows.empty(batch * (h // 2) * (w // 2), c, x.buf.device)
    _count(1)
    check(lib.hoisdf_maxpool3x3s2_split(x.hi_ptr, x.lo_ptr, x.ld, batch, h, w, c, out.hi_ptr, out.lo_ptr, out.ld,
                                        _stream()), "hoisdf_maxpool3x3s2_split")
    return out


def fold_weight_norm(g: Optional[torch.Tensor], v: torch.Tensor, cols_out: Optional[int] = None,
                     src_col: Optional[torch.Tensor] = None) -> torch.Tensor:
    v = _f32c(v.detach(), "weight_v")
    rows, cols = v.shape
    cols_out = cols_out or round_up(cols, 4)
    out = torch.empty(rows, cols_out, device=v.device, dtype=torch.float32)
    gg = None if g is None else _f32c(g.detach().reshape(-1), "weight_g")
    if src_col is None:
        src_col = torch.arange(cols_out, device=v.device, dtype=torch.int32)
        src_col[cols:] = -1
    _count(1)
    check(lib.hoisdf_fold_weight_norm(_ptr(gg), v.data_ptr(), rows, cols, out.data_ptr(), cols_out,
                                      src_col.data_ptr(), cols_out, _stream()), "hoisdf_fold_weight_norm")
    return out


# ----------------------------------------------------------------------------------------------------
# Pyramid / gather
# ----------------------------------------------------------------------------------------------------
def to_nhwc(x: torch.Tensor) -> torch.Tensor:
    """(B,C,H,W) logical NCHW tensor -> (B,H,W,C) contiguous; zero-copy when x is already channels_last."""
    assert x.dim() == 4 and x.dtype == torch.float32 and x.is_cuda
    b, c, h, w = x.shape
    p = x.permute(0, 2, 3, 1)
    if p.is_contiguous():
        return p
    x = x.contiguous()
    out = torch.empty(b, h, w, c, device=x.device, dtype=torch.float32)
    _count(1)
    check(lib.hoisdf_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), b, c, h, w, _stream()), "hoisdf_nchw_to_nhwc")
    return out


def make_pyramid(maps: Sequence[torch.Tensor], img_hw=(256, 256)) -> _capi.Pyramid:
    """maps: list of (B,H,W,C) contiguous fp32 tensors."""
    p = _capi.Pyramid()
    p.levels = len(maps)
    p.img_h, p.img_w = int(img_hw[0]), int(img_hw[1])
    for i, m in enumerate(maps):
        assert m.is_contiguous() and m.dtype == torch.float32 and m.is_cuda
        p.map[i] = m.data_ptr()
        p.h[i], p.w[i], p.c[i] = m.shape[1], m.shape[2], m.shape[3]
    return p


def gather(maps: Sequence[torch.Tensor], uv: torch.Tensor, batch: int, *, mode: int, out: torch.Tensor,
           row_offsets: Optional[torch.Tensor] = None, rows_per_sample: int = 0,
           bias: Optional[torch.Tensor] = None, act: int = ACT_NONE, img_hw=(256, 256)):
    rows = uv.shape[0]
    assert uv.is_contiguous() and uv.shape[1] == 2
    pyr = make_pyramid(maps, img_hw)
    _count(1)
    if isinstance(out, SplitRows):
        check(lib.hoisdf_gather_split_fwd(C.byref(pyr), uv.data_ptr(), rows, _ptr(row_offsets), batch, rows_per_sample,
                                          mode, _ptr(bias), act, out.hi_ptr, out.lo_ptr, out.ld, _stream()),
              "hoisdf_gather_split_fwd")
        return out
    assert out.stride(1) == 1
    check(lib.hoisdf_gather_fwd(C.byref(pyr), uv.data_ptr(), rows, _ptr(row_offsets), batch, rows_per_sample, mode,
                                _ptr(bias), act, out.data_ptr(), out.stride(0), _stream()), "hoisdf_gather_fwd")
    return out


# ----------------------------------------------------------------------------------------------------
# Lattice / projection
# ----------------------------------------------------------------------------------------------------
def lattice_count(center, cam_intr, bbox, sdf_scale: float, bins: int):
    b = center.shape[0]
    chunks = lib.hoisdf_lattice_chunks(bins)
    counts = torch.empty(b, chunks, device=center.device, dtype=torch.int32)
    offsets = torch.empty(b + 1, device=center.device, dtype=torch.int64)
    _count(2)
    check(lib.hoisdf_lattice_count(center.data_ptr(), cam_intr.data_ptr(), bbox.data_ptr(), float(sdf_scale), b,
                                   bins, counts.data_ptr(), offsets.data_ptr(), _stream()), "hoisdf_lattice_count")
    return counts, offsets


def lattice_compact(center, cam_intr, bbox, sdf_scale: float, bins: int, counts, offsets, total: int):
    b = center.shape[0]
    cand_index = torch.empty(max(total, 1), device=center.device, dtype=torch.int32)
    cand_uv = torch.empty(max(total, 1), 2, device=center.device, dtype=torch.float32)
    _count(1)
    check(lib.hoisdf_lattice_compact(center.data_ptr(), cam_intr.data_ptr(), bbox.data_ptr(), float(sdf_scale), b,
                                     bins, counts.data_ptr(), offsets.data_ptr(), cand_index.data_ptr(),
                                     cand_uv.data_ptr(), _stream()), "hoisdf_lattice_compact")
    return cand_index, cand_uv


def project_points(points, center, cam_intr, sdf_scale: float, want_cam: bool = True):
    b, p, _ = points.shape
    cam = torch.empty(b, p, 3, device=points.device, dtype=torch.float32) if want_cam else None
    uv = torch.empty(b * p, 2, device=points.device, dtype=torch.float32)
    _count(1)
    check(lib.hoisdf_project_points(points.data_ptr(), center.data_ptr(), cam_intr.data_ptr(), float(sdf_scale), b, p,
                                    _ptr(cam), uv.data_ptr(), _stream()), "hoisdf_project_points")
    return cam, uv


# ----------------------------------------------------------------------------------------------------
# SDF decoder
# ----------------------------------------------------------------------------------------------------
@dataclass
class PackedSdfDecoder:
    tensors: List[torch.Tensor]      # keeps the packed buffers alive
    struct_fma: _capi.SdfWeights     # fp32 FMA kernels
    struct_tc: Optional[_capi.SdfWeights]   # tcgen05 3xTF32 kernels (None when tensor cores are disabled)
    struct_tc1: Optional[_capi.SdfWeights] = None   # tcgen05 single-pass TF32 (candidate screening only)
    struct_h3: Optional[_capi.SdfWeightsH3] = None  # tcgen05 FP16x3 on split-half rows

    def struct(self, exact: bool = False, screening: bool = False):
        if exact or self.struct_tc is None or not USE_TENSOR_CORES:
            return self.struct_fma
        return self.struct_tc1 if screening else self.struct_tc


def pack_sdf_decoder(dec_params: dict) -> PackedSdfDecoder:
    """dec_params: {'linh0.weight_g': ..., 'linh0.weight_v': ..., 'linh0.bias': ..., ..., 'linh4.weight', 'linh4.bias'}.

    Folds weight-norm (upstream sdf_net.py:57-62) and lays linh2's columns out for the padded row buffer:
    upstream input of linh2 is cat[h1 (223), input (289)]; ours is [input (289), 0 0 0, h1 (223), 0].
    """
    dev = dec_params["linh0.weight_v"].device
    w0 = fold_weight_norm(dec_params["linh0.weight_g"], dec_params["linh0.weight_v"], DEC_IN_PAD)
    w1 = fold_weight_norm(dec_params["linh1.weight_g"], dec_params["linh1.weight_v"], 512)
    src = torch.full((ROW_LD,), -1, dtype=torch.int32)
    src[0:DEC_IN] = torch.arange(H1, H1 + DEC_IN, dtype=torch.int32)
    src[SKIP_OFF:SKIP_OFF + H1] = torch.arange(0, H1, dtype=torch.int32)
    w2 = fold_weight_norm(dec_params["linh2.weight_g"], dec_params["linh2.weight_v"], ROW_LD, src.to(dev))
    w3 = fold_weight_norm(dec_params["linh3.weight_g"], dec_params["linh3.weight_v"], 512)
    w4 = _f32c(dec_params["linh4.weight"].detach().reshape(-1).clone(), "linh4.weight")
    bs = [_f32c(dec_params["linh%d.bias" % i].detach().clone(), "bias") for i in range(5)]
    keep = [w0, w1, w2, w3, w4] + bs
    ws = (w0, w1, w2, w3)
    bp = [b.data_ptr() for b in bs]
    s_fma = _capi.SdfWeights(ws[0].data_ptr(), bp[0], ws[1].data_ptr(), bp[1], ws[2].data_ptr(), bp[2],
                             ws[3].data_ptr(), bp[3], w4.data_ptr(), bp[4], None, None, None, None, 3)
    splits = [split_tf32(w) for w in ws]
    keep += [t for hl in splits for t in hl]
    his, los = [hl[0].data_ptr() for hl in splits], [hl[1].data_ptr() for hl in splits]
    s_tc = _capi.SdfWeights(his[0], bp[0], his[1], bp[1], his[2], bp[2], his[3], bp[3], w4.data_ptr(), bp[4],
                            los[0], los[1], los[2], los[3], 3)
    s_tc1 = _capi.SdfWeights(his[0], bp[0], his[1], bp[1], his[2], bp[2], his[3], bp[3], w4.data_ptr(), bp[4],
                             los[0], los[1], los[2], los[3], 1)
    # FP16x3: the same folded weights as fp16 planes; linh2's columns follow the split-half row layout
    # [input (289) | 0 x7 | h1 (223) | 0]
    srch = torch.full((ROWH_LD,), -1, dtype=torch.int32)
    srch[0:DEC_IN] = torch.arange(H1, H1 + DEC_IN, dtype=torch.int32)
    srch[SKIP_OFF_H:SKIP_OFF_H + H1] = torch.arange(0, H1, dtype=torch.int32)
    w2h = fold_weight_norm(dec_params["linh2.weight_g"], dec_params["linh2.weight_v"], ROWH_LD, srch.to(dev))
    h3s = [PackedLinearH3.pack(w0, None, DEC_IN), PackedLinearH3.pack(w1, None, 512),
           PackedLinearH3.pack(w2h, None, SKIP_OFF_H + H1), PackedLinearH3.pack(w3, None, 512)]
    if any(hp.scale != 1.0 for hp in h3s):
        raise ValueError("SDF decoder weights >= 16 in magnitude are not supported by the FP16x3 decoder chain")
    s_h3 = _capi.SdfWeightsH3()
    for i, hp in enumerate(h3s):
        for j in range(3):
            s_h3.w[i][j] = hp.plane_ptr(j)
        s_h3.ldw[i] = hp.ld
        s_h3.b[i] = bp[i]
    s_h3.w4, s_h3.b4 = w4.data_ptr(), bp[4]
    keep += [hp.planes for hp in h3s]
    return PackedSdfDecoder(keep, s_fma, s_tc, s_tc1, s_h3)


def posenc(rows_buf, *, lattice_index=None, points=None, bins: int = 64):
    _count(1)
    if isinstance(rows_buf, SplitRows):
        check(lib.hoisdf_posenc_split_fwd(_ptr(lattice_index), _ptr(points), rows_buf.rows, bins, rows_buf.hi_ptr,
                                          rows_buf.lo_ptr, rows_buf.ld, _stream()), "hoisdf_posenc_split_fwd")
        return
    rows = rows_buf.shape[0]
    check(lib.hoisdf_posenc_fwd(_ptr(lattice_index), _ptr(points), rows, bins, rows_buf.data_ptr(),
                                rows_buf.stride(0), 256, _stream()), "hoisdf_posenc_fwd")


def sdf_decoder(packed: PackedSdfDecoder, rows_buf, h_a=None, h_b=None, clamp: float = 0.0,
                out: Optional[torch.Tensor] = None, exact: bool = False, screening: bool = False,
                chunk_kb: int = 0, single: bool = False) -> torch.Tensor:
    if isinstance(rows_buf, SplitRows):
        rows, dev = rows_buf.rows, rows_buf.buf.device
        h_a = h_a if h_a is not None else SplitRows.empty(rows, 512, dev)
        h_b = h_b if h_b is not None else SplitRows.empty(rows, 512, dev)
        out = out if out is not None else torch.empty(rows, device=dev, dtype=torch.float32)
        assert h_a.ld == h_b.ld and rows_buf.col0 == 0 and h_a.col0 == 0 and h_b.col0 == 0
        _count(5)
        packed.struct_h3.chunk_kb = int(chunk_kb)
        packed.struct_h3.single_pass = int(single)
        if PROFILE is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        check(lib.hoisdf_sdf_decoder_h3_fwd(C.byref(packed.struct_h3), rows_buf.hi_ptr, rows_buf.lo_ptr, rows_buf.ld,
                                            rows, h_a.hi_ptr, h_a.lo_ptr, h_b.hi_ptr, h_b.lo_ptr, h_a.ld,
                                            out.data_ptr(), float(clamp), _stream()), "hoisdf_sdf_decoder_h3_fwd")
        if PROFILE is not None:
            e1.record()
            PROFILE.append(("sdf_decoder_h3", SDF_DECODER_FLOPS * rows, e0, e1, "sdf_decoder rows=%d chunk=%d%s" % (
                rows, chunk_kb, " single" if single else "")))
        return out
    rows = rows_buf.shape[0]
    dev = rows_buf.device
    h_a = h_a if h_a is not None else torch.empty(rows, 512, device=dev, dtype=torch.float32)
    h_b = h_b if h_b is not None else torch.empty(rows, 512, device=dev, dtype=torch.float32)
    out = out if out is not None else torch.empty(rows, device=dev, dtype=torch.float32)
    _count(5)
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib.hoisdf_sdf_decoder_fwd(C.byref(packed.struct(exact, screening)), rows_buf.data_ptr(), rows_buf.stride(0), rows,
                                     h_a.data_ptr(), h_b.data_ptr(), out.data_ptr(), float(clamp), _stream()),
          "hoisdf_sdf_decoder_fwd")
    if PROFILE is not None:
        e1.record()
        PROFILE.append(("sdf_decoder", SDF_DECODER_FLOPS * rows, e0, e1))
    return out


def maps_to_half(maps: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """fp16 copies of NHWC fp32 maps (the projected pyramid the chain kernel's gather mode reads)."""
    out = []
    for m in maps:
        assert m.is_contiguous() and m.dtype == torch.float32 and m.numel() % 8 == 0
        h = torch.empty(m.shape, device=m.device, dtype=torch.float16)
        _count(1)
        check(lib.hoisdf_f32_to_f16(m.data_ptr(), h.data_ptr(), m.numel(), _stream()), "hoisdf_f32_to_f16")
        out.append(h)
    return out


def _pyramid_h(gmaps16: Sequence[torch.Tensor], img_hw) -> _capi.PyramidH:
    pyr = _capi.PyramidH()
    pyr.levels, pyr.c, pyr.img_h, pyr.img_w = len(gmaps16), 512, int(img_hw[0]), int(img_hw[1])
    for i, m in enumerate(gmaps16):
        assert m.is_contiguous() and m.dtype == torch.float16 and m.shape[3] == 512 and m.is_cuda
        pyr.map[i], pyr.h[i], pyr.w[i] = m.data_ptr(), m.shape[1], m.shape[2]
    return pyr


def gather_h16(gmaps16: Sequence[torch.Tensor], uv: torch.Tensor, batch: int, out: SplitRows, *, row_offsets=None,
               rows_per_sample: int = 0, bias=None, act: int = ACT_NONE, img_hw=(256, 256)) -> SplitRows:
    """SUM-mode gather of fp16 maps into the HI plane of `out` (the lo plane is left untouched): the screening path."""
    rows = uv.shape[0]
    assert uv.is_contiguous() and uv.shape[1] == 2 and out.cols >= 512 and out.rows >= rows
    pyr = _pyramid_h(gmaps16, img_hw)
    _count(1)
    check(lib.hoisdf_gather_sum_h16_fwd(C.byref(pyr), uv.data_ptr(), rows, _ptr(row_offsets), batch, rows_per_sample,
                                        _ptr(bias), act, out.hi_ptr, out.ld, _stream()), "hoisdf_gather_sum_h16_fwd")
    return out


def sdf_chain(packed: PackedSdfDecoder, out: torch.Tensor, *, sdfin1: Optional[PackedLinear] = None,
              a0: Optional[SplitRows] = None, x: Optional[SplitRows] = None, lattice_index=None, points=None,
              bins: int = 64, clamp: float = 0.0, gmaps16: Optional[Sequence[torch.Tensor]] = None, uv=None,
              row_offsets=None, batch: int = 0, rows_per_sample: int = 0, bias0=None, img_hw=(256, 256)) -> torch.Tensor:
    """The fused candidate chain (csrc/sdf_chain.cu): [gather ->] linear_sdfin.layers.1 -> posenc/xyz -> linh0..linh4 ->
    tanh in ONE persistent tcgen05 kernel, single-product fp16 (screening arithmetic).  Row source, exactly one of:
    `gmaps16` + `uv` (gather mode: fp16 projected maps (B,H,W,512), projected pixels (rows, 2)), `a0` = hi plane of
    relu(linear_sdfin.layers.0) rows (rows mode), `x` = hi plane of the decoder input rows (decoder-only mode)."""
    assert packed.struct_h3 is not None and sum(v is not None for v in (a0, x, gmaps16)) == 1
    a = _capi.SdfChainArgs()
    keep = None
    if x is not None:
        rows = x.rows
        assert x.cols >= DEC_IN and x.ld >= SKIP_OFF_H
        a.x, a.ldx = x.hi_ptr, x.ld
    else:
        assert sdfin1 is not None and sdfin1.h3 is not None and sdfin1.n == 256
        assert sdfin1.h3.scale == 1.0, "linear_sdfin.layers.1 weights >= 16 in magnitude: use the unfused chain"
        a.w_s1, a.ldw_s1, a.b_s1 = sdfin1.h3.plane_ptr(1), sdfin1.h3.ld, _ptr(sdfin1.b)
        a.lattice_index, a.points, a.bins = _ptr(lattice_index), _ptr(points), int(bins)
        if a0 is not None:
            rows = a0.rows
            assert a0.cols >= 512
            a.a0, a.lda0 = a0.hi_ptr, a0.ld
        else:
            rows = uv.shape[0]
            assert uv.is_contiguous() and uv.shape[1] == 2 and uv.dtype == torch.float32 and bias0 is not None
            pyr = _pyramid_h(gmaps16, img_hw)
            keep = pyr
            a.gmaps, a.uv, a.row_offsets = C.addressof(pyr), uv.data_ptr(), _ptr(row_offsets)
            a.batch, a.rows_per_sample, a.b_s0 = int(batch), int(rows_per_sample), bias0.data_ptr()
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() >= rows
    h3 = packed.struct_h3
    for l in range(4):
        a.w[l], a.ldw[l], a.b[l] = h3.w[l][1], h3.ldw[l], h3.b[l]
    a.w4, a.b4 = h3.w4, h3.b4
    a.rows, a.clamp, a.out_sdf = rows, float(clamp), out.data_ptr()
    _count(1)
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib.hoisdf_sdf_chain_fwd(C.byref(a), _stream()), "hoisdf_sdf_chain_fwd")
    del keep
    if PROFILE is not None:
        e1.record()
        flops = SDF_DECODER_FLOPS + (2.0 * 512 * 256 if x is None else 0.0)
        PROFILE.append(("sdf_chain", flops * rows, e0, e1, "sdf_chain rows=%d %s single" % (
            rows, "decoder" if x is not None else ("rows" if a0 is not None else "gather"))))
    return out


def sdf_pad_input(x: torch.Tensor):
    x = _f32c(x, "SDFDecoder input")
    rows = x.shape[0]
    if use_h3():
        buf = SplitRows(torch.zeros(rows, 2, ROWH_LD, device=x.device, dtype=torch.float16), ROWH_LD)
        split_rows(x, out=buf.window(0, DEC_IN), kpad=DEC_IN_PAD)
        return buf
    buf = torch.empty(rows, ROW_LD, device=x.device, dtype=torch.float32)
    _count(1)
    check(lib.hoisdf_sdf_pad_input(x.data_ptr(), rows, buf.data_ptr(), ROW_LD, _stream()), "hoisdf_sdf_pad_input")
    return buf


def select_points(sdf, offsets, cand_index, batch: int, num_points: int, bins: int, clamp: float,
                  order_by_row: bool = False):
    dev = sdf.device
    sel = torch.empty(batch, num_points, device=dev, dtype=torch.int32)
    row = torch.empty(batch, num_points, device=dev, dtype=torch.int32)
    pts = torch.empty(batch, num_points, 3, device=dev, dtype=torch.float32)
    out_sdf = torch.empty(batch, num_points, 1, device=dev, dtype=torch.float32)
    pe = torch.empty(batch, num_points, 30, device=dev, dtype=torch.float32)
    flag = torch.zeros(1, device=dev, dtype=torch.int32)
    _count(1)
    check(lib.hoisdf_select_points(sdf.data_ptr(), offsets.data_ptr(), cand_index.data_ptr(),
                                   batch, num_points, bins, float(clamp), int(order_by_row), sel.data_ptr(),
                                   row.data_ptr(), pts.data_ptr(), out_sdf.data_ptr(), pe.data_ptr(), flag.data_ptr(),
                                   _stream()),
          "hoisdf_select_points")
    return sel, pts, out_sdf, pe, flag, row


def tokens(xyz, pe, fea, sdf, beta, out_tokens: torch.Tensor, t0: int):
    b, p, _ = xyz.shape
    assert xyz.is_contiguous() and pe.is_contiguous() and sdf.is_contiguous() and out_tokens.is_contiguous()
    assert fea.stride(-1) == 1 and fea.stride(0) == p * fea.stride(-2)
    _count(1)
    check(lib.hoisdf_tokens_fwd(xyz.data_ptr(), pe.data_ptr(), fea.data_ptr(), fea.stride(-2), sdf.data_ptr(),
                                beta.data_ptr(), b, p, out_tokens.data_ptr(), out_tokens.shape[1], t0, _stream()),
          "hoisdf_tokens_fwd")


# ----------------------------------------------------------------------------------------------------
# Transformer pieces
# ----------------------------------------------------------------------------------------------------
_attn_ws = {}


def _attention_workspace(device, nbytes: int) -> torch.Tensor:
    """One grow-only scratch buffer per device for the bf16 hi/lo operand copies of the tensor-core attention."""
    if torch.cuda.is_current_stream_capturing():
        # CUDA-graph capture: the graph keeps using whatever address it captured, so it gets a buffer of its own from
        # the graph's memory pool instead of the grow-only shared one (which a later, larger call would replace)
        return torch.empty(nbytes, device=device, dtype=torch.uint8)
    buf = _attn_ws.get(device)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, device=device, dtype=torch.uint8)
        _attn_ws[device] = buf
    return buf


def attention(q, ldq, k, v, ldk, out, ldo, batch, heads, lq, lk, kv_valid=None, mask=None, tensor_cores=None):
    """q/k/v/out are tensors whose data_ptr is the first element of head 0 (may be column-offset views).
    tensor_cores: None = automatic (long query sequences), True = also for short ones (decoder cross-attention)."""
    ws, ws_bytes, n = None, 0, 1
    if USE_TENSOR_CORES and mask is None and (lq > 32 or tensor_cores):
        dev = out.buf.device if isinstance(out, SplitRows) else out.device
        ws_bytes = lib.hoisdf_attention_workspace_bytes(batch, heads, lq, lk)
        ws = _attention_workspace(dev, ws_bytes)
        n = 4
    _count(n)
    if isinstance(out, SplitRows):          # tensor-core path only: result straight in split-half format
        if ws is None:
            raise RuntimeError("split-half attention output needs the tensor-core attention kernel")
        check(lib.hoisdf_attention_split_fwd(q.data_ptr(), ldq, k.data_ptr(), v.data_ptr(), ldk, out.hi_ptr, out.lo_ptr,
                                             out.ld, batch, heads, lq, lk, lk if kv_valid is None else kv_valid,
                                             _ptr(ws), ws_bytes, _stream()), "hoisdf_attention_split_fwd")
        return out
    check(lib.hoisdf_attention_fwd(q.data_ptr(), ldq, k.data_ptr(), v.data_ptr(), ldk, out.data_ptr(), ldo, batch,
                                   heads, lq, lk, lk if kv_valid is None else kv_valid, _ptr(mask), _ptr(ws), ws_bytes,
                                   _stream()),
          "hoisdf_attention_fwd")
    return out


def add_layernorm(x, res, gamma, beta, out=None, gamma2=None, beta2=None, out2=None,
                  out_split: Optional[SplitRows] = None, out2_split: Optional[SplitRows] = None):
    """LayerNorm(x (+ res)); `out_split` / `out2_split` additionally receive the result(s) in split-half format."""
    rows = x.numel() // x.shape[-1]
    d = x.shape[-1]
    out = out if out is not None else torch.empty_like(x)
    _count(1)
    if out_split is None and out2_split is None:
        check(lib.hoisdf_add_layernorm_fwd(x.data_ptr(), _ptr(res), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(),
                                           _ptr(gamma2), _ptr(beta2), _ptr(out2), rows, d, _stream()),
              "hoisdf_add_layernorm_fwd")
        return out
    sp = lambda t: (None, None, 0) if t is None else (t.hi_ptr, t.lo_ptr, t.ld)  # noqa: E731
    for t in (out_split, out2_split):
        assert t is None or (t.rows >= rows and t.cols >= d)
    check(lib.hoisdf_add_layernorm_split_fwd(x.data_ptr(), _ptr(res), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(),
                                             _ptr(gamma2), _ptr(beta2), _ptr(out2), rows, d, *sp(out_split),
                                             *sp(out2_split), _stream()), "hoisdf_add_layernorm_split_fwd")
    return out


def vote_joints(points, off, cls) -> torch.Tensor:
    """points (B,P,3), off (L,B,P,60), cls (L,B,P,20) contiguous -> (L,B,20,3)."""
    l, b, p, _ = cls.shape
    out = torch.empty(l, b, 20, 3, device=cls.device, dtype=torch.float32)
    _count(1)
    check(lib.hoisdf_vote_joints_fwd(points.data_ptr(), off.data_ptr(), cls.data_ptr(), l, b, p, out.data_ptr(),
                                     _stream()), "hoisdf_vote_joints_fwd")
    return out


def mano(model_struct, pose6d, betas):
    """pose6d (N,16,6), betas (N,10) contiguous -> verts (N,778,3), joints (N,21,3) [m]."""
    n = pose6d.shape[0]
    verts = torch.empty(n, 778, 3, device=pose6d.device, dtype=torch.float32)
    joints = torch.empty(n, 21, 3, device=pose6d.device, dtype=torch.float32)
    _count(1)
    check(lib.hoisdf_mano_fwd(C.byref(model_struct), pose6d.data_ptr(), betas.data_ptr(), n, verts.data_ptr(),
                              joints.data_ptr(), _stream()), "hoisdf_mano_fwd")
    return verts, joints


def mano_aa(model_struct, pose_aa, betas):
    """pose_aa (N,48) axis-angle, betas (N,10) contiguous -> verts (N,778,3), joints (N,21,3) [m]."""
    n = pose_aa.shape[0]
    verts = torch.empty(n, 778, 3, device=pose_aa.device, dtype=torch.float32)
    joints = torch.empty(n, 21, 3, device=pose_aa.device, dtype=torch.float32)
    _count(1)
    check(lib.hoisdf_mano_aa_fwd(C.byref(model_struct), pose_aa.data_ptr(), betas.data_ptr(), n, verts.data_ptr(),
                                 joints.data_ptr(), _stream()), "hoisdf_mano_aa_fwd")
    return verts, joints


# ----------------------------------------------------------------------------------------------------
# Test-time metrics (upstream common/metrics.py)
# ----------------------------------------------------------------------------------------------------
def _metric_workspace(batch: int, n_verts: int, dev) -> torch.Tensor:
    nbytes = int(lib.hoisdf_obj_metrics_workspace_bytes(batch, n_verts))
    return torch.empty(max(nbytes // 4, 1), device=dev, dtype=torch.float32)


def obj_pose_metrics(templates, obj_ids, rot_pred, trans_pred, rot_gt, trans_gt):
    """templates (T,N,3), obj_ids (B) int64 or None, rot_pred / trans_pred (B,P,3) per-point votes, rot_gt / trans_gt
    (B,3) -> (adds, mme, mce, oce), each (B).  upstream common/metrics.py:110-185."""
    templates = _f32c(templates, "templates")
    rot_pred, trans_pred = _f32c(rot_pred, "rot_pred"), _f32c(trans_pred, "trans_pred")
    rot_gt, trans_gt = _f32c(rot_gt, "rot_gt"), _f32c(trans_gt, "trans_gt")
    if rot_pred.dim() == 2:                                   # already one pose per sample
        rot_pred, trans_pred = rot_pred[:, None], trans_pred[:, None]
    b, votes = rot_pred.shape[0], rot_pred.shape[1]
    t, n = templates.shape[0], templates.shape[1]
    if rot_pred.shape != trans_pred.shape or rot_gt.shape != (b, 3) or trans_gt.shape != (b, 3) or \
            templates.dim() != 3 or templates.shape[2] != 3 or rot_pred.shape[2] != 3:
        raise ValueError("obj_pose_metrics: inconsistent shapes")
    if obj_ids is not None:
        if obj_ids.dtype != torch.int64 or not obj_ids.is_cuda or obj_ids.shape != (b,):
            raise ValueError("obj_ids must be a CUDA int64 tensor of shape (B,)")
        obj_ids = obj_ids.contiguous()
    out = torch.empty(4, b, device=templates.device, dtype=torch.float32)
    ws = _metric_workspace(b, n, templates.device)
    _count(2)
    check(lib.hoisdf_obj_metrics_fwd(templates.data_ptr(), _ptr(obj_ids), t, n, rot_pred.data_ptr(),
                                     trans_pred.data_ptr(), votes, rot_gt.data_ptr(), trans_gt.data_ptr(), b,
                                     out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(),
                                     ws.data_ptr(), ws.numel() * 4, _stream()), "hoisdf_obj_metrics_fwd")
    return out[0], out[1], out[2], out[3]


def mesh_metrics(pred_meshes, target_meshes):
    """(B,N,3) x2 -> (adds, mme, mce), each (B).  upstream common/metrics.py:62-108."""
    pred_meshes, target_meshes = _f32c(pred_meshes, "pred_meshes"), _f32c(target_meshes, "target_meshes")
    if pred_meshes.shape != target_meshes.shape or pred_meshes.dim() != 3 or pred_meshes.shape[2] != 3:
        raise ValueError("mesh_metrics: meshes must both be (B, N, 3)")
    b, n = pred_meshes.shape[0], pred_meshes.shape[1]
    out = torch.empty(3, b, device=pred_meshes.device, dtype=torch.float32)
    ws = _metric_workspace(b, n, pred_meshes.device)
    _count(2)
    check(lib.hoisdf_mesh_metrics_fwd(pred_meshes.data_ptr(), target_meshes.data_ptr(), b, n, out[0].data_ptr(),
                                      out[1].data_ptr(), out[2].data_ptr(), ws.data_ptr(), ws.numel() * 4, _stream()),
          "hoisdf_mesh_metrics_fwd")
    return out[0], out[1], out[2]


def hand_joint_metrics(pred, gt, want_aligned: bool = False):
    """pred / gt (B,J,3) -> (mje (B), pamje (B)[, aligned (B,J,3)]).  upstream common/metrics.py:188-248."""
    pred, gt = _f32c(pred, "pred"), _f32c(gt, "gt")
    if pred.shape != gt.shape or pred.dim() != 3 or pred.shape[2] != 3:
        raise ValueError("hand_joint_metrics: pred and gt must both be (B, J, 3)")
    b, j = pred.shape[0], pred.shape[1]
    out = torch.empty(2, b, device=pred.device, dtype=torch.float32)
    aligned = torch.empty_like(pred) if want_aligned else None
    _count(1)
    check(lib.hoisdf_hand_joint_metrics_fwd(pred.data_ptr(), gt.data_ptr(), b, j, out[0].data_ptr(), out[1].data_ptr(),
                                            _ptr(aligned), _stream()), "hoisdf_hand_joint_metrics_fwd")
    return (out[0], out[1], aligned) if want_aligned else (out[0], out[1])
