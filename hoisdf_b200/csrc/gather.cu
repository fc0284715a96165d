// Multi-scale bilinear feature gather (upstream F.grid_sample(bilinear, border, align_corners=True) x 5
// levels + cat, main/model.py:164-175 / 203-214 / 316-328) on channels-last maps, and the NCHW->NHWC copy.
//
// HBM/L2-bound work: every tap of a point is one contiguous C_l-vector in NHWC, read with 128-bit loads
// by one warp (lanes stride the channel quads), so all global traffic is fully coalesced.  Consecutive
// rows of one sample are lattice neighbours along z and project to (almost) the same pixel, so the
// 8 warps of a CTA hit the same lines in L1.
#include "tc_common.cuh"

namespace hoisdf {

struct GatherParams {
  hoisdf_pyramid pyr;
  const float* __restrict__ uv;
  const int64_t* __restrict__ row_offsets;
  const float* __restrict__ bias;
  float* __restrict__ out;
  __half* __restrict__ out_hi;   // split-half output planes (tc_common.cuh) instead of `out` when not NULL
  __half* __restrict__ out_lo;
  int64_t rows, batch, rows_per_sample, ld_out;
  int act;
};

// store 4 consecutive columns of row `r` starting at column c, in the output format the caller asked for
__device__ __forceinline__ void store4(const GatherParams& p, int64_t r, int c, float4 a) {
  if (p.out_hi == nullptr) {
    *reinterpret_cast<float4*>(p.out + r * p.ld_out + c) = a;
    return;
  }
  __half h[4], l[4];
  tc::split_half(a.x, h[0], l[0]); tc::split_half(a.y, h[1], l[1]);
  tc::split_half(a.z, h[2], l[2]); tc::split_half(a.w, h[3], l[3]);
  uint2 ph, pl;
  ph.x = static_cast<uint32_t>(__half_as_ushort(h[0])) | (static_cast<uint32_t>(__half_as_ushort(h[1])) << 16);
  ph.y = static_cast<uint32_t>(__half_as_ushort(h[2])) | (static_cast<uint32_t>(__half_as_ushort(h[3])) << 16);
  pl.x = static_cast<uint32_t>(__half_as_ushort(l[0])) | (static_cast<uint32_t>(__half_as_ushort(l[1])) << 16);
  pl.y = static_cast<uint32_t>(__half_as_ushort(l[2])) | (static_cast<uint32_t>(__half_as_ushort(l[3])) << 16);
  *reinterpret_cast<uint2*>(p.out_hi + r * p.ld_out + c) = ph;
  *reinterpret_cast<uint2*>(p.out_lo + r * p.ld_out + c) = pl;
}

struct Taps {
  int64_t o00, o01, o10, o11;  // element offsets of the 4 taps (channel 0)
  float w00, w01, w10, w11;    // nw, ne, sw, se
};

// ATen grid_sampler_2d arithmetic (align_corners=True, border padding):
//   g = (uv - (img-1)/2) / ((img-1)/2);  x = ((g + 1) / 2) * (W - 1);  x = min(W-1, max(x, 0))
__device__ __forceinline__ Taps make_taps(float u, float v, int img_w, int img_h, int W, int H, int C, int64_t b) {
  const float nx = static_cast<float>(img_w - 1) / 2.0f;
  const float ny = static_cast<float>(img_h - 1) / 2.0f;
  const float gx = __fdiv_rn(__fsub_rn(u, nx), nx);
  const float gy = __fdiv_rn(__fsub_rn(v, ny), ny);
  float x = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), static_cast<float>(W - 1));
  float y = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), static_cast<float>(H - 1));
  x = fminf(static_cast<float>(W - 1), fmaxf(x, 0.f));
  y = fminf(static_cast<float>(H - 1), fmaxf(y, 0.f));
  const float x0f = floorf(x), y0f = floorf(y);
  const float tx = x - x0f, ty = y - y0f;  // exact (Sterbenz)
  const int x0 = static_cast<int>(x0f), y0 = static_cast<int>(y0f);
  const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
  Taps t;
  // weights as ATen forms them: (x_se - x) * (y_se - y) etc. with x_se = x0 + 1
  const float ax = __fsub_rn(x0f + 1.f, x), ay = __fsub_rn(y0f + 1.f, y);
  t.w00 = ax * ay;
  t.w01 = tx * ay;
  t.w10 = ax * ty;
  t.w11 = tx * ty;
  // taps beyond the border are skipped upstream; their weight is exactly 0 here because x was clamped
  if (x0 + 1 > W - 1) { t.w01 = 0.f; t.w11 = 0.f; }
  if (y0 + 1 > H - 1) { t.w10 = 0.f; t.w11 = 0.f; }
  const int64_t base = b * static_cast<int64_t>(H) * W;
  t.o00 = (base + static_cast<int64_t>(y0) * W + x0) * C;
  t.o01 = (base + static_cast<int64_t>(y0) * W + x1) * C;
  t.o10 = (base + static_cast<int64_t>(y1) * W + x0) * C;
  t.o11 = (base + static_cast<int64_t>(y1) * W + x1) * C;
  return t;
}

__device__ __forceinline__ int64_t sample_of_row(const GatherParams& p, int64_t r) {
  if (p.row_offsets == nullptr) return r / p.rows_per_sample;
  int64_t lo = 0, hi = p.batch;  // largest b with offsets[b] <= r
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (p.row_offsets[mid] <= r) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float4 blend(const float* m, const Taps& t, int c) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(m + t.o00 + c));
  const float4 b = __ldg(reinterpret_cast<const float4*>(m + t.o01 + c));
  const float4 d = __ldg(reinterpret_cast<const float4*>(m + t.o10 + c));
  const float4 e = __ldg(reinterpret_cast<const float4*>(m + t.o11 + c));
  float4 r;
  r.x = fmaf(e.x, t.w11, fmaf(d.x, t.w10, fmaf(b.x, t.w01, a.x * t.w00)));
  r.y = fmaf(e.y, t.w11, fmaf(d.y, t.w10, fmaf(b.y, t.w01, a.y * t.w00)));
  r.z = fmaf(e.z, t.w11, fmaf(d.z, t.w10, fmaf(b.z, t.w01, a.z * t.w00)));
  r.w = fmaf(e.w, t.w11, fmaf(d.w, t.w10, fmaf(b.w, t.w01, a.w * t.w00)));
  return r;
}

// one warp per row; CONCAT: out[r, off_l + c] = sample_l[c]
__global__ void __launch_bounds__(256) gather_concat_kernel(const GatherParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= p.rows) return;
  const int64_t b = sample_of_row(p, r);
  const float u = p.uv[r * 2 + 0], v = p.uv[r * 2 + 1];
  int off = 0;
#pragma unroll 1
  for (int l = 0; l < p.pyr.levels; ++l) {
    const int C = p.pyr.c[l];
    const Taps t = make_taps(u, v, p.pyr.img_w, p.pyr.img_h, p.pyr.w[l], p.pyr.h[l], C, b);
    const float* m = p.pyr.map[l];
    for (int c = lane * 4; c < C; c += 128) store4(p, r, off + c, blend(m, t, c));
    off += C;
  }
}

// one warp per row; SUM: out[r, c] = act(bias[c] + sum_l sample_l[c]); C <= 512 and C % 128 == 0
template <int CQ>  // float4 per lane
__global__ void __launch_bounds__(256) gather_sum_kernel(const GatherParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= p.rows) return;
  const int64_t b = sample_of_row(p, r);
  const float u = p.uv[r * 2 + 0], v = p.uv[r * 2 + 1];
  const int C = CQ * 128;
  float4 acc[CQ];
#pragma unroll
  for (int q = 0; q < CQ; ++q) {
    acc[q] = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + q * 128 + lane * 4))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll 1
  for (int l = 0; l < p.pyr.levels; ++l) {
    const Taps t = make_taps(u, v, p.pyr.img_w, p.pyr.img_h, p.pyr.w[l], p.pyr.h[l], C, b);
    const float* m = p.pyr.map[l];
#pragma unroll
    for (int q = 0; q < CQ; ++q) {
      const float4 s = blend(m, t, q * 128 + lane * 4);
      acc[q].x += s.x; acc[q].y += s.y; acc[q].z += s.z; acc[q].w += s.w;
    }
  }
#pragma unroll
  for (int q = 0; q < CQ; ++q) {
    float4 a = acc[q];
    if (p.act == HOISDF_ACT_RELU) {
      a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
    }
    store4(p, r, q * 128 + lane * 4, a);
  }
}

// SUM mode on fp16 maps with an fp16 (hi plane only) result: the candidate SCREENING path (single-product chain kernel,
// csrc/sdf_chain.cu, which reads nothing but the hi plane).  Half the L1 / L2 bytes per tap of the fp32 kernel above and a
// quarter of its output bytes.  One warp per row, C = 512: every lane owns 16 channels (2 x 16 bytes per tap).
struct GatherH16Params {
  const uint4* __restrict__ map[5];     // (B, H_l, W_l, 512) halfs
  int h[5], w[5];
  int levels, img_w, img_h;
  const float* __restrict__ uv;
  const int64_t* __restrict__ row_offsets;
  const float* __restrict__ bias;
  __half* __restrict__ out_hi;
  int64_t rows, batch, rows_per_sample, ld_out;
  int act;
};

__global__ void __launch_bounds__(256) gather_sum_h16_kernel(const GatherH16Params p) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= p.rows) return;
  int64_t b;
  if (p.row_offsets == nullptr) {
    b = r / p.rows_per_sample;
  } else {
    int64_t lo = 0, hi = p.batch;
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (p.row_offsets[mid] <= r) lo = mid; else hi = mid;
    }
    b = lo;
  }
  const float u = p.uv[r * 2 + 0], v = p.uv[r * 2 + 1];
  float acc[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 bb = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + lane * 16) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    acc[4 * q] = bb.x; acc[4 * q + 1] = bb.y; acc[4 * q + 2] = bb.z; acc[4 * q + 3] = bb.w;
  }
#pragma unroll
  for (int l = 0; l < 5; ++l) {
    if (l >= p.levels) break;
    const Taps t = make_taps(u, v, p.img_w, p.img_h, p.w[l], p.h[l], 64, b);     // offsets in uint4 units (64 per pixel)
    const uint4* m = p.map[l] + lane * 2;
#pragma unroll
    for (int hseg = 0; hseg < 2; ++hseg) {
      const uint4 a = __ldg(m + t.o00 + hseg), c = __ldg(m + t.o01 + hseg), d = __ldg(m + t.o10 + hseg), e = __ldg(m + t.o11 + hseg);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, cw[4] = {c.x, c.y, c.z, c.w}, dw[4] = {d.x, d.y, d.z, d.w},
                     ew[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&aw[j]));
        const float2 fc = __half22float2(*reinterpret_cast<const __half2*>(&cw[j]));
        const float2 fd = __half22float2(*reinterpret_cast<const __half2*>(&dw[j]));
        const float2 fe = __half22float2(*reinterpret_cast<const __half2*>(&ew[j]));
        acc[hseg * 8 + 2 * j] += fmaf(fe.x, t.w11, fmaf(fd.x, t.w10, fmaf(fc.x, t.w01, fa.x * t.w00)));
        acc[hseg * 8 + 2 * j + 1] += fmaf(fe.y, t.w11, fmaf(fd.y, t.w10, fmaf(fc.y, t.w01, fa.y * t.w00)));
      }
    }
  }
  uint32_t o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float x0 = acc[2 * j], x1 = acc[2 * j + 1];
    if (p.act == HOISDF_ACT_RELU) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
    o[j] = tc::cvt_f16x2_sat(x0, x1);
  }
  uint4* dst = reinterpret_cast<uint4*>(p.out_hi + r * p.ld_out + lane * 16);
  dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
  dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// (N, C, HW) -> (N, HW, C) through a 32x33 shared tile
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                           int C, int HW) {
  __shared__ float tile[32][33];
  const int64_t n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* s = src + n * static_cast<int64_t>(C) * HW;
  float* d = dst + n * static_cast<int64_t>(C) * HW;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int c = c0 + ty + j, pp = p0 + tx;
    tile[ty + j][tx] = (c < C && pp < HW) ? s[static_cast<int64_t>(c) * HW + pp] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int pp = p0 + ty + j, c = c0 + tx;
    if (c < C && pp < HW) d[static_cast<int64_t>(pp) * C + c] = tile[tx][ty + j];
  }
}

// (N, C, HW) fp32 -> split-half (N, HW, ld) planes at a channel offset, through a 32x33 shared tile
__global__ void __launch_bounds__(256) nchw_to_nhwc_split_kernel(const float* __restrict__ src, __half* __restrict__ hi,
                                                                 __half* __restrict__ lo, int C, int HW, int64_t ld) {
  __shared__ float tile[32][33];
  const int64_t n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* s = src + n * static_cast<int64_t>(C) * HW;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int c = c0 + ty + j, pp = p0 + tx;
    tile[ty + j][tx] = (c < C && pp < HW) ? s[static_cast<int64_t>(c) * HW + pp] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int pp = p0 + ty + j, c = c0 + tx;
    if (c < C && pp < HW) {
      __half h, l;
      tc::split_half(tile[tx][ty + j], h, l);
      const int64_t o = (n * HW + pp) * ld + c;
      hi[o] = h;
      lo[o] = l;
    }
  }
}

}  // namespace hoisdf

using namespace hoisdf;

static int gather_any(const hoisdf_pyramid* pyr, const float* uv, int64_t rows, const int64_t* row_offsets,
                      int64_t batch, int64_t rows_per_sample, int32_t mode, const float* bias, int32_t act, float* out,
                      uint16_t* out_hi, uint16_t* out_lo, int64_t ld_out, void* stream) {
  const bool split = out_hi != nullptr || out_lo != nullptr;
  if (pyr == nullptr || uv == nullptr) return HOISDF_E_NULL;
  if (split ? (out_hi == nullptr || out_lo == nullptr) : out == nullptr) return HOISDF_E_NULL;
  if (rows == 0) return HOISDF_OK;
  if (rows < 0 || batch <= 0 || pyr->levels < 1 || pyr->levels > 5) return HOISDF_E_SHAPE;
  if (row_offsets == nullptr && rows_per_sample <= 0) return HOISDF_E_SHAPE;
  if (split) {
    if ((reinterpret_cast<uintptr_t>(out_hi) & 7) || (reinterpret_cast<uintptr_t>(out_lo) & 7) || (ld_out & 3))
      return HOISDF_E_ALIGN;
  } else if (!aligned16(out) || (ld_out & 3)) {
    return HOISDF_E_ALIGN;
  }
  int ctot = 0;
  for (int l = 0; l < pyr->levels; ++l) {
    if (pyr->map[l] == nullptr) return HOISDF_E_NULL;
    if (!aligned16(pyr->map[l]) || (pyr->c[l] & 3)) return HOISDF_E_ALIGN;
    if (pyr->c[l] <= 0 || pyr->h[l] <= 0 || pyr->w[l] <= 0) return HOISDF_E_SHAPE;
    ctot += pyr->c[l];
  }
  GatherParams p;
  p.pyr = *pyr; p.uv = uv; p.row_offsets = row_offsets; p.bias = bias; p.out = out;
  p.out_hi = reinterpret_cast<__half*>(out_hi); p.out_lo = reinterpret_cast<__half*>(out_lo);
  p.rows = rows; p.batch = batch; p.rows_per_sample = rows_per_sample; p.ld_out = ld_out; p.act = act;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>(ceil_div(rows, 8));
  if (mode == HOISDF_GATHER_CONCAT) {
    if (ld_out < ctot) return HOISDF_E_SHAPE;
    HOISDF_LAUNCH(gather_concat_kernel, grid, 256, s, p);
  } else if (mode == HOISDF_GATHER_SUM) {
    const int C = pyr->c[0];
    for (int l = 1; l < pyr->levels; ++l)
      if (pyr->c[l] != C) return HOISDF_E_SHAPE;
    if (ld_out < C) return HOISDF_E_SHAPE;
    if (bias != nullptr && !aligned16(bias)) return HOISDF_E_ALIGN;
    if (C == 512) HOISDF_LAUNCH(gather_sum_kernel<4>, grid, 256, s, p);
    else if (C == 256) HOISDF_LAUNCH(gather_sum_kernel<2>, grid, 256, s, p);
    else if (C == 128) HOISDF_LAUNCH(gather_sum_kernel<1>, grid, 256, s, p);
    else return HOISDF_E_UNSUPPORTED;
  } else {
    return HOISDF_E_UNSUPPORTED;
  }
  return launch_status();
}

HOISDF_API int hoisdf_gather_fwd(const hoisdf_pyramid* pyr, const float* uv, int64_t rows,
                                 const int64_t* row_offsets, int64_t batch, int64_t rows_per_sample, int32_t mode,
                                 const float* bias, int32_t act, float* out, int64_t ld_out, void* stream) {
  if (out == nullptr) return HOISDF_E_NULL;
  return gather_any(pyr, uv, rows, row_offsets, batch, rows_per_sample, mode, bias, act, out, nullptr, nullptr, ld_out,
                    stream);
}

HOISDF_API int hoisdf_gather_split_fwd(const hoisdf_pyramid* pyr, const float* uv, int64_t rows,
                                       const int64_t* row_offsets, int64_t batch, int64_t rows_per_sample,
                                       int32_t mode, const float* bias, int32_t act, uint16_t* out_hi,
                                       uint16_t* out_lo, int64_t ld_out, void* stream) {
  if (out_hi == nullptr || out_lo == nullptr) return HOISDF_E_NULL;
  return gather_any(pyr, uv, rows, row_offsets, batch, rows_per_sample, mode, bias, act, nullptr, out_hi, out_lo,
                    ld_out, stream);
}

HOISDF_API int hoisdf_nchw_to_nhwc(const float* src, float* dst, int64_t n, int64_t c, int64_t h, int64_t w,
                                   void* stream) {
  if (src == nullptr || dst == nullptr) return HOISDF_E_NULL;
  if (n <= 0 || c <= 0 || h <= 0 || w <= 0 || n > 65535 || c > (1 << 20) || h * w > (1 << 24)) return HOISDF_E_SHAPE;
  const int HW = static_cast<int>(h * w);
  dim3 grid(static_cast<unsigned>(ceil_div(HW, 32)), static_cast<unsigned>(ceil_div(c, 32)),
            static_cast<unsigned>(n));
  HOISDF_LAUNCH(nchw_to_nhwc_kernel, grid, 256, static_cast<cudaStream_t>(stream), src, dst, static_cast<int>(c), HW);
  return launch_status();
}

HOISDF_API int hoisdf_nchw_to_nhwc_split(const float* src, uint16_t* dst_hi, uint16_t* dst_lo, int64_t n, int64_t c,
                                         int64_t h, int64_t w, int64_t ld, void* stream) {
  if (src == nullptr || dst_hi == nullptr || dst_lo == nullptr) return HOISDF_E_NULL;
  if (n <= 0 || c <= 0 || h <= 0 || w <= 0 || n > 65535 || c > (1 << 20) || h * w > (1 << 24) || ld < c)
    return HOISDF_E_SHAPE;
  const int HW = static_cast<int>(h * w);
  dim3 grid(static_cast<unsigned>(ceil_div(HW, 32)), static_cast<unsigned>(ceil_div(c, 32)),
            static_cast<unsigned>(n));
  HOISDF_LAUNCH(nchw_to_nhwc_split_kernel, grid, 256, static_cast<cudaStream_t>(stream), src,
                reinterpret_cast<__half*>(dst_hi), reinterpret_cast<__half*>(dst_lo), static_cast<int>(c), HW, ld);
  return launch_status();
}

HOISDF_API int hoisdf_gather_sum_h16_fwd(const hoisdf_pyramid_h* pyr, const float* uv, int64_t rows, const int64_t* row_offsets,
                                         int64_t batch, int64_t rows_per_sample, const float* bias, int32_t act,
                                         uint16_t* out_hi, int64_t ld_out, void* stream) {
  if (pyr == nullptr || uv == nullptr || out_hi == nullptr) return HOISDF_E_NULL;
  if (rows == 0) return HOISDF_OK;
  if (rows < 0 || batch <= 0 || pyr->levels < 1 || pyr->levels > 5 || pyr->c != 512 || ld_out < 512) return HOISDF_E_SHAPE;
  if (row_offsets == nullptr && rows_per_sample <= 0) return HOISDF_E_SHAPE;
  if (!aligned16(out_hi) || (ld_out & 7) || (bias != nullptr && !aligned16(bias))) return HOISDF_E_ALIGN;
  GatherH16Params p;
  for (int l = 0; l < 5; ++l) {
    const int ll = l < pyr->levels ? l : 0;
    if (pyr->map[ll] == nullptr) return HOISDF_E_NULL;
    if (!aligned16(pyr->map[ll])) return HOISDF_E_ALIGN;
    if (pyr->h[ll] <= 0 || pyr->w[ll] <= 0) return HOISDF_E_SHAPE;
    p.map[l] = reinterpret_cast<const uint4*>(pyr->map[ll]);
    p.h[l] = pyr->h[ll];
    p.w[l] = pyr->w[ll];
  }
  p.levels = pyr->levels; p.img_w = pyr->img_w; p.img_h = pyr->img_h;
  p.uv = uv; p.row_offsets = row_offsets; p.bias = bias; p.out_hi = reinterpret_cast<__half*>(out_hi);
  p.rows = rows; p.batch = batch; p.rows_per_sample = rows_per_sample; p.ld_out = ld_out; p.act = act;
  HOISDF_LAUNCH(gather_sum_h16_kernel, static_cast<unsigned>(ceil_div(rows, 8)), 256, static_cast<cudaStream_t>(stream), p);
  return launch_status();
}
