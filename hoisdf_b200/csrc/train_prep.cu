// Operand preparation for the tensor-core backward of a Linear (hoisdf_b200/autograd.py:LinearFn.backward; upstream
// main/train.py:131 `loss.backward()` through every nn.Linear of the hot path).  The backward runs the SAME FP16x3 GEMM as the
// forward on transposed operands:
//      dX   = dZ . W        = linear_h3( split(dZ / s),  pack(W^T) ) * s
//      dW^T = X^T . dZ      = linear_h3( split(X^T),     pack((dZ / s)^T) ) ,   dW = (dW^T)^T * s
// with dZ = dY * relu'(Y) and s a power of two that brings the gradient into the fp16 planes' range.  Done with PyTorch glue
// that is 8 passes over dY (clone, mask + bias sums, |.|, max, scale, two transposed copies, split, pack); here it is
//   hoisdf_absmax            max |dY|                                                      (1 read)
//   hoisdf_linear_bwd_prep   dZ = dY * [Y > 0];  s = 2^(ceil(log2(max|dY|)) - 3) on the device;  one pass writes
//                            dZ / s in split-half format (x operand of the dX GEMM), (dZ / s)^T as the three weight planes
//                            (w operand of the dW GEMM, tile transpose through shared memory) and adds the column sums
//                            into db                                                        (1 read of dY and Y)
//   hoisdf_split_rows_t      X (M, K) fp32 -> X^T in split-half format (K rows of M)        (1 read)
// and, so that a layer costs the host ONE call per direction instead of 3 / 7 (the 17-query decoder layers and the small heads
// are bound by the host's launch rate, not by the GPU):
//   hoisdf_linear_train_fwd  split(X), pack(W), GEMM                        -> Y
//   hoisdf_linear_train_bwd  absmax, prep, pack(W^T) + GEMM -> dX, X^T + split-K GEMM -> dW^T, db
// on a caller-owned workspace (hoisdf_linear_train_workspace_bytes).
#include <cstring>

#include "tc_common.cuh"

namespace hoisdf {
using namespace tc;
namespace {

__global__ void absmax_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld, unsigned* __restrict__ out) {
  float m = 0.f;
  const int64_t n = rows * cols;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    m = fmaxf(m, fabsf(x[r * ld + c]));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));     // non-negative floats order like their bits
}

__device__ __forceinline__ float pow2_scale(float amax) {
  // s = 2^(ceil(log2(amax)) - 3): amax / s in (4, 8]; 1 for an all-zero (or non-finite) gradient
  if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
  int e;
  const float f = frexpf(amax, &e);          // amax = f * 2^e, f in [0.5, 1)
  const int c = (f == 0.5f) ? e - 1 : e;     // ceil(log2(amax))
  return ldexpf(1.f, c - 3);
}

// 32 x 32 tile per block (32 x 8 threads, 4 rows each)
__global__ void __launch_bounds__(256)
linear_bwd_prep_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ y, int64_t ldy, int64_t m, int64_t n,
                       int act, const unsigned* __restrict__ amax_bits, __half* __restrict__ dz_hi, __half* __restrict__ dz_lo,
                       int64_t ld_dz, __half* __restrict__ ta, __half* __restrict__ tb, __half* __restrict__ tcp, int64_t ld_t,
                       float* __restrict__ db, float* __restrict__ scale_out) {
  __shared__ float tile[32][33];
  __shared__ float colsum[8][33];
  const float s = pow2_scale(__uint_as_float(*amax_bits));
  const float inv = 1.f / s;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) *scale_out = s;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 32, c0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t c = c0 + tx;
  float cs = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rl = ty + 8 * i;
    const int64_t r = r0 + rl;
    float v = 0.f;
    if (r < m && c < n) {
      v = dy[r * lddy + c];
      if (act == HOISDF_ACT_RELU && !(y[r * ldy + c] > 0.f)) v = 0.f;
      cs += v;
      const float vs = v * inv;
      __half h, l;
      split_half(vs, h, l);
      dz_hi[r * ld_dz + c] = h;
      dz_lo[r * ld_dz + c] = l;
      v = vs;
    }
    tile[rl][tx] = v;
  }
  colsum[ty][tx] = cs;
  __syncthreads();
  if (db != nullptr && ty == 0 && c < n) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += colsum[i][tx];
    atomicAdd(db + c, t);
  }
  // transposed write: thread (tx, ty) now owns row index tx of the tile (= column of the transposed matrix)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cl = ty + 8 * i;                       // column of the tile = row of the transposed matrix
    const int64_t cc = c0 + cl, rr = r0 + tx;
    if (cc < n && rr < m) {
      const float w = tile[tx][cl];
      // hoisdf_pack_h3's plane format: A = fp16(w * 2^11), B = fp16(A * 2^-11), C = fp16((w - A * 2^-11) * 2^11)
      const __half ha = __float2half_rn(fminf(fmaxf(w * kLoScale, -65504.f), 65504.f));
      const float whi = __half2float(ha) * kLoInv;
      ta[cc * ld_t + rr] = ha;
      tb[cc * ld_t + rr] = __float2half_rn(whi);
      tcp[cc * ld_t + rr] = __float2half_rn(fminf(fmaxf((w - whi) * kLoScale, -65504.f), 65504.f));
    }
  }
}

__global__ void __launch_bounds__(256)
split_rows_t_kernel(const float* __restrict__ x, int64_t m, int64_t k, int64_t ldx, __half* __restrict__ hi,
                    __half* __restrict__ lo, int64_t ldh) {
  __shared__ float tile[32][33];
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 32, c0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rl = ty + 8 * i;
    const int64_t r = r0 + rl, c = c0 + tx;
    tile[rl][tx] = (r < m && c < k) ? x[r * ldx + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cl = ty + 8 * i;
    const int64_t cc = c0 + cl, rr = r0 + tx;
    if (cc < k && rr < m) {
      __half h, l;
      split_half(tile[tx][cl], h, l);
      hi[cc * ldh + rr] = h;
      lo[cc * ldh + rr] = l;
    }
  }
}


// W (n, k) fp32 -> hoisdf_pack_h3 planes of W^T: (k rows, ldh >= n halfs); 32 x 32 tile per block (32 x 8 threads)
__global__ void __launch_bounds__(256)
pack_h3_t_kernel(const float* __restrict__ w, int64_t n, int64_t k, int64_t ldw, __half* __restrict__ a, __half* __restrict__ b,
                 __half* __restrict__ c, int64_t ldh) {
  __shared__ float tile[32][33];
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 32, c0 = static_cast<int64_t>(blockIdx.x) * 32;   // rows of w, cols of w
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rl = ty + 8 * i;
    const int64_t r = r0 + rl, cc = c0 + tx;
    tile[rl][tx] = (r < n && cc < k) ? w[r * ldw + cc] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cl = ty + 8 * i;
    const int64_t orow = c0 + cl, ocol = r0 + tx;              // element (orow, ocol) of W^T
    if (orow < k && ocol < n) {
      const float v = tile[tx][cl];
      const __half ha = __float2half_rn(fminf(fmaxf(v * kLoScale, -65504.f), 65504.f));
      const float whi = __half2float(ha) * kLoInv;
      a[orow * ldh + ocol] = ha;
      b[orow * ldh + ocol] = __float2half_rn(whi);
      c[orow * ldh + ocol] = __float2half_rn(fminf(fmaxf((v - whi) * kLoScale, -65504.f), 65504.f));
    }
  }
}

inline int64_t up8(int64_t v) { return (v + 7) / 8 * 8; }
inline int64_t up256(int64_t v) { return (v + 255) / 256 * 256; }

struct TrainWs {                       // byte offsets into the workspace
  int64_t scal, xs, wp, dz, dzt, wt, xt, total;
};
inline TrainWs train_ws(int64_t m, int64_t n, int64_t k) {
  TrainWs w;
  int64_t o = 0;
  w.scal = o; o += 256;                                   // amax | scale
  w.xs = o;   o += up256(2 * m * up8(k) * 2);             // forward: X split-half (hi plane | lo plane)
  w.wp = o;   o += up256(3 * n * up8(k) * 2);             // forward: W planes
  const int64_t fwd_end = o;
  o = 256;                                                // the backward reuses the same bytes
  w.dz = o;   o += up256(2 * m * up8(n) * 2);             // dZ / s split-half
  w.dzt = o;  o += up256(3 * n * up8(m) * 2);             // (dZ / s)^T planes
  w.wt = o;   o += up256(3 * k * up8(n) * 2);             // W^T planes
  w.xt = o;   o += up256(2 * k * up8(m) * 2);             // X^T split-half
  w.total = o > fwd_end ? o : fwd_end;
  return w;
}

}  // namespace
}  // namespace hoisdf

using namespace hoisdf;

HOISDF_API int hoisdf_absmax(const float* x, int64_t rows, int64_t cols, int64_t ld, float* out, void* stream) {
  if (x == nullptr || out == nullptr) return HOISDF_E_NULL;
  if (rows <= 0 || cols <= 0 || ld < cols) return HOISDF_E_SHAPE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(out, 0, sizeof(float), s) != cudaSuccess) return HOISDF_E_SHAPE;
  const int64_t n = rows * cols;
  const int64_t want = ceil_div(n, 256 * 8);
  const unsigned blocks = static_cast<unsigned>(want < 148 * 16 ? (want < 1 ? 1 : want) : 148 * 16);
  absmax_kernel<<<blocks, 256, 0, s>>>(x, rows, cols, ld, reinterpret_cast<unsigned*>(out));
  return launch_status();
}

HOISDF_API int hoisdf_linear_bwd_prep(const float* dy, int64_t lddy, const float* y, int64_t ldy, int64_t m, int64_t n,
                                      int32_t act, const float* amax, uint16_t* dz_hi, uint16_t* dz_lo, int64_t ld_dz,
                                      uint16_t* dzt_a, uint16_t* dzt_b, uint16_t* dzt_c, int64_t ld_dzt, float* db,
                                      float* scale_out, void* stream) {
  if (dy == nullptr || amax == nullptr || dz_hi == nullptr || dz_lo == nullptr || dzt_a == nullptr || dzt_b == nullptr ||
      dzt_c == nullptr || scale_out == nullptr || (act == HOISDF_ACT_RELU && y == nullptr))
    return HOISDF_E_NULL;
  if (m <= 0 || n <= 0 || lddy < n || ld_dz < n || ld_dzt < m || (y != nullptr && ldy < n)) return HOISDF_E_SHAPE;
  if (act != HOISDF_ACT_NONE && act != HOISDF_ACT_RELU) return HOISDF_E_UNSUPPORTED;
  const int64_t gy = ceil_div(m, 32), gx = ceil_div(n, 32);
  if (gy > 0x7fffffffLL || gx > 65535) return HOISDF_E_SHAPE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (db != nullptr && cudaMemsetAsync(db, 0, sizeof(float) * n, s) != cudaSuccess) return HOISDF_E_SHAPE;
  // (grid.x = row tiles: up to 2^31 - 1; grid.y = column tiles)
  const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(gy));
  if (gy > 65535) return HOISDF_E_SHAPE;
  linear_bwd_prep_kernel<<<grid, dim3(32, 8), 0, s>>>(dy, lddy, y, ldy, m, n, act, reinterpret_cast<const unsigned*>(amax),
                                                      reinterpret_cast<__half*>(dz_hi), reinterpret_cast<__half*>(dz_lo), ld_dz,
                                                      reinterpret_cast<__half*>(dzt_a), reinterpret_cast<__half*>(dzt_b),
                                                      reinterpret_cast<__half*>(dzt_c), ld_dzt, db, scale_out);
  return launch_status();
}

HOISDF_API int hoisdf_split_rows_t(const float* x, int64_t m, int64_t k, int64_t ldx, uint16_t* hi, uint16_t* lo, int64_t ldh,
                                   void* stream) {
  if (x == nullptr || hi == nullptr || lo == nullptr) return HOISDF_E_NULL;
  if (m <= 0 || k <= 0 || ldx < k || ldh < m) return HOISDF_E_SHAPE;
  const int64_t gy = ceil_div(m, 32), gx = ceil_div(k, 32);
  if (gy > 65535 || gx > 0x7fffffffLL) return HOISDF_E_SHAPE;
  split_rows_t_kernel<<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(gy)), dim3(32, 8), 0,
                        static_cast<cudaStream_t>(stream)>>>(x, m, k, ldx, reinterpret_cast<__half*>(hi),
                                                             reinterpret_cast<__half*>(lo), ldh);
  return launch_status();
}

HOISDF_API int64_t hoisdf_linear_train_workspace_bytes(int64_t m, int64_t n, int64_t k) {
  if (m <= 0 || n <= 0 || k <= 0) return 0;
  return train_ws(m, n, k).total;
}

HOISDF_API int hoisdf_linear_train_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, int64_t m,
                                       int64_t n, int64_t k, int32_t act, float* y, int64_t ldy, void* workspace,
                                       int64_t workspace_bytes, void* stream) {
  if (x == nullptr || w == nullptr || y == nullptr || workspace == nullptr) return HOISDF_E_NULL;
  if (m <= 0 || n <= 0 || k <= 0 || ldx < k || ldw < k || ldy < n) return HOISDF_E_SHAPE;
  const TrainWs o = train_ws(m, n, k);
  if (workspace_bytes < o.total) return HOISDF_E_WORKSPACE;
  if (!aligned16(workspace)) return HOISDF_E_ALIGN;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const int64_t ldh = up8(k);
  uint16_t* xh = reinterpret_cast<uint16_t*>(ws + o.xs);
  uint16_t* xl = xh + m * ldh;
  uint16_t* wa = reinterpret_cast<uint16_t*>(ws + o.wp);
  int st = hoisdf_split_rows(x, m, k, ldx, (k + 3) / 4 * 4, xh, xl, ldh, stream);
  if (st != HOISDF_OK) return st;
  st = hoisdf_pack_h3(w, n, k, ldw, wa, wa + n * ldh, wa + 2 * n * ldh, ldh, stream);
  if (st != HOISDF_OK) return st;
  hoisdf_linear_h3_args a;
  memset(&a, 0, sizeof(a));
  a.x_hi = xh; a.x_lo = xl; a.ldx = ldh;
  a.w_a = wa; a.w_b = wa + n * ldh; a.w_c = wa + 2 * n * ldh; a.ldw = ldh; a.bias = bias;
  a.y = y; a.ldy = ldy; a.m = m; a.n = n; a.k = k; a.act = act; a.chunk_kb = 4;
  return hoisdf_linear_h3_fwd(&a, stream);
}

HOISDF_API int hoisdf_linear_train_bwd(const float* dy, int64_t lddy, const float* y, int64_t ldy, const float* x, int64_t ldx,
                                       const float* w, int64_t ldw, int64_t m, int64_t n, int64_t k, int32_t act, float* dx,
                                       int64_t lddx, float* dwt, int64_t lddwt, float* db, void* workspace,
                                       int64_t workspace_bytes, void* stream) {
  if (dy == nullptr || workspace == nullptr || (dx != nullptr && w == nullptr) || (dwt != nullptr && x == nullptr) ||
      (act == HOISDF_ACT_RELU && y == nullptr))
    return HOISDF_E_NULL;
  if (m <= 0 || n <= 0 || k <= 0 || lddy < n || (dx != nullptr && (lddx < k || ldw < k)) ||
      (dwt != nullptr && (lddwt < n || ldx < k)))
    return HOISDF_E_SHAPE;
  const TrainWs o = train_ws(m, n, k);
  if (workspace_bytes < o.total) return HOISDF_E_WORKSPACE;
  if (!aligned16(workspace)) return HOISDF_E_ALIGN;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* amax = reinterpret_cast<float*>(ws + o.scal);
  float* scale = amax + 1;
  const int64_t ld_dz = up8(n), ld_t = up8(m);
  uint16_t* dzh = reinterpret_cast<uint16_t*>(ws + o.dz);
  uint16_t* dzl = dzh + m * ld_dz;
  uint16_t* ta = reinterpret_cast<uint16_t*>(ws + o.dzt);
  int st = hoisdf_absmax(dy, m, n, lddy, amax, stream);
  if (st != HOISDF_OK) return st;
  st = hoisdf_linear_bwd_prep(dy, lddy, y, ldy, m, n, act, amax, dzh, dzl, ld_dz, ta, ta + n * ld_t, ta + 2 * n * ld_t, ld_t, db,
                              scale, stream);
  if (st != HOISDF_OK) return st;
  hoisdf_linear_h3_args a;
  if (dx != nullptr) {                                     // dX = (dZ / s . W) * s: weight operand = W^T (k rows of n)
    uint16_t* wt = reinterpret_cast<uint16_t*>(ws + o.wt);
    const int64_t gy = ceil_div(n, 32), gx = ceil_div(k, 32);
    if (gy > 65535) return HOISDF_E_SHAPE;
    pack_h3_t_kernel<<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(gy)), dim3(32, 8), 0,
                       static_cast<cudaStream_t>(stream)>>>(w, n, k, ldw, reinterpret_cast<__half*>(wt),
                                                            reinterpret_cast<__half*>(wt + k * ld_dz),
                                                            reinterpret_cast<__half*>(wt + 2 * k * ld_dz), ld_dz);
    st = launch_status();
    if (st != HOISDF_OK) return st;
    memset(&a, 0, sizeof(a));
    a.x_hi = dzh; a.x_lo = dzl; a.ldx = ld_dz;
    a.w_a = wt; a.w_b = wt + k * ld_dz; a.w_c = wt + 2 * k * ld_dz; a.ldw = ld_dz;
    a.y = dx; a.ldy = lddx; a.m = m; a.n = k; a.k = n; a.act = HOISDF_ACT_NONE; a.chunk_kb = 4; a.y_scale = scale;
    st = hoisdf_linear_h3_fwd(&a, stream);
    if (st != HOISDF_OK) return st;
  }
  if (dwt != nullptr) {                                    // dW^T = (X^T . dZ / s) * s, split over the long contraction
    uint16_t* xth = reinterpret_cast<uint16_t*>(ws + o.xt);
    uint16_t* xtl = xth + k * ld_t;
    st = hoisdf_split_rows_t(x, m, k, ldx, xth, xtl, ld_t, stream);
    if (st != HOISDF_OK) return st;
    memset(&a, 0, sizeof(a));
    a.x_hi = xth; a.x_lo = xtl; a.ldx = ld_t;
    a.w_a = ta; a.w_b = ta + n * ld_t; a.w_c = ta + 2 * n * ld_t; a.ldw = ld_t;
    a.y = dwt; a.ldy = lddwt; a.m = k; a.n = n; a.k = m; a.act = HOISDF_ACT_NONE; a.chunk_kb = 4; a.y_scale = scale;
    a.split_k = 1;
    st = hoisdf_linear_h3_fwd(&a, stream);
    if (st != HOISDF_OK) return st;
  }
  return HOISDF_OK;
}
