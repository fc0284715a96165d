// Narrow Linear: Y (M, N <= 24) = act(X . W^T + b) for split-half X -- the last layer of the small heads (upstream
// main/model.py:81-90 linear_handcls / linear_obj_rot / linear_obj_rel_trans / linear_pose / linear_shape, and the 1x1
// convOut_* heads of common/nets/module.py:147-218).  A 128 x 256 tensor-core tile wastes > 90 % of its columns on
// these and is bound by per-tile latency; this kernel is a plain HBM-bound pass: one warp per row, each lane joins two
// (hi, lo) pairs per 64 columns, fp32 FMAs against L1-resident weights, butterfly reduction, lane n stores column n.
#include "tc_common.cuh"

namespace hoisdf {
using namespace tc;

enum { NARROW_ACT_NONE = 0, NARROW_ACT_RELU = 1, NARROW_ACT_SIGMOID = 2 };

template <int NMAX>
__global__ void __launch_bounds__(256) linear_narrow_kernel(const __half* __restrict__ xh, const __half* __restrict__ xl,
                                                            int64_t ldx, int64_t m, const float* __restrict__ w,
                                                            int64_t ldw, const float* __restrict__ bias, int n, int k,
                                                            int act, float* __restrict__ y, int64_t ldy) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t r = warp0; r < m; r += nwarps) {
    float acc[NMAX];
#pragma unroll
    for (int j = 0; j < NMAX; ++j) acc[j] = 0.f;
    for (int c = 2 * lane; c < k; c += 64) {        // k is even: a (c, c + 1) pair never straddles the end
      const uint32_t h2 = __ldg(reinterpret_cast<const unsigned int*>(xh + r * ldx + c));
      const uint32_t l2 = __ldg(reinterpret_cast<const unsigned int*>(xl + r * ldx + c));
      const float x0 = join_half(__ushort_as_half(static_cast<unsigned short>(h2 & 0xffffu)),
                                 __ushort_as_half(static_cast<unsigned short>(l2 & 0xffffu)));
      const float x1 = join_half(__ushort_as_half(static_cast<unsigned short>(h2 >> 16)),
                                 __ushort_as_half(static_cast<unsigned short>(l2 >> 16)));
#pragma unroll
      for (int j = 0; j < NMAX; ++j) {
        if (j < n) {
          const float2 ww = __ldg(reinterpret_cast<const float2*>(w + j * ldw + c));
          acc[j] = fmaf(x1, ww.y, fmaf(x0, ww.x, acc[j]));
        }
      }
    }
    float mine = 0.f;
#pragma unroll
    for (int j = 0; j < NMAX; ++j) {
      if (j < n) {
        float s = acc[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == j) mine = s;
      }
    }
    if (lane < n) {
      float v = mine + (bias != nullptr ? __ldg(bias + lane) : 0.f);
      if (act == NARROW_ACT_RELU) v = fmaxf(v, 0.f);
      else if (act == NARROW_ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
      y[r * ldy + lane] = v;
    }
  }
}

}  // namespace hoisdf

using namespace hoisdf;

HOISDF_API int hoisdf_linear_narrow_split_fwd(const uint16_t* x_hi, const uint16_t* x_lo, int64_t ldx, int64_t m,
                                              const float* w, int64_t ldw, const float* bias, int64_t n, int64_t k,
                                              int32_t act, float* y, int64_t ldy, void* stream) {
  if (x_hi == nullptr || x_lo == nullptr || w == nullptr || y == nullptr) return HOISDF_E_NULL;
  if (m == 0) return HOISDF_OK;
  if (m < 0 || n <= 0 || k <= 0 || ldx < k || ldw < k || ldy < n) return HOISDF_E_SHAPE;
  if (n > 24 || (k & 1) || act < 0 || act > 2) return HOISDF_E_UNSUPPORTED;
  if ((ldx & 1) || (ldw & 1) || (reinterpret_cast<uintptr_t>(x_hi) & 3) || (reinterpret_cast<uintptr_t>(x_lo) & 3) ||
      (reinterpret_cast<uintptr_t>(w) & 7))
    return HOISDF_E_ALIGN;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t blocks64 = ceil_div(m, 8);
  const unsigned blocks = static_cast<unsigned>(blocks64 < 148 * 16 ? blocks64 : 148 * 16);
  const __half* xh = reinterpret_cast<const __half*>(x_hi);
  const __half* xl = reinterpret_cast<const __half*>(x_lo);
  const int ni = static_cast<int>(n), ki = static_cast<int>(k);
  if (n <= 4) HOISDF_LAUNCH(linear_narrow_kernel<4>, blocks, 256, s, xh, xl, ldx, m, w, ldw, bias, ni, ki, act, y, ldy);
  else if (n <= 12) HOISDF_LAUNCH(linear_narrow_kernel<12>, blocks, 256, s, xh, xl, ldx, m, w, ldw, bias, ni, ki, act, y, ldy);
  else HOISDF_LAUNCH(linear_narrow_kernel<24>, blocks, 256, s, xh, xl, ldx, m, w, ldw, bias, ni, ki, act, y, ldy);
  return launch_status();
}
