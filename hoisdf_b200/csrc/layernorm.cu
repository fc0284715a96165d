// y = LayerNorm(x + res) (eps 1e-5, biased variance) with an optional second LayerNorm of the result -- the
// post-norm residual blocks of upstream common/nets/transformer.py:296-301,384-394 and the shared
// `inter_norm` / decoder `norm` applied to every layer output (transformer.py:196-197, 243-244).
// HBM-bound: one warp per row, 128-bit loads, the row stays in registers between the two passes.
#include "common.cuh"

namespace hoisdf {

template <int D>  // row width, multiple of 128
__global__ void __launch_bounds__(256) add_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ y,
                                                            const float* __restrict__ gamma2,
                                                            const float* __restrict__ beta2, float* __restrict__ y2,
                                                            int64_t rows) {
  constexpr int Q = D / 128;
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float v[Q * 4];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int c = q * 128 + lane * 4;
    float4 a = __ldg(reinterpret_cast<const float4*>(x + r * D + c));
    if (res != nullptr) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(res + r * D + c));
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    v[q * 4 + 0] = a.x; v[q * 4 + 1] = a.y; v[q * 4 + 2] = a.z; v[q * 4 + 3] = a.w;
  }
  auto normalise = [&](const float* g, const float* bt, float* dst) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < Q * 4; ++i) s += v[i];
    const float mean = warp_sum(s) * (1.0f / D);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < Q * 4; ++i) {
      const float d = v[i] - mean;
      ss = fmaf(d, d, ss);
    }
    const float rstd = rsqrtf(warp_sum(ss) * (1.0f / D) + 1e-5f);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int c = q * 128 + lane * 4;
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bt + c));
      v[q * 4 + 0] = (v[q * 4 + 0] - mean) * rstd * gg.x + bb.x;
      v[q * 4 + 1] = (v[q * 4 + 1] - mean) * rstd * gg.y + bb.y;
      v[q * 4 + 2] = (v[q * 4 + 2] - mean) * rstd * gg.z + bb.z;
      v[q * 4 + 3] = (v[q * 4 + 3] - mean) * rstd * gg.w + bb.w;
      *reinterpret_cast<float4*>(dst + r * D + c) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
    }
  };
  normalise(gamma, beta, y);
  if (y2 != nullptr) normalise(gamma2, beta2, y2);
}

}  // namespace hoisdf

using namespace hoisdf;

HOISDF_API int hoisdf_add_layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta,
                                        float* y, const float* gamma2, const float* beta2, float* y2, int64_t rows,
                                        int64_t d, void* stream) {
  if (x == nullptr || gamma == nullptr || beta == nullptr || y == nullptr) return HOISDF_E_NULL;
  if (y2 != nullptr && (gamma2 == nullptr || beta2 == nullptr)) return HOISDF_E_NULL;
  if (rows == 0) return HOISDF_OK;
  if (rows < 0 || d != 256) return d == 256 ? HOISDF_E_SHAPE : HOISDF_E_UNSUPPORTED;
  if (!aligned16(x) || !aligned16(y) || !aligned16(gamma) || !aligned16(beta) || (res && !aligned16(res)) ||
      (y2 && (!aligned16(y2) || !aligned16(gamma2) || !aligned16(beta2))))
    return HOISDF_E_ALIGN;
  add_layernorm_kernel<256><<<static_cast<unsigned>(ceil_div(rows, 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, res, gamma, beta, y, gamma2, beta2, y2, rows);
  return launch_status();
}
