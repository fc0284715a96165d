// y = LayerNorm(x + res) (eps 1e-5, biased variance) with an optional second LayerNorm of the result -- the
// post-norm residual blocks of upstream common/nets/transformer.py:296-301,384-394 and the shared
// `inter_norm` / decoder `norm` applied to every layer output (transformer.py:196-197, 243-244).
// HBM-bound: one warp per row, 128-bit loads, the row stays in registers between the two passes.
#include "tc_common.cuh"

namespace hoisdf {

template <int D>  // row width, multiple of 128
__global__ void __launch_bounds__(256) add_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ y,
                                                            const float* __restrict__ gamma2,
                                                            const float* __restrict__ beta2, float* __restrict__ y2,
                                                            int64_t rows, __half* __restrict__ yh_hi,
                                                            __half* __restrict__ yh_lo, int64_t ldyh,
                                                            __half* __restrict__ y2h_hi, __half* __restrict__ y2h_lo,
                                                            int64_t ldy2h) {
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
  auto normalise = [&](const float* g, const float* bt, float* dst, __half* dhi, __half* dlo, int64_t ldh) {
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
      if (dhi != nullptr) {      // the same values in split-half format for the next FP16x3 Linear
        __half h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) tc::split_half(v[q * 4 + j], h[j], l[j]);
        uint2 ph, pl;
        ph.x = static_cast<uint32_t>(__half_as_ushort(h[0])) | (static_cast<uint32_t>(__half_as_ushort(h[1])) << 16);
        ph.y = static_cast<uint32_t>(__half_as_ushort(h[2])) | (static_cast<uint32_t>(__half_as_ushort(h[3])) << 16);
        pl.x = static_cast<uint32_t>(__half_as_ushort(l[0])) | (static_cast<uint32_t>(__half_as_ushort(l[1])) << 16);
        pl.y = static_cast<uint32_t>(__half_as_ushort(l[2])) | (static_cast<uint32_t>(__half_as_ushort(l[3])) << 16);
        *reinterpret_cast<uint2*>(dhi + r * ldh + c) = ph;
        *reinterpret_cast<uint2*>(dlo + r * ldh + c) = pl;
      }
    }
  };
  normalise(gamma, beta, y, yh_hi, yh_lo, ldyh);
  if (y2 != nullptr) normalise(gamma2, beta2, y2, y2h_hi, y2h_lo, ldy2h);
}

}  // namespace hoisdf

using namespace hoisdf;

static int add_layernorm_launch(const float* x, const float* res, const float* gamma, const float* beta, float* y,
                                const float* gamma2, const float* beta2, float* y2, int64_t rows, int64_t d,
                                uint16_t* yh_hi, uint16_t* yh_lo, int64_t ldyh, uint16_t* y2h_hi, uint16_t* y2h_lo,
                                int64_t ldy2h, void* stream) {
  if (x == nullptr || gamma == nullptr || beta == nullptr || y == nullptr) return HOISDF_E_NULL;
  if (y2 != nullptr && (gamma2 == nullptr || beta2 == nullptr)) return HOISDF_E_NULL;
  if (rows == 0) return HOISDF_OK;
  if (rows < 0 || d != 256) return d == 256 ? HOISDF_E_SHAPE : HOISDF_E_UNSUPPORTED;
  if (!aligned16(x) || !aligned16(y) || !aligned16(gamma) || !aligned16(beta) || (res && !aligned16(res)) ||
      (y2 && (!aligned16(y2) || !aligned16(gamma2) || !aligned16(beta2))))
    return HOISDF_E_ALIGN;
  HOISDF_LAUNCH(add_layernorm_kernel<256>, static_cast<unsigned>(ceil_div(rows, 8)), 256,
                static_cast<cudaStream_t>(stream), x, res, gamma, beta, y, gamma2, beta2, y2, rows,
                reinterpret_cast<__half*>(yh_hi), reinterpret_cast<__half*>(yh_lo), ldyh, reinterpret_cast<__half*>(y2h_hi),
                reinterpret_cast<__half*>(y2h_lo), ldy2h);
  return launch_status();
}

HOISDF_API int hoisdf_add_layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta,
                                        float* y, const float* gamma2, const float* beta2, float* y2, int64_t rows,
                                        int64_t d, void* stream) {
  return add_layernorm_launch(x, res, gamma, beta, y, gamma2, beta2, y2, rows, d, nullptr, nullptr, 0, nullptr, nullptr,
                              0, stream);
}

HOISDF_API int hoisdf_add_layernorm_split_fwd(const float* x, const float* res, const float* gamma, const float* beta,
                                              float* y, const float* gamma2, const float* beta2, float* y2,
                                              int64_t rows, int64_t d, uint16_t* yh_hi, uint16_t* yh_lo, int64_t ldyh,
                                              uint16_t* y2h_hi, uint16_t* y2h_lo, int64_t ldy2h, void* stream) {
  if ((yh_hi == nullptr) != (yh_lo == nullptr) || (y2h_hi == nullptr) != (y2h_lo == nullptr)) return HOISDF_E_NULL;
  if (y2h_hi != nullptr && y2 == nullptr) return HOISDF_E_NULL;
  if ((yh_hi != nullptr && ((ldyh & 3) || ldyh < d || (reinterpret_cast<uintptr_t>(yh_hi) & 7) ||
                            (reinterpret_cast<uintptr_t>(yh_lo) & 7))) ||
      (y2h_hi != nullptr && ((ldy2h & 3) || ldy2h < d || (reinterpret_cast<uintptr_t>(y2h_hi) & 7) ||
                             (reinterpret_cast<uintptr_t>(y2h_lo) & 7))))
    return HOISDF_E_ALIGN;
  return add_layernorm_launch(x, res, gamma, beta, y, gamma2, beta2, y2, rows, d, yh_hi, yh_lo, ldyh, y2h_hi, y2h_lo,
                              ldy2h, stream);
}
