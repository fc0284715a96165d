// SDF decoder (upstream common/nets/sdf_net.py:87-122), its input tail (NeRF embedding,
// common/utils/sdf_utils.py:96-141) and token assembly (main/model.py:123-126, 520-562).
//
// Row-buffer layout used by the whole candidate / query path (one row per 3-D point, ld >= 516 floats):
//   [0,256)   relu(linear_sdfin)            written by the sdfin layer-1 GEMM epilogue
//   [256,286) NeRF posenc  [286,289) xyz    written by posenc_kernel
//   [289,292) 0 (K padding of linh0)
//   [292,515) relu(linh1)  (223)            written by the linh1 GEMM epilogue  } linh2 reads cols [0,516)
//   [515]     0                                                                  } with a permuted weight
// so the upstream `torch.cat([xh, input], 1)` skip connection (sdf_net.py:97-98) costs no copy.
#include "tc_common.cuh"

namespace hoisdf {

constexpr int kFea = 256;       // linear_sdfin output
constexpr int kDecIn = 289;     // 256 + 30 + 3
constexpr int kDecInPad = 292;
constexpr int kSkipOff = 292;
constexpr int kH1 = 223;
constexpr int kRowLd = 516;

// 37 threads of work per row: j<30 posenc, 30..32 xyz, 33..35 zero pad, 36 -> column 515 zero
__global__ void posenc_kernel(const int32_t* __restrict__ lattice_index, const float* __restrict__ points,
                              int64_t rows, int bins, float* __restrict__ out, int64_t ld, int64_t col0) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t r = i / 37;
  const int j = static_cast<int>(i - r * 37);
  if (r >= rows) return;
  float x[3];
  if (lattice_index != nullptr) {
    lattice_point(lattice_index[r], bins, x[0], x[1], x[2]);
  } else {
    x[0] = points[r * 3 + 0]; x[1] = points[r * 3 + 1]; x[2] = points[r * 3 + 2];
  }
  float* o = out + r * ld + col0;
  if (j < 30) {
    const int oct = j / 6, w = j % 6;
    const float a = x[w % 3] * static_cast<float>(1 << oct);  // exact scaling by 2^k
    o[j] = (w < 3) ? sinf(a) : cosf(a);
  } else if (j < 33) {
    o[j] = x[j - 30];
  } else if (j < 36) {
    o[j] = 0.f;
  } else {
    out[r * ld + (kRowLd - 1)] = 0.f;
  }
}

// Split-half row buffer of the FP16x3 path (pitch >= 520 halfs per plane; every window starts on a 16-byte boundary):
//   [0,256) relu(linear_sdfin) | [256,286) posenc | [286,289) xyz | [289,296) 0 | [296,519) relu(linh1) | [519] 0
constexpr int kSkipOffH = 296;
constexpr int kRowLdH = 520;

// 41 threads of work per row: j<30 posenc, 30..32 xyz, 33..39 zero pad, 40 -> column 519 zero
__global__ void posenc_split_kernel(const int32_t* __restrict__ lattice_index, const float* __restrict__ points,
                                    int64_t rows, int bins, __half* __restrict__ hi, __half* __restrict__ lo,
                                    int64_t ld) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t r = i / 41;
  const int j = static_cast<int>(i - r * 41);
  if (r >= rows) return;
  float x[3];
  if (lattice_index != nullptr) {
    lattice_point(lattice_index[r], bins, x[0], x[1], x[2]);
  } else {
    x[0] = points[r * 3 + 0]; x[1] = points[r * 3 + 1]; x[2] = points[r * 3 + 2];
  }
  float val = 0.f;
  int64_t col = kFea + j;
  if (j < 30) {
    const int oct = j / 6, w = j % 6;
    const float a = x[w % 3] * static_cast<float>(1 << oct);  // exact scaling by 2^k
    val = (w < 3) ? sinf(a) : cosf(a);
  } else if (j < 33) {
    val = x[j - 30];
  } else if (j == 40) {
    col = kRowLdH - 1;
  }
  __half h, l;
  tc::split_half(val, h, l);
  hi[r * ld + col] = h;
  lo[r * ld + col] = l;
}

// out[r] = tanh(h[r,:512] . w4 + b4) with h in split-half format (lo == nullptr: hi plane only); one warp per row
__global__ void __launch_bounds__(256) sdf_head_split_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo,
                                                             int64_t ldh, int64_t rows, const float* __restrict__ w4,
                                                             const float* __restrict__ b4, float* __restrict__ out,
                                                             float clamp) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int c = q * 256 + lane * 8;
    const uint4 ph = __ldg(reinterpret_cast<const uint4*>(hi + r * ldh + c));
    const uint4 pl = lo != nullptr ? __ldg(reinterpret_cast<const uint4*>(lo + r * ldh + c)) : make_uint4(0, 0, 0, 0);
    const float4 wa = __ldg(reinterpret_cast<const float4*>(w4 + c));
    const float4 wb = __ldg(reinterpret_cast<const float4*>(w4 + c + 4));
    const uint32_t hw[4] = {ph.x, ph.y, ph.z, ph.w}, lw[4] = {pl.x, pl.y, pl.z, pl.w};
    const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a0 = tc::join_half(__ushort_as_half(static_cast<unsigned short>(hw[j] & 0xffffu)),
                                     __ushort_as_half(static_cast<unsigned short>(lw[j] & 0xffffu)));
      const float a1 = tc::join_half(__ushort_as_half(static_cast<unsigned short>(hw[j] >> 16)),
                                     __ushort_as_half(static_cast<unsigned short>(lw[j] >> 16)));
      s = fmaf(a0, ww[2 * j], s);
      s = fmaf(a1, ww[2 * j + 1], s);
    }
  }
  s = warp_sum(s);
  if (lane == 0) {
    float t = tanhf(s + __ldg(b4));
    if (clamp > 0.f) t = fminf(fmaxf(t, -clamp), clamp);
    out[r] = t;
  }
}

// out[r] = tanh(h[r,:512] . w4 + b4); one warp per row
__global__ void __launch_bounds__(256) sdf_head_kernel(const float* __restrict__ h, int64_t ldh, int64_t rows,
                                                       const float* __restrict__ w4, const float* __restrict__ b4,
                                                       float* __restrict__ out, float clamp) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* hr = h + r * ldh;
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(hr + q * 128 + lane * 4));
    const float4 w = __ldg(reinterpret_cast<const float4*>(w4 + q * 128 + lane * 4));
    s = fmaf(a.x, w.x, s); s = fmaf(a.y, w.y, s); s = fmaf(a.z, w.z, s); s = fmaf(a.w, w.w, s);
  }
  s = warp_sum(s);
  if (lane == 0) {
    float t = tanhf(s + __ldg(b4));
    if (clamp > 0.f) t = fminf(fmaxf(t, -clamp), clamp);
    out[r] = t;
  }
}

__global__ void sdf_pad_input_kernel(const float* __restrict__ in, int64_t rows, float* __restrict__ x, int64_t ld) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t r = i / 293;
  const int c = static_cast<int>(i - r * 293);
  if (r >= rows) return;
  if (c < kDecIn) x[r * ld + c] = in[r * kDecIn + c];
  else if (c < kDecInPad) x[r * ld + c] = 0.f;
  else x[r * ld + (kRowLd - 1)] = 0.f;
}

// tokens[b, t0 + t, c] = c<3 ? xyz : c<33 ? posenc : fea * sigmoid(sdf/beta)/beta   (upstream model.py:123-126,520-531)
__global__ void tokens_kernel(const float* __restrict__ xyz, const float* __restrict__ posenc,
                              const float* __restrict__ fea, int64_t ld_fea, const float* __restrict__ sdf,
                              const float* __restrict__ beta, int64_t batch, int64_t p, float* __restrict__ tokens,
                              int64_t s_total, int64_t t0) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= batch * p * 256) return;
  const int c = static_cast<int>(i & 255);
  const int64_t bt = i >> 8;
  const int64_t b = bt / p, t = bt - b * p;
  float v;
  if (c < 3) {
    v = xyz[bt * 3 + c];
  } else if (c < 33) {
    v = posenc[bt * 30 + (c - 3)];
  } else {
    const float be = __ldg(beta);
    const float z = __fdiv_rn(sdf[bt], be);
    const float sig = __fdiv_rn(__fdiv_rn(1.f, 1.f + expf(-z)), be);
    v = fea[bt * ld_fea + (c - 33)] * sig;
  }
  tokens[(b * s_total + t0 + t) * 256 + c] = v;
}

}  // namespace hoisdf

using namespace hoisdf;

HOISDF_API int hoisdf_posenc_fwd(const int32_t* lattice_index, const float* points, int64_t rows, int32_t bins,
                                 float* out, int64_t ld_out, int64_t col0, void* stream) {
  if (out == nullptr || (lattice_index == nullptr && points == nullptr)) return HOISDF_E_NULL;
  if (rows == 0) return HOISDF_OK;
  if (rows < 0 || ld_out < kRowLd || col0 != kFea) return HOISDF_E_SHAPE;
  const int64_t n = rows * 37;
  HOISDF_LAUNCH(posenc_kernel, static_cast<unsigned>(ceil_div(n, 256)), 256, static_cast<cudaStream_t>(stream),
                lattice_index, points, rows, bins, out, ld_out, col0);
  return launch_status();
}

HOISDF_API int hoisdf_sdf_pad_input(const float* in, int64_t rows, float* x, int64_t ldx, void* stream) {
  if (in == nullptr || x == nullptr) return HOISDF_E_NULL;
  if (rows == 0) return HOISDF_OK;
  if (rows < 0 || ldx < kRowLd) return HOISDF_E_SHAPE;
  const int64_t n = rows * 293;
  HOISDF_LAUNCH(sdf_pad_input_kernel, static_cast<unsigned>(ceil_div(n, 256)), 256, static_cast<cudaStream_t>(stream), in,
                rows, x, ldx);
  return launch_status();
}

HOISDF_API int hoisdf_sdf_decoder_fwd(const hoisdf_sdf_weights* w, float* x, int64_t ldx, int64_t rows, float* h_a,
                                      float* h_b, float* out_sdf, float clamp, void* stream) {
  if (w == nullptr || x == nullptr || h_a == nullptr || h_b == nullptr || out_sdf == nullptr) return HOISDF_E_NULL;
  if (w->w0 == nullptr || w->w1 == nullptr || w->w2 == nullptr || w->w3 == nullptr || w->w4 == nullptr ||
      w->b0 == nullptr || w->b1 == nullptr || w->b2 == nullptr || w->b3 == nullptr || w->b4 == nullptr)
    return HOISDF_E_NULL;
  if (rows == 0) return HOISDF_OK;
  if (rows < 0 || ldx < kRowLd) return HOISDF_E_SHAPE;
  if ((ldx & 3) || !aligned16(x) || !aligned16(h_a) || !aligned16(h_b) || !aligned16(w->w4)) return HOISDF_E_ALIGN;
  hoisdf_linear_args a;
  int st;
  const bool tc = w->w0_lo != nullptr && w->w1_lo != nullptr && w->w2_lo != nullptr && w->w3_lo != nullptr;
  // linh0: x[:, 0:292] -> h_a (512), ReLU
  a = {x, ldx, 0, 0, w->w0, kDecInPad, w->b0, nullptr, h_a, 512, 0, 0, rows, 512, kDecInPad, HOISDF_ACT_RELU, tc ? w->w0_lo : nullptr, w->tf32_passes};
  if ((st = hoisdf_linear_fwd(&a, stream)) != HOISDF_OK) return st;
  // linh1: h_a -> x[:, 292:515] (223), ReLU
  a = {h_a, 512, 0, 0, w->w1, 512, w->b1, nullptr, x + kSkipOff, ldx, 0, 0, rows, kH1, 512, HOISDF_ACT_RELU, tc ? w->w1_lo : nullptr, w->tf32_passes};
  if ((st = hoisdf_linear_fwd(&a, stream)) != HOISDF_OK) return st;
  // linh2: x[:, 0:516] (input | pad | h1 | pad, weight columns permuted to match) -> h_a, ReLU
  a = {x, ldx, 0, 0, w->w2, kRowLd, w->b2, nullptr, h_a, 512, 0, 0, rows, 512, kRowLd, HOISDF_ACT_RELU, tc ? w->w2_lo : nullptr, w->tf32_passes};
  if ((st = hoisdf_linear_fwd(&a, stream)) != HOISDF_OK) return st;
  // linh3: h_a -> h_b, ReLU
  a = {h_a, 512, 0, 0, w->w3, 512, w->b3, nullptr, h_b, 512, 0, 0, rows, 512, 512, HOISDF_ACT_RELU, tc ? w->w3_lo : nullptr, w->tf32_passes};
  if ((st = hoisdf_linear_fwd(&a, stream)) != HOISDF_OK) return st;
  // linh4 + tanh
  HOISDF_LAUNCH(sdf_head_kernel, static_cast<unsigned>(ceil_div(rows, 8)), 256, static_cast<cudaStream_t>(stream), h_b,
                int64_t(512), rows, w->w4, w->b4, out_sdf, clamp);
  return launch_status();
}

HOISDF_API int hoisdf_tokens_fwd(const float* xyz, const float* posenc, const float* fea, int64_t ld_fea,
                                 const float* sdf, const float* beta, int64_t batch, int64_t p, float* tokens,
                                 int64_t s_total, int64_t t0, void* stream) {
  if (xyz == nullptr || posenc == nullptr || fea == nullptr || sdf == nullptr || beta == nullptr ||
      tokens == nullptr)
    return HOISDF_E_NULL;
  if (batch <= 0 || p <= 0 || t0 < 0 || t0 + p > s_total || ld_fea < 223) return HOISDF_E_SHAPE;
  const int64_t n = batch * p * 256;
  HOISDF_LAUNCH(tokens_kernel, static_cast<unsigned>(ceil_div(n, 256)), 256, static_cast<cudaStream_t>(stream), xyz, posenc,
                fea, ld_fea, sdf, beta, batch, p, tokens, s_total, t0);
  return launch_status();
}

HOISDF_API int hoisdf_posenc_split_fwd(const int32_t* lattice_index, const float* points, int64_t rows, int32_t bins,
                                       uint16_t* out_hi, uint16_t* out_lo, int64_t ld_out, void* stream) {
  if (out_hi == nullptr || out_lo == nullptr || (lattice_index == nullptr && points == nullptr)) return HOISDF_E_NULL;
  if (rows == 0) return HOISDF_OK;
  if (rows < 0 || ld_out < kRowLdH) return HOISDF_E_SHAPE;
  const int64_t n = rows * 41;
  HOISDF_LAUNCH(posenc_split_kernel, static_cast<unsigned>(ceil_div(n, 256)), 256, static_cast<cudaStream_t>(stream),
                lattice_index, points, rows, bins, reinterpret_cast<__half*>(out_hi), reinterpret_cast<__half*>(out_lo),
                ld_out);
  return launch_status();
}

HOISDF_API int hoisdf_sdf_decoder_h3_fwd(const hoisdf_sdf_weights_h3* w, uint16_t* x_hi, uint16_t* x_lo, int64_t ldx,
                                         int64_t rows, uint16_t* ha_hi, uint16_t* ha_lo, uint16_t* hb_hi,
                                         uint16_t* hb_lo, int64_t ldh, float* out_sdf, float clamp, void* stream) {
  if (w == nullptr || x_hi == nullptr || x_lo == nullptr || ha_hi == nullptr || ha_lo == nullptr || hb_hi == nullptr ||
      hb_lo == nullptr || out_sdf == nullptr || w->w4 == nullptr || w->b4 == nullptr)
    return HOISDF_E_NULL;
  for (int l = 0; l < 4; ++l)
    if (w->w[l][0] == nullptr || w->w[l][1] == nullptr || w->w[l][2] == nullptr || w->b[l] == nullptr)
      return HOISDF_E_NULL;
  if (rows == 0) return HOISDF_OK;
  if (rows < 0 || ldx < kRowLdH || ldh < 512) return HOISDF_E_SHAPE;
  if ((ldx & 7) || (ldh & 7) || !aligned16(x_hi) || !aligned16(x_lo) || !aligned16(ha_hi) || !aligned16(ha_lo) ||
      !aligned16(hb_hi) || !aligned16(hb_lo) || !aligned16(w->w4))
    return HOISDF_E_ALIGN;
  hoisdf_linear_h3_args a;
  int st;
  auto layer = [&](int l, const uint16_t* xh, const uint16_t* xl, int64_t ld_in, int64_t k, uint16_t* yh, uint16_t* yl,
                   int64_t ld_out, int64_t n) {
    a = {xh, xl, ld_in, 0, 0, w->w[l][0], w->w[l][1], w->w[l][2], w->ldw[l], w->b[l], nullptr,
         nullptr, 0, yh, yl, ld_out, rows, n, k, HOISDF_ACT_RELU, w->chunk_kb, nullptr, nullptr, 0, w->single_pass};
    return hoisdf_linear_h3_fwd(&a, stream);
  };
  // linh0: x[:, 0:289] -> h_a (512);  linh1: h_a -> x[:, 296:519] (223);  linh2: x[:, 0:519] (weight columns
  // permuted to [input | 0 | h1]) -> h_a;  linh3: h_a -> h_b;  all ReLU
  if ((st = layer(0, x_hi, x_lo, ldx, kDecIn, ha_hi, ha_lo, ldh, 512)) != HOISDF_OK) return st;
  if ((st = layer(1, ha_hi, ha_lo, ldh, 512, x_hi + kSkipOffH, x_lo + kSkipOffH, ldx, kH1)) != HOISDF_OK) return st;
  if ((st = layer(2, x_hi, x_lo, ldx, kSkipOffH + kH1, ha_hi, ha_lo, ldh, 512)) != HOISDF_OK) return st;
  if ((st = layer(3, ha_hi, ha_lo, ldh, 512, hb_hi, hb_lo, ldh, 512)) != HOISDF_OK) return st;
  HOISDF_LAUNCH(sdf_head_split_kernel, static_cast<unsigned>(ceil_div(rows, 8)), 256, static_cast<cudaStream_t>(stream),
                reinterpret_cast<const __half*>(hb_hi),
                w->single_pass ? nullptr : reinterpret_cast<const __half*>(hb_lo),      // single-product chain: only
                ldh, rows, w->w4, w->b4, out_sdf, clamp);                               // the hi planes were written
  return launch_status();
}
