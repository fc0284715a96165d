// ResNet-50 stem pieces in split-half format (upstream common/nets/resnet.py:70-76 -- torchvision's conv1 7x7 s2 p3 and
// maxpool 3x3 s2 p1).  Everything else of the backbone is the FP16x3 implicit-GEMM kernel (linear_h3.cu).
// Both kernels are HBM-bound elementwise passes: 128-bit stores, one thread per 8 output halfs.
#include "tc_common.cuh"

namespace hoisdf {
using namespace tc;

constexpr int kStemK = 147;      // 7 * 7 * 3
constexpr int kStemKPad = 160;   // 5 K blocks of 32 halfs

__device__ __forceinline__ uint32_t pack2(__half a, __half b) {
  return static_cast<uint32_t>(__half_as_ushort(a)) | (static_cast<uint32_t>(__half_as_ushort(b)) << 16);
}

// one thread = 8 consecutive im2col columns of one output pixel
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ img, int64_t batch, int h, int w,
                                                          __half* __restrict__ hi, __half* __restrict__ lo,
                                                          int64_t ldh) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int oh = h >> 1, ow = w >> 1;
  constexpr int G = kStemKPad / 8;
  if (i >= batch * oh * ow * G) return;
  const int g = static_cast<int>(i % G);
  const int64_t pix = i / G;
  const int ox = static_cast<int>(pix % ow);
  const int oy = static_cast<int>((pix / ow) % oh);
  const int64_t b = pix / (static_cast<int64_t>(ow) * oh);
  const float* base = img + b * 3 * h * w;
  __half hh[8], ll[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = g * 8 + j;
    float v = 0.f;
    if (col < kStemK) {
      const int tap = col / 3, c = col - tap * 3;
      const int ky = tap / 7, kx = tap - ky * 7;
      const int iy = 2 * oy - 3 + ky, ix = 2 * ox - 3 + kx;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w) v = __ldg(base + (static_cast<int64_t>(c) * h + iy) * w + ix);
    }
    split_half(v, hh[j], ll[j]);
  }
  *reinterpret_cast<uint4*>(hi + pix * ldh + g * 8) =
      make_uint4(pack2(hh[0], hh[1]), pack2(hh[2], hh[3]), pack2(hh[4], hh[5]), pack2(hh[6], hh[7]));
  *reinterpret_cast<uint4*>(lo + pix * ldh + g * 8) =
      make_uint4(pack2(ll[0], ll[1]), pack2(ll[2], ll[3]), pack2(ll[4], ll[5]), pack2(ll[6], ll[7]));
}

// one thread = 8 channels of one output pixel
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const __half* __restrict__ xh, const __half* __restrict__ xl,
                                                           int64_t ldx, int64_t batch, int h, int w, int c,
                                                           __half* __restrict__ yh, __half* __restrict__ yl,
                                                           int64_t ldy) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int oh = h >> 1, ow = w >> 1, G = c >> 3;
  if (i >= batch * oh * ow * G) return;
  const int g = static_cast<int>(i % G);
  const int64_t pix = i / G;
  const int ox = static_cast<int>(pix % ow);
  const int oy = static_cast<int>((pix / ow) % oh);
  const int64_t b = pix / (static_cast<int64_t>(ow) * oh);
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = 2 * oy - 1 + ky;
    if (iy < 0 || iy >= h) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = 2 * ox - 1 + kx;
      if (ix < 0 || ix >= w) continue;
      const int64_t off = ((b * h + iy) * w + ix) * ldx + g * 8;
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(xh + off));
      const uint4 d = __ldg(reinterpret_cast<const uint4*>(xl + off));
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        m[2 * j] = fmaxf(m[2 * j], join_half(__ushort_as_half(static_cast<unsigned short>(aw[j] & 0xffffu)),
                                             __ushort_as_half(static_cast<unsigned short>(dw[j] & 0xffffu))));
        m[2 * j + 1] = fmaxf(m[2 * j + 1], join_half(__ushort_as_half(static_cast<unsigned short>(aw[j] >> 16)),
                                                     __ushort_as_half(static_cast<unsigned short>(dw[j] >> 16))));
      }
    }
  }
  __half hh[8], ll[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) split_half(m[j], hh[j], ll[j]);
  *reinterpret_cast<uint4*>(yh + pix * ldy + g * 8) =
      make_uint4(pack2(hh[0], hh[1]), pack2(hh[2], hh[3]), pack2(hh[4], hh[5]), pack2(hh[6], hh[7]));
  *reinterpret_cast<uint4*>(yl + pix * ldy + g * 8) =
      make_uint4(pack2(ll[0], ll[1]), pack2(ll[2], ll[3]), pack2(ll[4], ll[5]), pack2(ll[6], ll[7]));
}

}  // namespace hoisdf

using namespace hoisdf;

HOISDF_API int hoisdf_stem_im2col_split(const float* img, int64_t batch, int64_t h, int64_t w, uint16_t* hi,
                                        uint16_t* lo, int64_t ldh, void* stream) {
  if (img == nullptr || hi == nullptr || lo == nullptr) return HOISDF_E_NULL;
  if (batch == 0) return HOISDF_OK;
  if (batch < 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1) || h > 32768 || w > 32768 || ldh < kStemKPad)
    return HOISDF_E_SHAPE;
  if ((ldh & 7) || !aligned16(hi) || !aligned16(lo)) return HOISDF_E_ALIGN;
  const int64_t total = batch * (h / 2) * (w / 2) * (kStemKPad / 8);
  stem_im2col_kernel<<<static_cast<unsigned>(ceil_div(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      img, batch, static_cast<int>(h), static_cast<int>(w), reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo),
      ldh);
  return launch_status();
}

HOISDF_API int hoisdf_maxpool3x3s2_split(const uint16_t* x_hi, const uint16_t* x_lo, int64_t ldx, int64_t batch,
                                         int64_t h, int64_t w, int64_t c, uint16_t* y_hi, uint16_t* y_lo, int64_t ldy,
                                         void* stream) {
  if (x_hi == nullptr || x_lo == nullptr || y_hi == nullptr || y_lo == nullptr) return HOISDF_E_NULL;
  if (batch == 0) return HOISDF_OK;
  if (batch < 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1) || h > 32768 || w > 32768 || c <= 0 || (c & 7) || ldx < c ||
      ldy < c)
    return HOISDF_E_SHAPE;
  if ((ldx & 7) || (ldy & 7) || !aligned16(x_hi) || !aligned16(x_lo) || !aligned16(y_hi) || !aligned16(y_lo))
    return HOISDF_E_ALIGN;
  const int64_t total = batch * (h / 2) * (w / 2) * (c / 8);
  maxpool3x3s2_kernel<<<static_cast<unsigned>(ceil_div(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(x_hi), reinterpret_cast<const __half*>(x_lo), ldx, batch, static_cast<int>(h),
      static_cast<int>(w), static_cast<int>(c), reinterpret_cast<__half*>(y_hi), reinterpret_cast<__half*>(y_lo), ldy);
  return launch_status();
}
