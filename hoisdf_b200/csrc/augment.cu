// Photometric augmentation of the training feed (upstream data/ho3d.py:355-364, data/dexycb.py:310-321):
//   img = img.filter(ImageFilter.GaussianBlur(random.random() * blur_radius))
//   img = dataset_util.color_jitter(img, brightness, saturation, hue, contrast)      (data/dataset_util.py:144-201)
// on the batch of warped 8-bit frames (hoisdf_image_crop_fwd's out_u8) in device memory, BIT-EXACT with the libraries upstream
// calls: Pillow 12.2.0 (libImaging BoxBlur.c, Blend.c, Convert.c) and torchvision's PIL branch of adjust_brightness /
// adjust_saturation / adjust_hue / adjust_contrast.  Every operation is integer or single-rounding float arithmetic on bytes, so
// "the same bytes as PIL" is a well-defined target; tests/test_feed_augment.py runs these kernels on the CPU emulator against
// Pillow / torchvision themselves (the colour conversions over all 2^24 colours).
//
// GaussianBlur(r) in Pillow = three box blurs per axis with the real-valued box radius R(r) of Gwosdek et al. (host helper
// hoisdf_gaussian_blur_params restates `_gaussian_blur_radius` in C's float / double mixture).  One box-blur pass of a line is
//     out[x] = (ww * sum_{|k| <= n} in[clamp(x + k)] + fw * (in[clamp(x - n - 1)] + in[clamp(x + n + 1)]) + 2^23) >> 24
// with n = (int) R, ww = (uint32)(2^24 / (2 R + 1)) (float division), fw = (2^24 - (2 n + 1) ww) / 2 and edge replication:
// Pillow's running accumulator (ImagingLineBoxBlur8/32) computes exactly this window sum, so the passes are evaluated here as
// direct taps from shared memory -- all three passes of an axis in one kernel, the line (or a strip of columns) resident in
// shared memory between them; intermediate results are rounded to bytes after every pass as in Pillow.
// Upstream's radius is < 0.5 (n = 0: a 3-tap filter); larger radii cost 2 n + 3 taps per byte and pass.
#include <cmath>

#include "common.cuh"

namespace hoisdf {
namespace {

struct BlurParams { uint32_t n, ww, fw; };

__device__ inline uint8_t box_tap(const uint8_t* line, int pos, int len, int step, BlurParams p) {
  // line[i * step] = element i of the line; clamp = edge replication
  uint32_t acc = 0;
  const int n = static_cast<int>(p.n);
  for (int k = -n; k <= n; ++k) acc += line[min(max(pos + k, 0), len - 1) * step];
  const uint32_t far = static_cast<uint32_t>(line[max(pos - n - 1, 0) * step]) + line[min(pos + n + 1, len - 1) * step];
  return static_cast<uint8_t>((acc * p.ww + far * p.fw + (1u << 23)) >> 24);
}

// horizontal: one block per image row; shared = 2 * w * ch bytes (ping-pong); passes box blurs along x
__global__ void __launch_bounds__(256)
blur_rows_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int h, int w, int ch,
                 const uint32_t* __restrict__ params, int passes) {
  HOISDF_DYNAMIC_SMEM(uint8_t, smem);
  const int b = blockIdx.y, y = blockIdx.x, line = w * ch;
  const BlurParams p{params[b * 3], params[b * 3 + 1], params[b * 3 + 2]};
  const int64_t base = (static_cast<int64_t>(b) * h + y) * line;
  uint8_t* cur = smem;
  uint8_t* nxt = smem + line;
  for (int i = threadIdx.x; i < line; i += blockDim.x) cur[i] = src[base + i];
  __syncthreads();
  for (int pass = 0; pass < passes; ++pass) {
    for (int i = threadIdx.x; i < line; i += blockDim.x) {
      const int x = i / ch, c = i - x * ch;
      nxt[i] = box_tap(cur + c, x, w, ch, p);
    }
    __syncthreads();
    uint8_t* t = cur; cur = nxt; nxt = t;
  }
  for (int i = threadIdx.x; i < line; i += blockDim.x) dst[base + i] = cur[i];
}

// vertical: one block per strip of `strip` byte columns (a byte column is one channel of one pixel column); shared = 2 * h * strip
__global__ void __launch_bounds__(256)
blur_cols_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int h, int line, int strip,
                 const uint32_t* __restrict__ params, int passes) {
  HOISDF_DYNAMIC_SMEM(uint8_t, smem);
  const int b = blockIdx.y, c0 = blockIdx.x * strip, cols = min(strip, line - c0);
  const BlurParams p{params[b * 3], params[b * 3 + 1], params[b * 3 + 2]};
  const int64_t base = static_cast<int64_t>(b) * h * line + c0;
  uint8_t* cur = smem;
  uint8_t* nxt = smem + h * strip;
  for (int i = threadIdx.x; i < h * cols; i += blockDim.x) {
    const int y = i / cols, c = i - y * cols;
    cur[y * strip + c] = src[base + static_cast<int64_t>(y) * line + c];
  }
  __syncthreads();
  for (int pass = 0; pass < passes; ++pass) {
    for (int i = threadIdx.x; i < h * cols; i += blockDim.x) {
      const int y = i / cols, c = i - y * cols;
      nxt[y * strip + c] = box_tap(cur + c, y, h, strip, p);
    }
    __syncthreads();
    uint8_t* t = cur; cur = nxt; nxt = t;
  }
  for (int i = threadIdx.x; i < h * cols; i += blockDim.x) {
    const int y = i / cols, c = i - y * cols;
    dst[base + static_cast<int64_t>(y) * line + c] = cur[y * strip + c];
  }
}

}  // namespace
}  // namespace hoisdf

using namespace hoisdf;

// Pillow's `_gaussian_blur_radius` (libImaging/BoxBlur.c) and the two fixed-point weights of ImagingHorizontalBoxBlur, in the
// same C types: float variables, double constants in the sqrt / floor lines.  out = {n, ww, fw}.  Host function (no GPU work).
HOISDF_API int hoisdf_gaussian_blur_params(float radius, int32_t passes, uint32_t* out) {
  if (out == nullptr) return HOISDF_E_NULL;
  if (!(radius >= 0.0f) || passes <= 0) return HOISDF_E_SHAPE;
  volatile float sigma2 = radius * radius / passes;
  volatile float L = static_cast<float>(std::sqrt(12.0 * sigma2 + 1.0));
  volatile float l = static_cast<float>(std::floor((L - 1.0) / 2.0));
  volatile float t0 = 2 * l + 1, t1 = l * (l + 1), t2 = 3 * sigma2;
  volatile float t3 = t1 - t2;
  volatile float a = t0 * t3;
  volatile float t4 = (l + 1) * (l + 1);
  volatile float t5 = sigma2 - t4;
  volatile float t6 = 6 * t5;
  a = a / t6;
  volatile float box = l + a;
  const int n = static_cast<int>(box);
  volatile float denom = box * 2 + 1;
  const uint32_t ww = static_cast<uint32_t>(static_cast<float>(1u << 24) / denom);
  const uint32_t fw = ((1u << 24) - static_cast<uint32_t>(n * 2 + 1) * ww) / 2;
  out[0] = static_cast<uint32_t>(n);
  out[1] = ww;
  out[2] = fw;
  return HOISDF_OK;
}

HOISDF_API int hoisdf_gaussian_blur_u8(const uint8_t* src, uint8_t* dst, uint8_t* scratch, int64_t batch, int64_t h, int64_t w,
                                       int64_t channels, const uint32_t* params, int32_t passes, void* stream) {
  if (src == nullptr || dst == nullptr || scratch == nullptr || params == nullptr) return HOISDF_E_NULL;
  if (batch <= 0 || batch > 65535 || h <= 0 || w <= 0 || (channels != 1 && channels != 3) || passes <= 0 || passes > 8)
    return HOISDF_E_SHAPE;
  const int64_t line = w * channels;
  if (2 * line > 48 * 1024) return HOISDF_E_SHAPE;
  int strip = 32;
  while (strip > 4 && 2 * h * strip > 48 * 1024) strip /= 2;
  if (2 * h * strip > 48 * 1024) return HOISDF_E_SHAPE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const dim3 g1(static_cast<unsigned>(h), static_cast<unsigned>(batch));
  HOISDF_LAUNCH_SMEM(blur_rows_kernel, g1, 256, static_cast<size_t>(2 * line), s, src, scratch, static_cast<int>(h),
                     static_cast<int>(w), static_cast<int>(channels), params, passes);
  const dim3 g2(static_cast<unsigned>(ceil_div(line, static_cast<int64_t>(strip))), static_cast<unsigned>(batch));
  HOISDF_LAUNCH_SMEM(blur_cols_kernel, g2, 256, static_cast<size_t>(2 * h * strip), s, scratch, dst, static_cast<int>(h),
                     static_cast<int>(line), strip, params, passes);
  return launch_status();
}
