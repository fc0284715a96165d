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
#include <algorithm>
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

// ---- colour jitter (data/dataset_util.py:167-201: up to four torchvision adjustments in a shuffled order) -------------------
// op codes of one step, in upstream's creation order; factor = the adjustment's factor (hue: the byte added to H, 0..255)
enum { JIT_NONE = 0, JIT_BRIGHTNESS = 1, JIT_SATURATION = 2, JIT_HUE = 3, JIT_CONTRAST = 4 };

// Pillow's RGB -> L (Convert.c `L24`)
__device__ inline int luma(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// Pillow's ImagingBlend on one byte (Blend.c): in1 + alpha * (in2 - in1) in float, truncated; clipped when alpha is outside [0, 1]
__device__ inline uint8_t blend(int degenerate, int value, float alpha, bool inside) {
  const float t = __fadd_rn(static_cast<float>(degenerate), __fmul_rn(alpha, static_cast<float>(value - degenerate)));
  if (inside) return static_cast<uint8_t>(t);
  return t <= 0.0f ? 0 : (t >= 255.0f ? 255 : static_cast<uint8_t>(t));
}

__device__ inline int clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// Pillow's rgb2hsv_row (Convert.c): float variables, double constants
__device__ inline void rgb_to_hsv(int r, int g, int b, int& uh, int& us, int& uv) {
  const int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
  uv = maxc;
  if (minc == maxc) { uh = 0; us = 0; return; }
  const float cr = static_cast<float>(maxc - minc);
  const float s = __fdiv_rn(cr, static_cast<float>(maxc));
  const float rc = __fdiv_rn(static_cast<float>(maxc - r), cr), gc = __fdiv_rn(static_cast<float>(maxc - g), cr),
              bc = __fdiv_rn(static_cast<float>(maxc - b), cr);
  float h;
  if (r == maxc) h = __fsub_rn(bc, gc);
  else if (g == maxc) h = __double2float_rn(__dsub_rn(__dadd_rn(2.0, static_cast<double>(rc)), static_cast<double>(bc)));
  else h = __double2float_rn(__dsub_rn(__dadd_rn(4.0, static_cast<double>(gc)), static_cast<double>(rc)));
  h = __double2float_rn(fmod(__dadd_rn(__ddiv_rn(static_cast<double>(h), 6.0), 1.0), 1.0));
  uh = clip8(static_cast<int>(__dmul_rn(static_cast<double>(h), 255.0)));
  us = clip8(static_cast<int>(__dmul_rn(static_cast<double>(s), 255.0)));
}

// Pillow's hsv2rgb (Convert.c)
__device__ inline void hsv_to_rgb(int h, int s, int v, int& r, int& g, int& b) {
  if (s == 0) { r = g = b = v; return; }
  const double h6 = __ddiv_rn(__dmul_rn(static_cast<double>(static_cast<float>(h)), 6.0), 255.0);
  const int i = static_cast<int>(floor(h6));
  const float f = __double2float_rn(__dsub_rn(h6, static_cast<double>(static_cast<float>(i))));
  const float fs = __double2float_rn(__ddiv_rn(static_cast<double>(static_cast<float>(s)), 255.0));
  const double vd = static_cast<double>(static_cast<float>(v)), fsd = static_cast<double>(fs), fd = static_cast<double>(f);
  const int p = clip8(static_cast<int>(round(__dmul_rn(vd, __dsub_rn(1.0, fsd)))));
  const int q = clip8(static_cast<int>(round(__dmul_rn(vd, __dsub_rn(1.0, __dmul_rn(fsd, fd))))));
  const int t = clip8(static_cast<int>(round(__dmul_rn(vd, __dsub_rn(1.0, __dmul_rn(fsd, __dsub_rn(1.0, fd)))))));
  switch (i % 6) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

// sum of L over the image for the samples whose op at `step` is the contrast adjustment (ImageEnhance.Contrast needs the
// mean of the CURRENT image's grey version); integer partial sums -> one 64-bit atomic per block: order-independent
__global__ void __launch_bounds__(256)
jitter_luma_sum_kernel(const uint8_t* __restrict__ img, int64_t pixels, const int32_t* __restrict__ ops, int step,
                       unsigned long long* __restrict__ sums) {
  const int b = blockIdx.y;
  if (ops[b * 4 + step] != JIT_CONTRAST) return;
  __shared__ unsigned int warp_sums[8];
  unsigned int local = 0;
  const uint8_t* p = img + static_cast<int64_t>(b) * pixels * 3;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < pixels;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    local += static_cast<unsigned int>(luma(p[i * 3], p[i * 3 + 1], p[i * 3 + 2]));
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long total = 0;
    for (int w = 0; w < 8; ++w) total += warp_sums[w];
    atomicAdd(&sums[b], total);
  }
}

__global__ void __launch_bounds__(256)
jitter_apply_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t pixels, const int32_t* __restrict__ ops,
                    const float* __restrict__ factors, int step, const unsigned long long* __restrict__ sums) {
  const int b = blockIdx.y;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= pixels) return;
  const int op = ops[b * 4 + step];
  const float factor = factors[b * 4 + step];
  const uint8_t* p = src + (static_cast<int64_t>(b) * pixels + i) * 3;
  uint8_t* o = dst + (static_cast<int64_t>(b) * pixels + i) * 3;
  int r = p[0], g = p[1], bl = p[2];
  const bool inside = factor >= 0.0f && factor <= 1.0f;
  if (op == JIT_BRIGHTNESS) {                       // ImageEnhance.Brightness: blend with black
    r = blend(0, r, factor, inside); g = blend(0, g, factor, inside); bl = blend(0, bl, factor, inside);
  } else if (op == JIT_SATURATION) {                // ImageEnhance.Color: blend with the grey version
    const int l = luma(r, g, bl);
    r = blend(l, r, factor, inside); g = blend(l, g, factor, inside); bl = blend(l, bl, factor, inside);
  } else if (op == JIT_CONTRAST) {                  // ImageEnhance.Contrast: blend with the mean grey, int(mean + 0.5)
    const int m = static_cast<int>(__dadd_rn(__ddiv_rn(static_cast<double>(sums[b]), static_cast<double>(pixels)), 0.5));
    r = blend(m, r, factor, inside); g = blend(m, g, factor, inside); bl = blend(m, bl, factor, inside);
  } else if (op == JIT_HUE) {                       // torchvision adjust_hue: H of Pillow's HSV shifted with byte wrap-around
    int h, s, v;
    rgb_to_hsv(r, g, bl, h, s, v);
    h = (h + static_cast<int>(factor)) & 255;
    hsv_to_rgb(h, s, v, r, g, bl);
  }
  o[0] = static_cast<uint8_t>(r); o[1] = static_cast<uint8_t>(g); o[2] = static_cast<uint8_t>(bl);
}

// ---- the whole training image in ONE kernel ------------------------------------------------------------------------------
// frame -> affine warp -> GaussianBlur -> colour jitter -> ToTensor / 255 (upstream ho3d.py:351-364,550) with ONE CTA per
// image and the warped res x res x 3 image RESIDENT IN SHARED MEMORY from the gather to the final store: 256 x 772 bytes =
// 193 KB of the SM's 227 KB (rows padded to an odd number of 32-bit words so that walking a column is bank-conflict free).
// HBM traffic per image: the gathered source bytes in, 12 bytes per pixel out -- against 13 launches and ~ 40 bytes per pixel
// through L2 / HBM for the step-by-step entry points above (which stay: masks, other sizes, box radii >= 1).  Same arithmetic,
// same bytes: the blur's three passes per axis run IN PLACE with a sliding (previous, current, next) window per row / byte
// column (box radius < 1: a 3-tap filter, all of upstream's radii), the jitter steps run element-wise with a block-wide
// integer reduction for the contrast mean.
struct TrainImageArgs {
  const uint8_t* src; int64_t src_pitch, src_stride; int src_w, src_h;
  const double* coef; const int32_t* mirror; const uint32_t* blur; const int32_t* ops; const float* factors;
  int res, row; float* out_f32; uint8_t* out_u8;
};

__device__ inline int aug_floor(double v) { return v < 0.0 ? static_cast<int>(floor(v)) : static_cast<int>(v); }
__device__ inline int aug_fix(double v) { return aug_floor(v * 65536.0 + 0.5); }          // Pillow's FIX (Geometry.c)

__device__ inline void blur_line_in_place(uint8_t* p, int len, int step, uint32_t ww, uint32_t fw) {
  uint32_t prev = p[0], cur = p[0];
  for (int i = 0; i < len; ++i) {
    const uint32_t nxt = p[min(i + 1, len - 1) * step];
    p[i * step] = static_cast<uint8_t>((cur * ww + (prev + nxt) * fw + (1u << 23)) >> 24);
    prev = cur;
    cur = nxt;
  }
}

__global__ void __launch_bounds__(1024) train_image_kernel(TrainImageArgs a) {
  HOISDF_DYNAMIC_SMEM(uint8_t, smem);
  __shared__ unsigned int warp_sums[32];
  __shared__ int mean_grey;
  const int b = blockIdx.x, res = a.res, row = a.row, tid = threadIdx.x, nthr = blockDim.x;
  int* tab = reinterpret_cast<int*>(smem + static_cast<size_t>(res) * row);      // 2 * res source columns / rows (scale-only crops)
  const double* c = a.coef + static_cast<int64_t>(b) * 6;
  const bool scale_only = c[1] == 0.0 && c[3] == 0.0;
  // 1. Pillow's ImagingScaleAffine tables: repeated double additions, one thread per table
  if (scale_only && (tid == 0 || tid == 32)) {
    const bool rows = tid == 32;
    const double stepd = rows ? c[4] : c[0];
    double o = (rows ? c[5] : c[2]) + stepd * 0.5;
    const int lim = rows ? a.src_h : a.src_w;
    for (int i = 0; i < res; ++i) {
      const int v = o < 0.0 ? -1 : static_cast<int>(o);
      tab[(rows ? res : 0) + i] = (v >= 0 && v < lim) ? v : -1;
      o += stepd;
    }
  }
  __syncthreads();
  // 2. the warp: gather the frame's bytes into shared memory
  const bool flip = a.mirror != nullptr && a.mirror[b] != 0;
  const unsigned a0 = static_cast<unsigned>(aug_fix(c[0])), a1 = static_cast<unsigned>(aug_fix(c[1]));
  const unsigned a3 = static_cast<unsigned>(aug_fix(c[3])), a4 = static_cast<unsigned>(aug_fix(c[4]));
  const unsigned a2 = static_cast<unsigned>(aug_fix(c[2] + (c[0] * 0.5 + c[1] * 0.5)));
  const unsigned a5 = static_cast<unsigned>(aug_fix(c[5] + (c[3] * 0.5 + c[4] * 0.5)));
  for (int i = tid; i < res * res; i += nthr) {
    const int y = i / res, x = i - y * res;
    int xin, yin;
    if (scale_only) {
      xin = tab[x];
      yin = tab[res + y];
    } else {
      xin = static_cast<int>(a2 + static_cast<unsigned>(y) * a1 + static_cast<unsigned>(x) * a0) >> 16;
      yin = static_cast<int>(a5 + static_cast<unsigned>(y) * a4 + static_cast<unsigned>(x) * a3) >> 16;
      if (xin < 0 || xin >= a.src_w || yin < 0 || yin >= a.src_h) xin = yin = -1;
    }
    uint8_t* o = smem + y * row + x * 3;
    if (xin >= 0 && yin >= 0) {
      if (flip) xin = a.src_w - 1 - xin;
      const uint8_t* p = a.src + b * a.src_stride + yin * a.src_pitch + static_cast<int64_t>(xin) * 3;
      o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
    } else {
      o[0] = o[1] = o[2] = 0;
    }
  }
  __syncthreads();
  // 3. GaussianBlur: three box passes along x, then three along y, in place
  const uint32_t ww = a.blur[b * 3 + 1], fw = a.blur[b * 3 + 2];
  if (fw != 0) {                                                     // (fw == 0: radius 0, the identity)
    for (int pass = 0; pass < 3; ++pass) {
      for (int t = tid; t < res * 3; t += nthr) blur_line_in_place(smem + (t / 3) * row + (t % 3), res, 3, ww, fw);
      __syncthreads();
    }
    for (int pass = 0; pass < 3; ++pass) {
      for (int t = tid; t < res * 3; t += nthr) blur_line_in_place(smem + t, res, row, ww, fw);
      __syncthreads();
    }
  }
  // 4. colour jitter: up to four adjustments in the sample's order
  for (int step = 0; step < 4; ++step) {
    const int op = a.ops[b * 4 + step];
    if (op == JIT_NONE) continue;                                     // (uniform over the block)
    const float factor = a.factors[b * 4 + step];
    const bool inside = factor >= 0.0f && factor <= 1.0f;
    if (op == JIT_CONTRAST) {
      unsigned int local = 0;
      for (int i = tid; i < res * res; i += nthr) {
        const uint8_t* p = smem + (i / res) * row + (i % res) * 3;
        local += static_cast<unsigned int>(luma(p[0], p[1], p[2]));
      }
      for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
      if ((tid & 31) == 0) warp_sums[tid >> 5] = local;
      __syncthreads();
      if (tid == 0) {
        unsigned long long total = 0;
        for (int w = 0; w < (nthr >> 5); ++w) total += warp_sums[w];
        mean_grey = static_cast<int>(__dadd_rn(__ddiv_rn(static_cast<double>(total), static_cast<double>(res * res)), 0.5));
      }
      __syncthreads();
    }
    const int shift = static_cast<int>(factor);
    for (int i = tid; i < res * res; i += nthr) {
      uint8_t* p = smem + (i / res) * row + (i % res) * 3;
      int r = p[0], g = p[1], bl = p[2];
      if (op == JIT_HUE) {
        int h, s, v;
        rgb_to_hsv(r, g, bl, h, s, v);
        hsv_to_rgb((h + shift) & 255, s, v, r, g, bl);
      } else {
        const int l = op == JIT_BRIGHTNESS ? 0 : (op == JIT_SATURATION ? luma(r, g, bl) : mean_grey);
        r = blend(l, r, factor, inside); g = blend(l, g, factor, inside); bl = blend(l, bl, factor, inside);
      }
      p[0] = static_cast<uint8_t>(r); p[1] = static_cast<uint8_t>(g); p[2] = static_cast<uint8_t>(bl);
    }
    __syncthreads();
  }
  // 5. ToTensor / 255: three fp32 planes, coalesced along x (and / or the bytes themselves)
  const int64_t plane = static_cast<int64_t>(res) * res;
  for (int i = tid; i < res * res; i += nthr) {
    const uint8_t* p = smem + (i / res) * row + (i % res) * 3;
    if (a.out_f32 != nullptr) {
      float* o = a.out_f32 + static_cast<int64_t>(b) * 3 * plane + i;
      o[0] = __fdiv_rn(static_cast<float>(p[0]), 255.0f);
      o[plane] = __fdiv_rn(static_cast<float>(p[1]), 255.0f);
      o[2 * plane] = __fdiv_rn(static_cast<float>(p[2]), 255.0f);
    }
    if (a.out_u8 != nullptr) {
      uint8_t* o = a.out_u8 + (static_cast<int64_t>(b) * plane + i) * 3;
      o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
    }
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

HOISDF_API int hoisdf_color_jitter_u8(const uint8_t* src, uint8_t* dst, int64_t batch, int64_t h, int64_t w, const int32_t* ops,
                                      const float* factors, uint64_t* sums, void* stream) {
  if (src == nullptr || dst == nullptr || ops == nullptr || factors == nullptr || sums == nullptr) return HOISDF_E_NULL;
  if (batch <= 0 || batch > 65535 || h <= 0 || w <= 0 || h * w > (int64_t{1} << 24)) return HOISDF_E_SHAPE;   // 32-bit block sums
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t pixels = h * w;
  cudaError_t e = cudaMemsetAsync(sums, 0, static_cast<size_t>(4 * batch) * sizeof(uint64_t), s);
  if (e != cudaSuccess) return static_cast<int>(e);
  const dim3 grid(static_cast<unsigned>(ceil_div(pixels, static_cast<int64_t>(256))), static_cast<unsigned>(batch));
  const dim3 sum_grid(static_cast<unsigned>(std::min<int64_t>(ceil_div(pixels, static_cast<int64_t>(256 * 16)), 64)),
                      static_cast<unsigned>(batch));
  const uint8_t* cur = src;
  for (int step = 0; step < 4; ++step) {
    unsigned long long* step_sums = reinterpret_cast<unsigned long long*>(sums) + static_cast<int64_t>(step) * batch;
    HOISDF_LAUNCH(jitter_luma_sum_kernel, sum_grid, 256, s, cur, pixels, ops, step, step_sums);
    HOISDF_LAUNCH(jitter_apply_kernel, grid, 256, s, cur, dst, pixels, ops, factors, step,
                  static_cast<const unsigned long long*>(step_sums));
    cur = dst;
  }
  return launch_status();
}

// row pitch of the shared-memory image: res * 3 rounded up to whole 32-bit words, an ODD number of them (bank-conflict-free columns)
static int train_image_row_bytes(int64_t res) {
  int64_t words = (res * 3 + 3) / 4;
  if ((words & 1) == 0) ++words;
  return static_cast<int>(words * 4);
}

HOISDF_API int64_t hoisdf_train_image_smem_bytes(int64_t res) {
  if (res <= 0 || res > 4096) return HOISDF_E_SHAPE;
  return res * train_image_row_bytes(res) + 2 * res * static_cast<int64_t>(sizeof(int));
}

HOISDF_API int hoisdf_train_image_fwd(const uint8_t* src, int64_t batch, int64_t src_h, int64_t src_w, int64_t src_pitch,
                                      int64_t src_stride, const double* coef, const int32_t* mirror, const uint32_t* blur,
                                      const int32_t* ops, const float* factors, int64_t res, float* out_f32, uint8_t* out_u8,
                                      void* stream) {
  if (src == nullptr || coef == nullptr || blur == nullptr || ops == nullptr || factors == nullptr ||
      (out_f32 == nullptr && out_u8 == nullptr))
    return HOISDF_E_NULL;
  if (batch <= 0 || src_h <= 0 || src_w <= 0 || src_h > 32767 || src_w > 32767 || res <= 0 || src_pitch < 3 * src_w ||
      src_stride < src_pitch * src_h)
    return HOISDF_E_SHAPE;
  const int64_t smem = hoisdf_train_image_smem_bytes(res);
  if (smem < 0 || smem > 227 * 1024) return HOISDF_E_UNSUPPORTED;        // the image does not fit one SM: use the step-by-step calls
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#ifndef HOISDF_EMULATE
  cudaError_t e = cudaFuncSetAttribute(train_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
#endif
  TrainImageArgs a{src, src_pitch, src_stride, static_cast<int>(src_w), static_cast<int>(src_h), coef, mirror, blur, ops, factors,
                   static_cast<int>(res), train_image_row_bytes(res), out_f32, out_u8};
  HOISDF_LAUNCH_SMEM(train_image_kernel, static_cast<unsigned>(batch), 1024, static_cast<size_t>(smem), s, a);
  return launch_status();
}
