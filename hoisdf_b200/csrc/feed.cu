// Image crop of the data feed (upstream data/ho3d.py:401-427 `data_crop` and :351-353 of `data_aug`,
// data/dexycb.py likewise): `dataset_util.transform_img` (data/dataset_util.py:44-51) = PIL `Image.transform(res, AFFINE,
// inverse coefficients)` with PIL's default NEAREST resampling, `.crop((0, 0, res, res))`, then
// `ToTensor()(np.asarray(img).astype(np.float32)) / 255.0` (ho3d.py:550,624): a pure byte gather -- here one kernel for a whole
// batch of frames already in device memory, BIT-EXACT with Pillow (12.2.0, libImaging/Geometry.c), whose two code paths for
// 8-bit images are restated:
//   * no rotation / shear (coefficient[1] == 0 and coefficient[3] == 0: the evaluation crop):
//     ImagingScaleAffine -- source column / row tables built by REPEATED double-precision addition
//        xo = a2 + a0 * 0.5;  for x: xin = COORD(xo); xo += a0        COORD(v) = v < 0 ? -1 : (int) v
//     (the running sum's rounding is part of the result, so the tables are produced sequentially by one thread each);
//   * general affine (the rotation augmentation): affine_fixed -- 16.16 fixed point with wrapping int32 arithmetic
//        FIX(v) = FLOOR(v * 65536 + 0.5);  a2' = FIX(a2 + (a0 * 0.5 + a1 * 0.5));  xx(x, y) = a2' + y * FIX(a1) + x * FIX(a0);
//        xin = xx >> 16   (same for the row with a5, a3, a4)
//     taken by Pillow when the four output corners map to |coordinates| < 32768 (checked by the caller,
//     hoisdf_b200/feed.py); its floating-point loop for larger coordinates is not restated.
// Pixels that map outside the source are 0 (PIL fills a new image with zeros).
// `channels` = 3 (RGB frames) or 1 (mode "L": the hand / object segmentation masks of the training feed, ho3d.py:366-381,
// which upstream warps with the same call and then shrinks with `resize((64, 64), Image.NEAREST)` = the scale-only path again
// with coefficients (w_in / w_out, 0, 0, 0, h_in / h_out, 0)); `divisor` = 255 for images (`ToTensor(...) / 255.0`), 1 for masks.
// HBM-bound byte work: 3 bytes read (scattered rows, contiguous along x for the crop) and 12 bytes written per output pixel.
#include "common.cuh"

namespace hoisdf {
namespace {

__host__ __device__ inline int pil_floor(double v) { return v < 0.0 ? static_cast<int>(floor(v)) : static_cast<int>(v); }
__host__ __device__ inline int pil_coord(double v) { return v < 0.0 ? -1 : static_cast<int>(v); }
__host__ __device__ inline int pil_fix(double v) { return pil_floor(v * 65536.0 + 0.5); }

// tables[(b * 2 + 0) * size + x] = source column of output column x, [(b * 2 + 1) * size + y] = source row (-1: outside);
// one block of 64 threads per sample, thread 0 walks the columns, thread 32 the rows
__global__ void crop_tables_kernel(const double* __restrict__ coef, int size, int src_w, int src_h, int* __restrict__ tables) {
  const int b = blockIdx.x;
  const double* a = coef + static_cast<int64_t>(b) * 6;
  if (a[1] != 0.0 || a[3] != 0.0) return;                 // general affine: no tables
  if (threadIdx.x != 0 && threadIdx.x != 32) return;
  const bool rows = threadIdx.x == 32;
  const double step = rows ? a[4] : a[0];
  double o = (rows ? a[5] : a[2]) + step * 0.5;
  const int lim = rows ? src_h : src_w;
  int* t = tables + (static_cast<int64_t>(b) * 2 + (rows ? 1 : 0)) * size;
  for (int i = 0; i < size; ++i) {
    const int v = pil_coord(o);
    t[i] = (v >= 0 && v < lim) ? v : -1;
    o += step;
  }
}

// one thread per output pixel; grid (ceil(size * size / 256), batch)
__global__ void __launch_bounds__(256)
affine_crop_kernel(const uint8_t* __restrict__ src, int64_t src_pitch, int64_t src_stride, int src_w, int src_h, int channels,
                   const double* __restrict__ coef, const int* __restrict__ tables, int size, float divisor,
                   float* __restrict__ out_f32, uint8_t* __restrict__ out_u8) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= size * size) return;
  const int y = i / size, x = i - y * size;
  const double* a = coef + static_cast<int64_t>(b) * 6;
  int xin, yin;
  if (a[1] == 0.0 && a[3] == 0.0) {
    xin = tables[(static_cast<int64_t>(b) * 2) * size + x];
    yin = tables[(static_cast<int64_t>(b) * 2 + 1) * size + y];
  } else {
    const unsigned a0 = static_cast<unsigned>(pil_fix(a[0])), a1 = static_cast<unsigned>(pil_fix(a[1]));
    const unsigned a3 = static_cast<unsigned>(pil_fix(a[3])), a4 = static_cast<unsigned>(pil_fix(a[4]));
    const unsigned a2 = static_cast<unsigned>(pil_fix(a[2] + (a[0] * 0.5 + a[1] * 0.5)));
    const unsigned a5 = static_cast<unsigned>(pil_fix(a[5] + (a[3] * 0.5 + a[4] * 0.5)));
    // (unsigned: the C loop's int additions wrap; two's complement makes the sums identical)
    const int xx = static_cast<int>(a2 + static_cast<unsigned>(y) * a1 + static_cast<unsigned>(x) * a0);
    const int yy = static_cast<int>(a5 + static_cast<unsigned>(y) * a4 + static_cast<unsigned>(x) * a3);
    xin = xx >> 16;
    yin = yy >> 16;
    if (xin < 0 || xin >= src_w || yin < 0 || yin >= src_h) xin = yin = -1;
  }
  uint8_t px[3] = {0, 0, 0};
  if (xin >= 0 && yin >= 0) {
    const uint8_t* p = src + b * src_stride + yin * src_pitch + static_cast<int64_t>(xin) * channels;
    for (int c = 0; c < channels; ++c) px[c] = p[c];
  }
  const int64_t plane = static_cast<int64_t>(size) * size;
  if (out_f32 != nullptr) {          // ToTensor layout (channels, size, size), value / divisor in fp32 (IEEE division)
    float* o = out_f32 + static_cast<int64_t>(b) * channels * plane + i;
    for (int c = 0; c < channels; ++c) o[c * plane] = __fdiv_rn(static_cast<float>(px[c]), divisor);
  }
  if (out_u8 != nullptr) {           // the PIL image itself: (size, size, channels) bytes
    uint8_t* o = out_u8 + (static_cast<int64_t>(b) * plane + i) * channels;
    for (int c = 0; c < channels; ++c) o[c] = px[c];
  }
}

}  // namespace
}  // namespace hoisdf

using namespace hoisdf;

HOISDF_API int hoisdf_image_crop_fwd(const uint8_t* src, int64_t batch, int64_t src_h, int64_t src_w, int64_t channels,
                                     int64_t src_pitch, int64_t src_stride, const double* coef, int64_t size, float divisor,
                                     float* out_f32, uint8_t* out_u8, int32_t* tables, void* stream) {
  if (src == nullptr || coef == nullptr || tables == nullptr || (out_f32 == nullptr && out_u8 == nullptr)) return HOISDF_E_NULL;
  if (batch <= 0 || batch > 65535 || src_h <= 0 || src_w <= 0 || src_h > 32767 || src_w > 32767 || size <= 0 || size > 8192 ||
      (channels != 1 && channels != 3) || !(divisor > 0.0f) || src_pitch < channels * src_w || src_stride < src_pitch * src_h)
    return HOISDF_E_SHAPE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  HOISDF_LAUNCH(crop_tables_kernel, static_cast<unsigned>(batch), 64, s, coef, static_cast<int>(size), static_cast<int>(src_w),
                static_cast<int>(src_h), tables);
  const dim3 grid(static_cast<unsigned>(ceil_div(size * size, 256)), static_cast<unsigned>(batch));
  HOISDF_LAUNCH(affine_crop_kernel, grid, 256, s, src, src_pitch, src_stride, static_cast<int>(src_w), static_cast<int>(src_h),
                static_cast<int>(channels), coef, tables, static_cast<int>(size), divisor, out_f32, out_u8);
  return launch_status();
}
