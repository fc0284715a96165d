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
// which upstream warps with the same call and then shrinks with `resize((128, 128), Image.NEAREST)` = the scale-only path again
// with coefficients (w_in / w_out, 0, 0, 0, h_in / h_out, 0)); `divisor` = 255 for images (`ToTensor(...) / 255.0`), 1 for masks.
// `mirror[b]` != 0: the source frame is read left-right mirrored (data/dexycb.py:427-430,479-481: left hands are flipped with
// `img[:, ::-1, :]` before the warp) -- the same pixels as warping a mirrored copy, without making one.
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
template <int channels>
__global__ void __launch_bounds__(256)
affine_crop_kernel(const uint8_t* __restrict__ src, int64_t src_pitch, int64_t src_stride, int src_w, int src_h,
                   const double* __restrict__ coef, const int32_t* __restrict__ mirror, const int* __restrict__ tables,
                   int size, float divisor,
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
  uint8_t px[channels] = {};
  if (xin >= 0 && yin >= 0) {
    if (mirror != nullptr && mirror[b] != 0) xin = src_w - 1 - xin;      // the warp reads the frame's mirror image
    const uint8_t* p = src + b * src_stride + yin * src_pitch + static_cast<int64_t>(xin) * channels;
#pragma unroll
    for (int c = 0; c < channels; ++c) px[c] = p[c];
  }
  const int64_t plane = static_cast<int64_t>(size) * size;
  if (out_f32 != nullptr) {          // ToTensor layout (channels, size, size), value / divisor in fp32 (IEEE division)
    float* o = out_f32 + static_cast<int64_t>(b) * channels * plane + i;
#pragma unroll
    for (int c = 0; c < channels; ++c) o[c * plane] = __fdiv_rn(static_cast<float>(px[c]), divisor);
  }
  if (out_u8 != nullptr) {           // the PIL image itself: (size, size, channels) bytes
    uint8_t* o = out_u8 + (static_cast<int64_t>(b) * plane + i) * channels;
#pragma unroll
    for (int c = 0; c < channels; ++c) o[c] = px[c];
  }
}

// ---- SDF point sets of a training / evaluation sample (upstream data/ho3d.py:484-486 row gather `sdf_data[all_idx]`, :333
// rotation of the augmentation, :524-548 normalisation, :561-579 the `inputs` / `targets` entries; data/dexycb.py:515-548 with
// its mirror flip :547-548) from the packed rows `[x, y, z, sdf_hand, sdf_obj, label]` (tool/pre_process_sdf.py:140-147) of a
// batch of frames resident in device memory.  One thread per selected row: 24 bytes in, 12-16 bytes out.
// The indices are the caller's (`np.random.choice` draws: their values are defined by numpy's generator state); groups along
// the index axis, in upstream's order: [hand (n_hand) | object (n_obj) | hand_pre (n_hand) | obj_pre (n_obj)], the last two
// only when n_sel = 2 * (n_hand + n_obj).  Arithmetic as numpy's float32: rotation = the sgemm accumulation order
// x * r0 -> fma(y, r1, .) -> fma(z, r2, .), subtraction and scaling as separately rounded float32 operations.
struct SdfRowsArgs {
  const float* rows; const int64_t* row_offsets; const int64_t* index; int64_t n_sel; int n_hand, n_obj;
  const float* rot; const int32_t* flip; const float* hand_root; const float* obj_centre; float hand_scale, obj_scale;
  float* hand_points; float* obj_points; float* hand_pre; float* obj_pre; float* hand_sdf; float* obj_sdf; int* status;
};

__global__ void __launch_bounds__(256) sdf_rows_kernel(SdfRowsArgs a) {
  const int b = blockIdx.y;
  const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= a.n_sel) return;
  const int64_t first = a.row_offsets[b], count = a.row_offsets[b + 1] - first;
  const int64_t idx = a.index[b * a.n_sel + j];
  if (idx < 0 || idx >= count) {            // an index outside the frame's rows: flagged, nothing written for it
    atomicExch(a.status, 1);
    return;
  }
  const float2* r2 = reinterpret_cast<const float2*>(a.rows + (first + idx) * 6);       // rows are 24 bytes: 8-byte aligned
  const float2 p01 = __ldg(r2), p23 = __ldg(r2 + 1), p45 = __ldg(r2 + 2);
  float x = p01.x, y = p01.y, z = p23.x;
  if (a.flip != nullptr && a.flip[b] != 0) x = -x;                                        // dexycb.py:547-548
  if (a.rot != nullptr) {                                                                 // ho3d.py:333 `.dot(rot_mat.T)`
    const float* r = a.rot + static_cast<int64_t>(b) * 9;
    const float nx = __fmaf_rn(z, r[2], __fmaf_rn(y, r[1], __fmul_rn(x, r[0])));
    const float ny = __fmaf_rn(z, r[5], __fmaf_rn(y, r[4], __fmul_rn(x, r[3])));
    const float nz = __fmaf_rn(z, r[8], __fmaf_rn(y, r[7], __fmul_rn(x, r[6])));
    x = nx; y = ny; z = nz;
  }
  const int per = a.n_hand + a.n_obj;
  const int g = j >= per ? 2 : 0, k = static_cast<int>(j - (g ? per : 0));
  const bool hand = k < a.n_hand;
  const int slot = hand ? k : k - a.n_hand, n = hand ? a.n_hand : a.n_obj;
  const float* c = (hand ? a.hand_root : a.obj_centre) + static_cast<int64_t>(b) * 3;
  const float sc = hand ? a.hand_scale : a.obj_scale;
  float* dst = g ? (hand ? a.hand_pre : a.obj_pre) : (hand ? a.hand_points : a.obj_points);
  float* o = dst + (static_cast<int64_t>(b) * n + slot) * 3;
  o[0] = __fmul_rn(__fsub_rn(x, c[0]), sc);
  o[1] = __fmul_rn(__fsub_rn(y, c[1]), sc);
  o[2] = __fmul_rn(__fsub_rn(z, c[2]), sc);
  if (g == 0) {                                                                           // ho3d.py:575-576
    if (hand) a.hand_sdf[static_cast<int64_t>(b) * n + slot] = __fmul_rn(p23.y, sc);
    else a.obj_sdf[static_cast<int64_t>(b) * n + slot] = __fmul_rn(p45.x, sc);
  }
}

// ---- segmentation masks of a training / DexYCB sample in ONE launch: `transform_img(mask)` to (res, res), then
// `.resize((out_res, out_res), Image.NEAREST)`, then `.astype(float32)` (ho3d.py:366-381,551-552; dexycb.py:323-336,389-402).
// One CTA per mask.  The shrink is Pillow's scale-only path with step res / out_res: its two tables (which warped column / row
// each output column / row takes) and, for an un-rotated warp, the warp's two tables are built in shared memory by four threads
// (repeated double-precision additions, as Pillow builds them); every output pixel is then ONE byte gathered from the frame.
__global__ void __launch_bounds__(1024)
mask_crop_kernel(const uint8_t* __restrict__ src, int64_t src_pitch, int64_t src_stride, int src_w, int src_h,
                 const double* __restrict__ coef, const int32_t* __restrict__ mirror, int res, int out_res,
                 float* __restrict__ out) {
  HOISDF_DYNAMIC_SMEM(int, tabs);                 // [0, out_res): mid column, [out_res, 2 out_res): mid row, then 2 * res warp tables
  const int b = blockIdx.x, tid = threadIdx.x;
  const double* a = coef + static_cast<int64_t>(b) * 6;
  const bool scale_only = a[1] == 0.0 && a[3] == 0.0;
  if (tid < 128 && (tid & 31) == 0) {
    const int which = tid >> 5;                   // 0, 1: shrink columns / rows; 2, 3: warp columns / rows
    if (which < 2) {
      const double step = static_cast<double>(res) / static_cast<double>(out_res);
      double o = 0.0 + step * 0.5;
      for (int i = 0; i < out_res; ++i) {
        const int v = pil_coord(o);
        tabs[which * out_res + i] = (v >= 0 && v < res) ? v : -1;
        o += step;
      }
    } else if (scale_only) {
      const bool rows = which == 3;
      const double step = rows ? a[4] : a[0];
      double o = (rows ? a[5] : a[2]) + step * 0.5;
      const int lim = rows ? src_h : src_w;
      int* t = tabs + 2 * out_res + (rows ? res : 0);
      for (int i = 0; i < res; ++i) {
        const int v = pil_coord(o);
        t[i] = (v >= 0 && v < lim) ? v : -1;
        o += step;
      }
    }
  }
  __syncthreads();
  const unsigned a0 = static_cast<unsigned>(pil_fix(a[0])), a1 = static_cast<unsigned>(pil_fix(a[1]));
  const unsigned a3 = static_cast<unsigned>(pil_fix(a[3])), a4 = static_cast<unsigned>(pil_fix(a[4]));
  const unsigned a2 = static_cast<unsigned>(pil_fix(a[2] + (a[0] * 0.5 + a[1] * 0.5)));
  const unsigned a5 = static_cast<unsigned>(pil_fix(a[5] + (a[3] * 0.5 + a[4] * 0.5)));
  const bool flip = mirror != nullptr && mirror[b] != 0;
  for (int i = tid; i < out_res * out_res; i += blockDim.x) {
    const int oy = i / out_res, ox = i - oy * out_res;
    const int mx = tabs[ox], my = tabs[out_res + oy];                 // pixel of the (res, res) warp this output takes
    uint8_t v = 0;
    if (mx >= 0 && my >= 0) {
      int xin, yin;
      if (scale_only) {
        xin = tabs[2 * out_res + mx];
        yin = tabs[2 * out_res + res + my];
      } else {
        xin = static_cast<int>(a2 + static_cast<unsigned>(my) * a1 + static_cast<unsigned>(mx) * a0) >> 16;
        yin = static_cast<int>(a5 + static_cast<unsigned>(my) * a4 + static_cast<unsigned>(mx) * a3) >> 16;
        if (xin < 0 || xin >= src_w || yin < 0 || yin >= src_h) xin = yin = -1;
      }
      if (xin >= 0 && yin >= 0) {
        if (flip) xin = src_w - 1 - xin;
        v = src[b * src_stride + yin * src_pitch + xin];
      }
    }
    out[static_cast<int64_t>(b) * out_res * out_res + i] = static_cast<float>(v);
  }
}

}  // namespace
}  // namespace hoisdf

using namespace hoisdf;

HOISDF_API int hoisdf_image_crop_fwd(const uint8_t* src, int64_t batch, int64_t src_h, int64_t src_w, int64_t channels,
                                     int64_t src_pitch, int64_t src_stride, const double* coef, const int32_t* mirror,
                                     int64_t size, float divisor, float* out_f32, uint8_t* out_u8, int32_t* tables, void* stream) {
  if (src == nullptr || coef == nullptr || tables == nullptr || (out_f32 == nullptr && out_u8 == nullptr)) return HOISDF_E_NULL;
  if (batch <= 0 || batch > 65535 || src_h <= 0 || src_w <= 0 || src_h > 32767 || src_w > 32767 || size <= 0 || size > 8192 ||
      (channels != 1 && channels != 3) || !(divisor > 0.0f) || src_pitch < channels * src_w || src_stride < src_pitch * src_h)
    return HOISDF_E_SHAPE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  HOISDF_LAUNCH(crop_tables_kernel, static_cast<unsigned>(batch), 64, s, coef, static_cast<int>(size), static_cast<int>(src_w),
                static_cast<int>(src_h), tables);
  const dim3 grid(static_cast<unsigned>(ceil_div(size * size, 256)), static_cast<unsigned>(batch));
  if (channels == 3)
    HOISDF_LAUNCH(affine_crop_kernel<3>, grid, 256, s, src, src_pitch, src_stride, static_cast<int>(src_w),
                  static_cast<int>(src_h), coef, mirror, tables, static_cast<int>(size), divisor, out_f32, out_u8);
  else
    HOISDF_LAUNCH(affine_crop_kernel<1>, grid, 256, s, src, src_pitch, src_stride, static_cast<int>(src_w),
                  static_cast<int>(src_h), coef, mirror, tables, static_cast<int>(size), divisor, out_f32, out_u8);
  return launch_status();
}

HOISDF_API int hoisdf_sdf_rows_fwd(const float* rows, const int64_t* row_offsets, const int64_t* index, int64_t batch,
                                   int64_t n_sel, int64_t n_hand, int64_t n_obj, const float* rot, const int32_t* flip,
                                   const float* hand_root, const float* obj_centre, float hand_scale, float obj_scale,
                                   float* hand_points, float* obj_points, float* hand_pre, float* obj_pre, float* hand_sdf,
                                   float* obj_sdf, int32_t* status, void* stream) {
  if (rows == nullptr || row_offsets == nullptr || index == nullptr || hand_root == nullptr || obj_centre == nullptr ||
      hand_points == nullptr || obj_points == nullptr || hand_sdf == nullptr || obj_sdf == nullptr || status == nullptr)
    return HOISDF_E_NULL;
  if (batch <= 0 || batch > 65535 || n_hand <= 0 || n_obj <= 0 || n_hand + n_obj > (1 << 28) ||
      (n_sel != n_hand + n_obj && n_sel != 2 * (n_hand + n_obj)))
    return HOISDF_E_SHAPE;
  if (n_sel == 2 * (n_hand + n_obj) && (hand_pre == nullptr || obj_pre == nullptr)) return HOISDF_E_NULL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SdfRowsArgs a{rows, row_offsets, index, n_sel, static_cast<int>(n_hand), static_cast<int>(n_obj), rot, flip, hand_root,
                obj_centre, hand_scale, obj_scale, hand_points, obj_points, hand_pre, obj_pre, hand_sdf, obj_sdf, status};
  const dim3 grid(static_cast<unsigned>(ceil_div(n_sel, 256)), static_cast<unsigned>(batch));
  HOISDF_LAUNCH(sdf_rows_kernel, grid, 256, s, a);
  return launch_status();
}

HOISDF_API int hoisdf_mask_crop_fwd(const uint8_t* src, int64_t batch, int64_t src_h, int64_t src_w, int64_t src_pitch,
                                    int64_t src_stride, const double* coef, const int32_t* mirror, int64_t res, int64_t out_res,
                                    float* out, void* stream) {
  if (src == nullptr || coef == nullptr || out == nullptr) return HOISDF_E_NULL;
  if (batch <= 0 || src_h <= 0 || src_w <= 0 || src_h > 32767 || src_w > 32767 || res <= 0 || res > 4096 || out_res <= 0 ||
      out_res > res || src_pitch < src_w || src_stride < src_pitch * src_h)
    return HOISDF_E_SHAPE;
  const size_t smem = static_cast<size_t>(2 * out_res + 2 * res) * sizeof(int);       // <= 64 KB
#ifndef HOISDF_EMULATE
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mask_crop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
  }
#endif
  HOISDF_LAUNCH_SMEM(mask_crop_kernel, static_cast<unsigned>(batch), 1024, smem, static_cast<cudaStream_t>(stream), src, src_pitch,
                     src_stride, static_cast<int>(src_w), static_cast<int>(src_h), coef, mirror, static_cast<int>(res),
                     static_cast<int>(out_res), out);
  return launch_status();
}
