/*
 * hoisdf_b200 -- C ABI of the B200-native (sm_100a) HOISDF hot path.
 *
 * The upstream project (amathislab/HOISDF) has no FFI layer: its operator surface is the Python code
 * in main/model.py and common/nets/*.  Each entry point below replaces the stock-PyTorch call sites
 * of one upstream operator (cited per function); `hoisdf_b200/` binds them with ctypes and re-exposes
 * the upstream signatures (see INTEGRATION.md for the binding a maintainer would add upstream).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 unless stated otherwise; the caller owns all buffers
 *     (inputs, outputs, workspaces); the library allocates nothing and keeps no pointer after return;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no internal synchronisation,
 *     no host read-back;
 *   - sizes are int64_t, leading dimensions are in ELEMENTS;
 *   - return value: 0 = ok, negative = HOISDF_E_* (bad argument), positive = cudaError_t observed after
 *     the launch.  Nothing throws or exits across the boundary.
 */
#ifndef HOISDF_B200_H
#define HOISDF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HOISDF_ABI_VERSION 39

enum {
  HOISDF_OK = 0,
  HOISDF_E_NULL = -1,      /* required pointer is NULL */
  HOISDF_E_SHAPE = -2,     /* size out of the supported range */
  HOISDF_E_ALIGN = -3,     /* pointer or leading dimension not 16-byte aligned */
  HOISDF_E_UNSUPPORTED = -4,
  HOISDF_E_WORKSPACE = -5,        /* caller-owned workspace too small for this input */
  HOISDF_E_TOO_FEW_POINTS = -6    /* a sample has fewer lattice points inside its bbox than num_points (upstream raises too) */
};

enum { HOISDF_ACT_NONE = 0, HOISDF_ACT_RELU = 1 };

int hoisdf_abi_version(void);
const char* hoisdf_status_string(int status);

/* ---------------------------------------------------------------------------------------------------
 * nn.Linear (+ReLU) over rows -- upstream common/nets/layer.py:192-201 (MLP.forward), every
 * `nn.Linear` of common/nets/sdf_net.py:95-107 and common/nets/transformer.py:294-299,378-392.
 *   Y[r, 0:N] = act( X[r, 0:K] . W[0:N, 0:K]^T + bias ) (+ residual[r, 0:N])
 * Rows may be "batched": row r lives at  base + (r / rows_per_batch) * batch_stride + (r % rows_per_batch) * ld
 * (rows_per_batch = 0 means plain `base + r * ld`).  K, ldx, ldw must be multiples of 4 and X, W 16-byte
 * aligned; columns K..ldw of W are never read.  `residual` (optional) shares Y's addressing.
 *
 * Two implementations with the same contract:
 *   w_lo == NULL : fp32 FMA kernel (bit-faithful fp32 products).
 *   w_lo != NULL : tcgen05 tensor-core kernel, 3xTF32 split (fp32-grade accuracy): `w` must then hold the
 *                  TF32-rounded weights and `w_lo` the residual W - w (both from hoisdf_split_tf32, same pitch).
 *                  Input rows may be batched (m must then be a multiple of x_rows_per_batch); batched OUTPUT
 *                  rows always take the FMA kernel.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  const float* x; int64_t ldx; int64_t x_rows_per_batch; int64_t x_batch_stride;
  const float* w; int64_t ldw;
  const float* bias;                 /* may be NULL */
  const float* residual;             /* may be NULL */
  float* y; int64_t ldy; int64_t y_rows_per_batch; int64_t y_batch_stride;
  int64_t m; int64_t n; int64_t k;
  int32_t act;
  const float* w_lo;                 /* may be NULL (see above) */
  int32_t tf32_passes;               /* tensor-core kernel only: 0 or 3 = 3xTF32 (fp32-grade); 1 = one TF32 pass
                                        (~5e-4 relative; used ONLY to pre-screen top-k candidates that an exact
                                        pass re-ranks, never for a result that is returned) */
} hoisdf_linear_args;

int hoisdf_linear_fwd(const hoisdf_linear_args* args, void* stream);

/* Elementwise split of `count` floats: w_hi = round-to-nearest TF32 of w, w_lo = w - w_hi (exact). */
int hoisdf_split_tf32(const float* w, int64_t count, float* w_hi, float* w_lo, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * The same nn.Linear on the tensor cores at FP16 rate with fp32-grade accuracy ("FP16x3", csrc/linear_h3.cu).
 * Activations travel in SPLIT-HALF format: an fp32 value x is two IEEE fp16 numbers
 *     hi = fp16(x),   lo = fp16((x - hi) * 2^11)        x ~= hi + lo * 2^-11      (|x| <= 65504)
 * stored as two planes (hi, lo) of uint16 with a common row pitch (in halfs, multiple of 8; bases 16-byte aligned).
 * Weights are packed once by hoisdf_pack_h3 into three fp16 planes A = w_hi * 2^11, B = w_hi, C = w_lo * 2^11
 * (|w| < 32).  Y = act(X . W^T + bias) (+ residual) is written EITHER as fp32 (`y`, pitch ldy floats) OR in
 * split-half format (`y_hi`, `y_lo`, pitch ldyh halfs; no residual) for the next layer.  X rows may be batched
 * like hoisdf_linear_fwd's (x_batch_stride in halfs); output rows are dense.  K is contracted over [0, k): columns
 * beyond k of X and W are never read (no zero padding needed).
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  const uint16_t* x_hi; const uint16_t* x_lo; int64_t ldx; int64_t x_rows_per_batch; int64_t x_batch_stride;
  const uint16_t* w_a; const uint16_t* w_b; const uint16_t* w_c; int64_t ldw;
  const float* bias;                 /* may be NULL */
  const float* residual;             /* may be NULL; fp32, shares y's addressing */
  float* y; int64_t ldy;             /* fp32 output, or NULL when y_hi / y_lo are given */
  uint16_t* y_hi; uint16_t* y_lo; int64_t ldyh;
  int64_t m; int64_t n; int64_t k;
  int32_t act;
  int32_t chunk_kb;                  /* K blocks (of 32) accumulated in TMEM before the partial sum is drained into
                                        fp32 registers (round-to-nearest adds).  The tensor core truncates on every
                                        accumulate, so the error of a sum held in TMEM grows ~linearly with its
                                        length: 0 = default (4 blocks = 128 of K: fp32-FMA-grade results); 1 = most
                                        accurate (deep convolution chains); >= K/32 = one drain per tile (fastest;
                                        only for values that an exact pass re-ranks) */
  const uint16_t* res_hi; const uint16_t* res_lo; int64_t ldr;
                                     /* optional residual in split-half format, row r of the dense output reads row
                                        r of these planes (pitch ldr halfs); added before the activation; needs the
                                        TMA-store epilogue (no fp32 `residual`, n % 32 == 0) -- the shortcut of the
                                        ResNet bottleneck, upstream common/nets/resnet.py (torchvision Bottleneck) */
  int32_t single_pass;               /* 1: ONE tensor-core product x_hi . w_hi instead of three (11-bit operands,
                                        ~5e-4 relative error, 3x less tensor work): ONLY for pre-screening top-k
                                        candidates that an FP16x3 pass and then an exact pass re-rank, never for a
                                        result that is returned */
  float w_scale;                     /* power-of-two factor the packed weights were divided by before packing (layers
                                        with |w| >= 32, e.g. BatchNorm folds with a tiny running variance); the
                                        epilogue multiplies it back.  0 = 1 */
  const float* y_scale;              /* optional: ONE float in device memory the product x . w^T is multiplied by in the
                                        epilogue (before the bias) -- the power-of-two factor hoisdf_linear_bwd_prep took
                                        out of a gradient, put back without a host read-back or an extra pass */
  int32_t split_k;                   /* 1: a launch with only a handful of output tiles over a long contraction (K >= 2048;
                                        the weight gradients of the training step: 256 x 256 outputs over K = rows) is
                                        split along K over the idle SMs; the partial products are ADDED into the
                                        zero-initialised output by TMA reductions, so the fp32 summation order -- not the
                                        value up to rounding -- varies from run to run.  fp32 `y` output, act == NONE,
                                        no residual; ignored otherwise.  0 (default): never, results are reproducible */
} hoisdf_linear_h3_args;

int hoisdf_linear_h3_fwd(const hoisdf_linear_h3_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Convolution as an implicit GEMM on the same FP16x3 kernel -- the U-Net decoder of upstream
 * common/nets/module.py:147-218 (nn.Conv2d 3x3 / 1x1 and, one output-parity class at a time, nn.ConvTranspose2d
 * k4 s2 p1), BatchNorm folded into weights and bias by the caller.
 *   x: NHWC image batch in split-half format, (batch, in_h, in_w) pixels of ldx halfs each, cin % 32 == 0;
 *   w: hoisdf_pack_h3 planes of a (cout, taps * cin) matrix, K index = tap * cin + channel;
 *   output pixel (b, y, x) of the (out_h, out_w) grid reads input pixel (y * stride + tap_dy[t], x * stride +
 *   tap_dx[t]) for tap t; pixels outside the image contribute zero (= zero padding);
 *   element (b, y, x, c) of the result is written at  base + b * y_sb + y * y_sy + x * y_sx + c  (strides in
 *   elements), fp32 (`y`) or split-half (`y_hi`, `y_lo`) -- strided placement lets a transposed convolution
 *   interleave its four parity classes and lets a layer write into a channel window of a concat buffer.
 * out_w must be a power of two <= 128 or a multiple of 128.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  const uint16_t* x_hi; const uint16_t* x_lo;
  int64_t batch; int64_t in_h; int64_t in_w; int64_t cin; int64_t ldx;
  const uint16_t* w_a; const uint16_t* w_b; const uint16_t* w_c; int64_t ldw;
  const float* bias;                 /* may be NULL */
  int32_t taps; int32_t tap_dy[16]; int32_t tap_dx[16]; int32_t stride;
  int64_t out_h; int64_t out_w; int64_t cout;
  float* y; uint16_t* y_hi; uint16_t* y_lo; int64_t y_sx; int64_t y_sy; int64_t y_sb;
  int32_t act; int32_t chunk_kb;     /* see hoisdf_linear_h3_args */
  const uint16_t* res_hi; const uint16_t* res_lo; int64_t ldr;
                                     /* optional split-half residual: output pixel (b, y, x) reads row
                                        (b * out_h + y) * out_w + x of these planes; cout % 32 == 0 */
  int32_t single_pass;               /* see hoisdf_linear_h3_args */
  float w_scale;                     /* see hoisdf_linear_h3_args */
} hoisdf_conv_h3_args;

int hoisdf_conv_h3_fwd(const hoisdf_conv_h3_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * ResNet-50 stem pieces (upstream common/nets/resnet.py:70-76 = torchvision conv1 7x7 s2 p3 + maxpool 3x3 s2 p1),
 * the rest of the backbone being hoisdf_conv_h3_fwd / hoisdf_linear_h3_fwd calls.
 *   hoisdf_stem_im2col_split: img (batch, 3, h, w) NCHW fp32 -> one row per output pixel of the (h/2, w/2) grid,
 *     160 columns in split-half format: column (ky * 7 + kx) * 3 + c = img[b, c, 2y - 3 + ky, 2x - 3 + kx] (0 outside
 *     the image), columns 147..159 zero.  The 7x7 convolution is then a Linear with K = 160.
 *   hoisdf_maxpool3x3s2_split: NHWC split-half (batch, h, w, c) -> (batch, h/2, w/2, c), window 3x3 stride 2 pad 1
 *     (padding never wins: max over the in-image taps), c % 8 == 0.
 * ------------------------------------------------------------------------------------------------- */
int hoisdf_stem_im2col_split(const float* img, int64_t batch, int64_t h, int64_t w, uint16_t* hi, uint16_t* lo,
                             int64_t ldh, void* stream);
int hoisdf_maxpool3x3s2_split(const uint16_t* x_hi, const uint16_t* x_lo, int64_t ldx, int64_t batch, int64_t h,
                              int64_t w, int64_t c, uint16_t* y_hi, uint16_t* y_lo, int64_t ldy, void* stream);

/* Narrow Linear for the last layer of the small heads (upstream main/model.py:81-90; convOut_* of
 * common/nets/module.py): y (m, n <= 24) = act(x . w^T + bias), x in split-half format (k even), w fp32 (n, ldw),
 * act: 0 none, 1 ReLU, 2 sigmoid.  fp32 FMA arithmetic on the joined values; HBM-bound (one warp per row). */
int hoisdf_linear_narrow_split_fwd(const uint16_t* x_hi, const uint16_t* x_lo, int64_t ldx, int64_t m, const float* w,
                                   int64_t ldw, const float* bias, int64_t n, int64_t k, int32_t act, float* y,
                                   int64_t ldy, void* stream);

/* W (n, ldw) fp32 with k valid columns -> planes A, B, C, each (n, ldh) halfs, columns [k, ldh) zeroed. */
int hoisdf_pack_h3(const float* w, int64_t n, int64_t k, int64_t ldw, uint16_t* w_a, uint16_t* w_b, uint16_t* w_c,
                   int64_t ldh, void* stream);
/* fp32 rows (m, ldx), k valid columns -> split-half planes (m, ldh); columns [k, kpad) are zeroed (kpad % 4 == 0). */
int hoisdf_split_rows(const float* x, int64_t m, int64_t k, int64_t ldx, int64_t kpad, uint16_t* hi, uint16_t* lo,
                      int64_t ldh, void* stream);
/* split-half planes -> fp32 rows. */
int hoisdf_join_rows(const uint16_t* hi, const uint16_t* lo, int64_t ldh, int64_t m, int64_t k, float* x,
                     int64_t ldx, void* stream);

/* nn.utils.weight_norm(dim=0) fold  W = g * v / ||v||_row  (upstream common/nets/sdf_net.py:57-62),
 * written into a (rows, ld_out) matrix at column offset 0; optional column permutation `src_col`
 * (int32[cols_out], -1 = write 0) lets the caller lay the skip-concat layer out for the padded
 * activation layout used by hoisdf_sdf_decoder_fwd. g may be NULL (plain copy/permutation of v). */
int hoisdf_fold_weight_norm(const float* g, const float* v, int64_t rows, int64_t cols, float* out,
                            int64_t ld_out, const int32_t* src_col, int64_t cols_out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Layout: NCHW -> NHWC copy of one pyramid level (upstream keeps NCHW: common/nets/module.py:172-218).
 * ------------------------------------------------------------------------------------------------- */
int hoisdf_nchw_to_nhwc(const float* src, float* dst, int64_t n, int64_t c, int64_t h, int64_t w, void* stream);
/* Same transpose with the result in split-half format: planes (n, h*w, ld halfs), channels at columns [0, c). */
int hoisdf_nchw_to_nhwc_split(const float* src, uint16_t* dst_hi, uint16_t* dst_lo, int64_t n, int64_t c, int64_t h,
                              int64_t w, int64_t ld, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * sdf_infer candidate generation -- upstream main/model.py:257-302: the sheared 64^3 lattice,
 * p_cam = s / sdf_scale + center, pinhole projection with K, STRICT bbox test, stable compaction.
 * Bit-exact with the upstream CPU arithmetic (separately rounded mul/add, IEEE division, fma chain of
 * the 3-term dot products).
 *   center (B,3), cam_intr (B,3,3), bbox (B,4) xyxy.
 * Pass 1 (`_count`): chunk_counts int32 (B, hoisdf_lattice_chunks(bins)) -- per-1024-index chunk counts, turned in
 * place into exclusive global row offsets; offsets int64 (B+1) exclusive prefix of the per-sample N_f.
 * The caller reads offsets[B] (total rows) to size the row buffers, then pass 2 (`_compact`) writes
 *   cand_index int32 (M)   lattice index of every surviving point, sample-major, ascending
 *   cand_uv    fp32  (M,2) its projection in pixels
 * ------------------------------------------------------------------------------------------------- */
int hoisdf_lattice_chunks(int32_t bins);
int hoisdf_lattice_count(const float* center, const float* cam_intr, const float* bbox, float sdf_scale,
                         int64_t batch, int32_t bins, int32_t* chunk_counts, int64_t* offsets, void* stream);
int hoisdf_lattice_compact(const float* center, const float* cam_intr, const float* bbox, float sdf_scale,
                           int64_t batch, int32_t bins, const int32_t* chunk_counts, const int64_t* offsets,
                           int32_t* cand_index, float* cand_uv, void* stream);

/* Pinhole projection of given points -- upstream main/model.py:148-150 / 190-192.
 *   points (B,P,3) normalised coords -> cam (B,P,3) = points/scale + center, uv (B,P,2) */
int hoisdf_project_points(const float* points, const float* center, const float* cam_intr, float sdf_scale,
                          int64_t batch, int64_t p, float* cam, float* uv, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Multi-scale bilinear gather -- upstream F.grid_sample(bilinear, border, align_corners=True) x5 + cat
 * (main/model.py:164-175, 203-214, 316-328).  Maps are NHWC.  Pixel coords use the IMAGE size for every
 * level: g = (uv - (img-1)/2) / ((img-1)/2), x = (g + 1)/2 * (W_l - 1), clamped to [0, W_l - 1].
 * mode CONCAT: out[r, off_l : off_l + C_l] = sample of level l  (the (N, C) matrix upstream feeds its MLPs)
 * mode SUM   : out[r, 0:C] = act(bias + sum_l sample of level l) (all levels carry C channels; used with
 *              the per-level projected maps  G_l = F_l . W0_l^T , which is linear_sdfin layer 0 applied
 *              before instead of after the interpolation -- bilinear sampling is linear)
 * Row r belongs to sample  b = (row_offsets ? upper_bound(row_offsets, r) - 1 : r / rows_per_sample).
 * ------------------------------------------------------------------------------------------------- */
enum { HOISDF_GATHER_CONCAT = 0, HOISDF_GATHER_SUM = 1 };
typedef struct {
  const float* map[5];  /* NHWC (B, H_l, W_l, C_l) */
  int32_t c[5]; int32_t h[5]; int32_t w[5];
  int32_t levels;
  int32_t img_h, img_w;
} hoisdf_pyramid;

int hoisdf_gather_fwd(const hoisdf_pyramid* pyr, const float* uv, int64_t rows, const int64_t* row_offsets,
                      int64_t batch, int64_t rows_per_sample, int32_t mode, const float* bias, int32_t act,
                      float* out, int64_t ld_out, void* stream);
/* Same gather with the output written in split-half format (planes out_hi / out_lo, pitch ld_out halfs, % 4 == 0)
 * for the FP16x3 Linear that follows. */
int hoisdf_gather_split_fwd(const hoisdf_pyramid* pyr, const float* uv, int64_t rows, const int64_t* row_offsets,
                            int64_t batch, int64_t rows_per_sample, int32_t mode, const float* bias, int32_t act,
                            uint16_t* out_hi, uint16_t* out_lo, int64_t ld_out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * SDF decoder input tail: NeRF embedding (upstream common/utils/sdf_utils.py:96-141, 5 octaves, sin then
 * cos per octave) and xyz written at columns [col0, col0+36) of the row buffer: 30 posenc, 3 xyz, 3 zero.
 * Points come either from lattice indices (sdf_infer) or from an explicit (rows,3) array.
 * ------------------------------------------------------------------------------------------------- */
int hoisdf_posenc_fwd(const int32_t* lattice_index, const float* points, int64_t rows, int32_t bins,
                      float* out, int64_t ld_out, int64_t col0, void* stream);
/* Same, written in split-half format into the FP16x3 row buffer (pitch >= 520 halfs per plane):
 * columns [256,286) posenc, [286,289) xyz, [289,296) and [519] zero. */
int hoisdf_posenc_split_fwd(const int32_t* lattice_index, const float* points, int64_t rows, int32_t bins,
                            uint16_t* out_hi, uint16_t* out_lo, int64_t ld_out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * SDFDecoder.forward -- upstream common/nets/sdf_net.py:87-122 (eval: dropout off):
 *   289 -> 512 -> 223 (+289 skip) -> 512 -> 512 -> 1, ReLU after layers 0..3, tanh.
 * x is the padded row buffer (rows, ldx>=516): cols [0,289) decoder input, [289,292) zero,
 * [292,515) scratch for relu(linh1), col 515 zero.  Weights are packed by hoisdf_fold_weight_norm:
 *   w0 (512,292), w1 (223,512), w2 (512,516) [skip-permuted], w3 (512,512), w4 (512).
 * h_a, h_b: two (rows,512) scratch buffers.  out_sdf (rows): tanh output, clamped to +-clamp when clamp > 0
 * (upstream main/model.py:241; sdf_infer clamps after the selection instead, model.py:354).
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  const float* w0; const float* b0;
  const float* w1; const float* b1;
  const float* w2; const float* b2;
  const float* w3; const float* b3;
  const float* w4; const float* b4;
  /* optional TF32 residuals of w0..w3 (all four or none): selects the tensor-core Linear kernel */
  const float* w0_lo; const float* w1_lo; const float* w2_lo; const float* w3_lo;
  int32_t tf32_passes;               /* as in hoisdf_linear_args */
} hoisdf_sdf_weights;

int hoisdf_sdf_decoder_fwd(const hoisdf_sdf_weights* wts, float* x, int64_t ldx, int64_t rows, float* h_a,
                           float* h_b, float* out_sdf, float clamp, void* stream);

/* The same decoder on the FP16x3 tensor-core Linear (hoisdf_linear_h3_fwd), activations in split-half format.
 * x: row buffer, two planes with pitch ldx >= 520 halfs: cols [0,289) decoder input, [289,296) zero, [296,519)
 * scratch for relu(linh1), col 519 zero.  Weights: hoisdf_pack_h3 planes (A, B, C) of w0 (512,289), w1 (223,512),
 * w2 (512,519) [columns permuted to input | 0 | h1], w3 (512,512); fp32 biases; fp32 w4 (512), b4.
 * h_a, h_b: split-half scratch (rows, ldh >= 512). */
typedef struct {
  const uint16_t* w[4][3]; int64_t ldw[4]; const float* b[4];
  const float* w4; const float* b4;
  int32_t chunk_kb;                  /* passed to every hoisdf_linear_h3_fwd of the chain (0 = default) */
  int32_t single_pass;               /* likewise (candidate pre-screening) */
} hoisdf_sdf_weights_h3;

int hoisdf_sdf_decoder_h3_fwd(const hoisdf_sdf_weights_h3* wts, uint16_t* x_hi, uint16_t* x_lo, int64_t ldx,
                              int64_t rows, uint16_t* ha_hi, uint16_t* ha_lo, uint16_t* hb_hi, uint16_t* hb_lo,
                              int64_t ldh, float* out_sdf, float clamp, void* stream);

/* The candidate chain of sdf_infer as ONE persistent tcgen05 kernel (csrc/sdf_chain.cu) -- upstream
 * main/model.py:330-346 (linear_sdfin.layers.1 -> NeRF embedding + xyz -> SDFDecoder, common/nets/sdf_net.py:87-122)
 * on single-product fp16 tensor-core arithmetic (the SCREENING stage of the selection cascade; ~5e-5 absolute).
 * Activations never leave the SM: 4 bytes per row (the SDF value) are written.
 *   rows mode     a0 != NULL: (rows, 512) fp16 = relu(bias + gathered projected maps), i.e. the output of
 *                 linear_sdfin.layers.0; posenc / xyz are computed in the kernel from lattice_index (or points).
 *   decoder mode  x != NULL: (rows, >= 296) fp16 decoder input rows [fea 256 | posenc 30 | xyz 3 | 0 x 7];
 *                 w_s1 / lattice_index / points unused (SDFDecoder.forward in isolation, BASELINE configs[4]).
 * Weights are the fp16 "B" planes (w_hi) that hoisdf_pack_h3 writes: w_s1 (256, >= 512); w[0] linh0 (512, >= 296,
 * zero beyond column 289); w[1] linh1 (223, >= 512); w[2] linh2 (512, >= 520) in the column layout
 * [input 289 | 0 x 7 | h1 223 | 0]; w[3] linh3 (512, >= 512); w4 (512) / b4 (1) fp32 (linh4). */
/* The projected pyramid G_l = F_l . W0_l^T (linear_sdfin.layers.0 applied to every level once per image) as NHWC fp16
 * maps (B, H_l, W_l, c = 512): what the chain kernel's gather mode interpolates. */
typedef struct {
  const uint16_t* map[5];
  int32_t h[5]; int32_t w[5];
  int32_t levels; int32_t c;
  int32_t img_h, img_w;
} hoisdf_pyramid_h;

typedef struct {
  const uint16_t* a0; int64_t lda0;
  const uint16_t* x; int64_t ldx;
  const int32_t* lattice_index; const float* points; int32_t bins;
  const uint16_t* w_s1; int64_t ldw_s1; const float* b_s1;
  const uint16_t* w[4]; int64_t ldw[4]; const float* b[4];
  const float* w4; const float* b4;
  int64_t rows; float clamp; float* out_sdf;
  /* gather mode (a0 == NULL and x == NULL): the kernel itself computes relu(b_s0 + sum_l bilinear sample of gmaps level l)
   * -- upstream's 5 x F.grid_sample + cat + linear_sdfin.layers.0 (main/model.py:316-330) -- for every row from its
   * projected pixel uv (rows, 2); row r belongs to sample upper_bound(row_offsets, r) - 1, or r / rows_per_sample. */
  const hoisdf_pyramid_h* gmaps; const float* uv; const int64_t* row_offsets; int64_t batch; int64_t rows_per_sample;
  const float* b_s0;
} hoisdf_sdf_chain_args;

int hoisdf_sdf_chain_fwd(const hoisdf_sdf_chain_args* args, void* stream);
/* SUM-mode gather (see hoisdf_gather_fwd) on fp16 maps with an fp16 result: out_hi[r, 0:512] = fp16(act(bias + sum_l sample_l)),
 * row pitch ld_out halfs.  The candidate-screening form of upstream main/model.py:316-330 feeding hoisdf_sdf_chain_fwd (rows mode). */
int hoisdf_gather_sum_h16_fwd(const hoisdf_pyramid_h* pyr, const float* uv, int64_t rows, const int64_t* row_offsets,
                              int64_t batch, int64_t rows_per_sample, const float* bias, int32_t act, uint16_t* out_hi,
                              int64_t ld_out, void* stream);
/* fp32 -> fp16 (round to nearest, saturating); n % 8 == 0.  Produces the fp16 copy of the projected maps. */
int hoisdf_f32_to_f16(const float* x, uint16_t* y, int64_t n, void* stream);

/* Expand a plain (rows, 289) decoder input (the upstream SDFDecoder.forward argument) into the padded
 * row buffer layout above. */
int hoisdf_sdf_pad_input(const float* in, int64_t rows, float* x, int64_t ldx, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Near-surface point selection -- upstream main/model.py:345-354: per sample, the `num_points` rows with
 * the smallest |sdf| in ascending order (ties: lower row first), then gather of lattice coords, posenc
 * and the clamped SDF value.
 *   sdf (M) raw decoder output, offsets int64 (B+1), cand_index int32 (M)
 *   -> sel_index int32 (B,P) lattice indices, sel_row int32 (B,P) global row of each pick (optional),
 *      points (B,P,3), out_sdf (B,P) clamped to +-clamp, posenc (B,P,30).
 *      Samples with fewer than P candidates set *status_flag (int32, device) to 1.
 * order_by_row != 0 ("screening" mode): the same SET of P rows, emitted in ascending row order instead of
 * |sdf| order -- used when a tensor-core pass pre-selects P + margin rows that an fp32 pass then re-ranks.
 * ------------------------------------------------------------------------------------------------- */
int hoisdf_select_points(const float* sdf, const int64_t* offsets, const int32_t* cand_index,
                         int64_t batch, int64_t num_points, int32_t bins, float clamp, int32_t order_by_row,
                         int32_t* sel_index, int32_t* sel_row, float* points, float* out_sdf, float* posenc,
                         int32_t* status_flag, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Model.sdf_infer for the whole batch behind ONE entry point -- upstream main/model.py:246-355 (csrc/sdf_infer.cu):
 * candidate generation, the verified coarse-to-fine selection cascade (stage A: fp16 gather + the fused single-product
 * chain kernel on every candidate; final stage: FP16x3 draining TMEM every K block on the P + margin survivors), the
 * device-side verdict of the screening step and the final top-P by |sdf|.
 *   gmaps / gmaps16: the pyramid projected through linear_sdfin.layers.0 (fp32 NHWC (B,H,W,512) and its fp16 copy),
 *     bias0 that layer's bias; s1_*: hoisdf_pack_h3 planes (256, ld_s1) + bias of linear_sdfin.layers.1; dec: the SDF
 *     decoder's planes.
 *   Outputs (device): points (B,P,3) lattice coordinates in selection order, sdf (B,P) clamped to +-clamp, posenc (B,P,30),
 *     sel_index (B,P) lattice indices; screen_err (1) = max |coarse - fine|, screen_gap (B) = coarse rank-(P+margin)
 *     |sdf| minus fine rank-P |sdf|, verified (1) = all(gap > 3 err): when 0 the caller re-runs the general path;
 *     status_flag: as hoisdf_select_points.  n_f (HOST, B, may be NULL): candidate count per sample.
 *   Optional diagnostics (device, may be NULL): cand_sdf / cand_index (total rows), exact_sdf / exact_index / screen_rows
 *     (B * hoisdf_sdf_infer_keep(P, margin)).
 *   Workspace: hoisdf_sdf_infer_workspace_bytes(batch, max_rows, ...) bytes, max_rows >= the total candidate count.
 *   Host interaction: the B + 1 row offsets (they size the launches) are read from `host_offsets` (pinned, caller-owned).
 *     planned = 1: the caller ran hoisdf_sdf_infer_plan earlier on this stream with the same chunk_counts / offsets /
 *     host_offsets and has waited for that copy (e.g. an event recorded right after it, before queueing the image
 *     encoder: nothing stalls); planned = 0: this call plans and synchronises the stream itself.
 *   Returns HOISDF_E_TOO_FEW_POINTS like upstream's shape error (model.py:348), HOISDF_E_UNSUPPORTED when a sample has no
 *   room for the screening margin (the caller then ranks every row exactly), HOISDF_E_WORKSPACE when max_rows is too small.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  const float* center; const float* cam_intr; const float* bbox;
  float sdf_scale; int32_t bins; int64_t batch; int64_t num_points; int64_t margin; float clamp;
  const hoisdf_pyramid* gmaps; const hoisdf_pyramid_h* gmaps16; const float* bias0;
  const uint16_t* s1_a; const uint16_t* s1_b; const uint16_t* s1_c; int64_t ld_s1; const float* b_s1; float s1_scale;
  const hoisdf_sdf_weights_h3* dec;
  void* workspace; int64_t workspace_bytes; int64_t max_rows;
  int32_t planned; int32_t* chunk_counts; int64_t* offsets; int64_t* host_offsets; int64_t* n_f;
  float* points; float* sdf; float* posenc; int32_t* sel_index; int32_t* status_flag;
  float* screen_err; float* screen_gap; int32_t* verified;
  float* cand_sdf; int32_t* cand_index; float* exact_sdf; int32_t* exact_index; int32_t* screen_rows;
} hoisdf_sdf_infer_args;

int64_t hoisdf_sdf_infer_keep(int64_t num_points, int64_t margin);
int64_t hoisdf_sdf_infer_workspace_bytes(int64_t batch, int64_t max_rows, int64_t num_points, int64_t margin, int32_t bins);
int hoisdf_sdf_infer_plan(const float* center, const float* cam_intr, const float* bbox, float sdf_scale, int64_t batch,
                          int32_t bins, int32_t* chunk_counts, int64_t* offsets, int64_t* host_offsets, void* stream);
int hoisdf_sdf_infer_fwd(const hoisdf_sdf_infer_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Token assembly -- upstream main/model.py:123-126 (sdf_activation) + :520-562:
 *   tokens[b, t, :] = cat[ xyz (3), posenc (30), fea (223) * sigmoid(sdf/beta)/beta ]
 * written batch-major (B, S, 256) at token offset t0 .. t0+P.  beta is a device scalar (already floored).
 * ------------------------------------------------------------------------------------------------- */
int hoisdf_tokens_fwd(const float* xyz, const float* posenc, const float* fea, int64_t ld_fea, const float* sdf,
                      const float* beta, int64_t batch, int64_t p, float* tokens, int64_t s_total, int64_t t0,
                      void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Multi-head attention core -- upstream nn.MultiheadAttention inside common/nets/transformer.py:294,378,383.
 *   q (B, Lq, ldq) , k/v (B, Lk, ldk) batch-major with heads as contiguous 64-wide column slices;
 *   out (B, Lq, ldo).  scores = q.k^T / sqrt(64); keys >= kv_valid are masked (memory_mask of
 *   common/utils/misc.py:42-47); optional dense bool mask (Lq, Lk) uint8, 1 = blocked (misc.py:11-31).
 *   Streaming (flash-style) softmax: the Lq x Lk score matrix is never materialised.
 * Three implementations behind one contract:
 *   workspace != NULL and no dense mask       : tcgen05 tensor-core kernel (BF16x3 split, fp32 accumulate in TMEM);
 *       `workspace` must hold hoisdf_attention_workspace_bytes(...) bytes (bf16 hi/lo copies of q, k, v^T);
 *   dense mask, or lq <= 32 w/o workspace     : small kernel (decoder: 17 queries);
 *   otherwise                                 : fp32 FMA streaming kernel.
 * ------------------------------------------------------------------------------------------------- */
int64_t hoisdf_attention_workspace_bytes(int64_t batch, int64_t heads, int64_t lq, int64_t lk);
int hoisdf_attention_fwd(const float* q, int64_t ldq, const float* k, const float* v, int64_t ldk, float* out,
                         int64_t ldo, int64_t batch, int64_t heads, int64_t lq, int64_t lk, int64_t kv_valid,
                         const uint8_t* mask, void* workspace, int64_t workspace_bytes, void* stream);

/* Training forward / backward of the same operator on the tensor cores (upstream main/train.py:131 `loss.backward()`
 * through common/nets/transformer.py:294, with nn.MultiheadAttention's dropout on the probabilities, cfg.dropout):
 *   out = dropout(softmax(q k^T / 8), p_drop) . v
 * The S x S probabilities never leave the SM in either direction.  The keep decisions are a counter hash of
 * (seed, (sample * heads + head) * lq + query, key) -- also what hoisdf_softmax_dropout_rows_fwd / _bwd (below) regenerate
 * on materialised (B, H, Lq, Lk) probabilities.  No dense mask; 0 <= p_drop < 1 (0: no dropout).
 *   _train_fwd : also writes lse (B*H*lq floats, may be NULL): log2-sum-exp of each scaled score row, from which the
 *                backward recomputes P.  min(lk, kv_valid) >= 128 (HOISDF_E_UNSUPPORTED otherwise: use the materialised
 *                form); workspace = hoisdf_attention_workspace_bytes(...) (required).
 *   _bwd       : dq (B*lq, ldg), dk, dv (B*lk, ldg) from q, k, v, out, dout (pitch ldo) and lse, same p_drop / seed as
 *                the forward; two tcgen05 kernels (dQ per query tile, dK / dV per key tile; no atomics), BF16x3 products.
 *                workspace = hoisdf_attention_bwd_workspace_bytes(...) (required). */
int hoisdf_attention_train_fwd(const float* q, int64_t ldq, const float* k, const float* v, int64_t ldk, float* out,
                               int64_t ldo, float* lse, int64_t batch, int64_t heads, int64_t lq, int64_t lk,
                               int64_t kv_valid, float p_drop, uint64_t seed, void* workspace, int64_t workspace_bytes,
                               void* stream);
int64_t hoisdf_attention_bwd_workspace_bytes(int64_t batch, int64_t heads, int64_t lq, int64_t lk);
int hoisdf_attention_bwd(const float* q, int64_t ldq, const float* k, const float* v, int64_t ldk, const float* out,
                         const float* dout, int64_t ldo, const float* lse, float* dq, float* dk, float* dv, int64_t ldg,
                         int64_t batch, int64_t heads, int64_t lq, int64_t lk, int64_t kv_valid, float p_drop,
                         uint64_t seed, void* workspace, int64_t workspace_bytes, void* stream);

/* The tensor-core path of hoisdf_attention_fwd (workspace required, no dense mask) writing its result in split-half
 * format (two fp16 planes, pitch ldo halfs, multiple of 8) -- what the FP16x3 out-projection GEMM reads. */
int hoisdf_attention_split_fwd(const float* q, int64_t ldq, const float* k, const float* v, int64_t ldk,
                               uint16_t* out_hi, uint16_t* out_lo, int64_t ldo, int64_t batch, int64_t heads, int64_t lq,
                               int64_t lk, int64_t kv_valid, void* workspace, int64_t workspace_bytes, void* stream);

/* y = LayerNorm(x (+ res)) * gamma + beta, eps 1e-5, rows of 256 (transformer.py:296-301); optional second
 * output y2 = LayerNorm(y) with (gamma2, beta2) -- the shared `inter_norm` of transformer.py:196-197. */
int hoisdf_add_layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* y,
                             const float* gamma2, const float* beta2, float* y2, int64_t rows, int64_t d,
                             void* stream);
/* The same, additionally writing y (and y2) in split-half format (pitch in halfs, multiple of 4; NULL = skip) so the
 * following FP16x3 Linear reads it without a separate hoisdf_split_rows pass. */
int hoisdf_add_layernorm_split_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* y,
                                   const float* gamma2, const float* beta2, float* y2, int64_t rows, int64_t d,
                                   uint16_t* yh_hi, uint16_t* yh_lo, int64_t ldyh, uint16_t* y2h_hi, uint16_t* y2h_lo,
                                   int64_t ldy2h, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * A whole transformer ENCODER stack behind one entry point (csrc/transformer.cu) -- upstream
 * common/nets/transformer.py:175-202 (TransformerEncoder.forward) over :279-302 (forward_post), pos_embed == 0
 * (main/model.py:541-543), d_model 256, heads of 64, ReLU feed-forward of width d_ff, eval mode.
 *   x (B*S, 256) fp32 batch-major tokens -> out (B*S, 256) = last layer's output (+ its split-half copy when out_hi is
 *   given: what the decoder's key / value projection reads) and, when `inter` is given, inter (L, B*S, 256) =
 *   inter_norm(out_l) for every layer (+ split-half copies: what the vote heads read).
 *   Weights: hoisdf_pack_h3 planes (a, b, c; pitch ld) + fp32 bias per Linear; `scale` as hoisdf_linear_h3_args.w_scale.
 *   seq <= 32 returns HOISDF_E_UNSUPPORTED (short sequences use the SIMT attention kernel through hoisdf_attention_fwd).
 * ------------------------------------------------------------------------------------------------- */
typedef struct { const uint16_t* a; const uint16_t* b; const uint16_t* c; int64_t ld; const float* bias; float scale; } hoisdf_h3_linear;
typedef struct {
  hoisdf_h3_linear qkv;   /* in_proj (768, 256): rows [q; k; v] */
  hoisdf_h3_linear out;   /* out_proj (256, 256) */
  hoisdf_h3_linear lin1;  /* linear1 (d_ff, 256) */
  hoisdf_h3_linear lin2;  /* linear2 (256, d_ff) */
  const float* norm1_g; const float* norm1_b; const float* norm2_g; const float* norm2_b;
} hoisdf_encoder_layer;
typedef struct {
  const hoisdf_encoder_layer* layers; int32_t num_layers; int32_t heads; int64_t d_ff;
  const float* inter_g; const float* inter_b;
  int64_t batch; int64_t seq;
  const float* x;
  float* out; uint16_t* out_hi; uint16_t* out_lo; int64_t ld_out;
  float* inter; uint16_t* inter_hi; uint16_t* inter_lo; int64_t ld_inter;
  void* workspace; int64_t workspace_bytes;
} hoisdf_encoder_args;
int64_t hoisdf_encoder_workspace_bytes(int64_t batch, int64_t seq, int64_t d_ff, int32_t heads);
int hoisdf_encoder_fwd(const hoisdf_encoder_args* args, void* stream);

/* The transformer DECODER stack over `queries` learned queries (csrc/transformer.cu) -- upstream
 * common/nets/transformer.py:214-252 (TransformerDecoder.forward) over :366-395 (forward_post) with tgt = 0,
 * query_pos = the query embedding expanded over the batch ((B * queries, 256) fp32), pos = 0:
 *   hs (L, B * queries, 256) = norm(out_l) for every layer.
 * memory: the encoder's last output in split-half format (hoisdf_encoder_fwd's out_hi / out_lo; row pitch ld_memory),
 * B * seq rows; tgt_mask uint8 (queries, queries), non-zero = blocked (may be NULL); keys >= kv_valid of the memory are
 * blocked for every query.  Weights as in hoisdf_encoder_layer: in_proj row windows [q; k] / [v] of the self-attention,
 * [q] / [k; v] of the cross-attention. */
typedef struct {
  hoisdf_h3_linear sa_qk; hoisdf_h3_linear sa_v; hoisdf_h3_linear sa_out;
  hoisdf_h3_linear ca_q; hoisdf_h3_linear ca_kv; hoisdf_h3_linear ca_out;
  hoisdf_h3_linear lin1; hoisdf_h3_linear lin2;
  const float* norm1_g; const float* norm1_b; const float* norm2_g; const float* norm2_b; const float* norm3_g; const float* norm3_b;
} hoisdf_decoder_layer;
typedef struct {
  const hoisdf_decoder_layer* layers; int32_t num_layers; int32_t heads; int64_t d_ff;
  const float* norm_g; const float* norm_b;
  int64_t batch; int64_t queries; int64_t seq; int64_t kv_valid;
  const float* query_pos; const uint8_t* tgt_mask;
  const uint16_t* memory_hi; const uint16_t* memory_lo; int64_t ld_memory;
  float* hs;
  void* workspace; int64_t workspace_bytes;
} hoisdf_decoder_args;
int64_t hoisdf_decoder_workspace_bytes(int64_t batch, int64_t queries, int64_t seq, int64_t d_ff, int32_t heads);
int hoisdf_decoder_fwd(const hoisdf_decoder_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Joint voting -- upstream common/nets/loss.py:31-36,54-57 (the part of JointvoteLoss that produces
 * `hand_joints`): softmax over points of the class logits, weighted sum of point + offset.
 *   points (B,P,3), off (L,B,P,60), cls (L,B,P,20) batch-major -> joints (L,B,20,3)
 * ------------------------------------------------------------------------------------------------- */
int hoisdf_vote_joints_fwd(const float* points, const float* off, const float* cls, int64_t layers,
                           int64_t batch, int64_t p, float* joints, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * ManoHead + ManoLayer -- upstream common/nets/mano_head.py:185-256 and
 * manopth/manopth/manolayer.py:111-276 (axis-angle, flat hand mean, centre on joint 0, right hand):
 *   pose6d (N,16,6), betas (N,10) -> verts (N,778,3), joints (N,21,3) in metres.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  const float* shapedirs;   /* (778,3,10) */
  const float* posedirs;    /* (778,3,135) */
  const float* v_template;  /* (778,3) */
  const float* j_regressor; /* (16,778) */
  const float* weights;     /* (778,16) */
  const float* hands_mean;  /* (45) */
} hoisdf_mano_model;

int hoisdf_mano_fwd(const hoisdf_mano_model* model, const float* pose6d, const float* betas, int64_t n,
                    float* verts, float* joints, void* stream);
/* Same MANO forward from axis-angle parameters (N,48) -- the ground-truth branch of ManoHead.forward
 * (upstream common/nets/mano_head.py:258-276: training and the dexycb evaluation). */
int hoisdf_mano_aa_fwd(const hoisdf_mano_model* model, const float* pose_aa, const float* betas, int64_t n,
                       float* verts, float* joints, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Test-time metrics -- upstream common/metrics.py:62-185 (eval_batched_obj_direct with
 * compute_obj_metrics_dexycb / compute_obj_metrics_ho3d; called from main/test.py:131-135) and :188-248
 * (rigid_transform_3D / rigid_align / eval_hand_joint; main/test.py:186-193, main/train.py:241): the
 * consumers of the path's obj_rot_out / obj_trans_out / mano_joints_out.
 *
 * hoisdf_obj_metrics_fwd: per sample b
 *     rot = mean_p rot_pred[b,p,:], trans = mean_p trans_pred[b,p,:]            (metrics.py:115-116)
 *     pred mesh = template . R(rot)^T + trans, target mesh = template . R(rot_gt[b])^T + trans_gt[b]
 *                 with R = manopth batch_rodrigues (rodrigues_layer.py:15-56)     (metrics.py:152-167)
 *     adds[b] = mean_i min_j |target_j - pred_i|      (ADD-S, :64-67)   mme[b] = mean_i |target_i - pred_i| (:106)
 *     mce[b]  = mean over the 8 axis-aligned bounding-box corners of |corner_pred - corner_target| (:70-94)
 *     oce[b]  = |trans - trans_gt[b]|                                              (:172,179)
 *   templates (T, N, 3); obj_ids (B) int64 selects the template of each sample (NULL: T >= B, sample b uses
 *   template b); rot_pred / trans_pred (B, votes, 3) axis-angle / metres; rot_gt / trans_gt (B, 3).
 *   adds / mme / mce / oce: (B) each, any may be NULL.  The (N x N) distance tensor upstream materialises never
 *   exists.  workspace: hoisdf_obj_metrics_workspace_bytes(batch, n_verts) bytes (per-CTA partials).
 * hoisdf_mesh_metrics_fwd: the same three mesh metrics for given meshes (B, N, 3) -- compute_obj_metrics_dexycb /
 *   compute_obj_metrics_ho3d called directly.
 * hoisdf_hand_joint_metrics_fwd: per sample, mje = mean_i |pred_i - gt_i| and pamje = the same after the similarity
 *   (scale, rotation, translation) alignment of rigid_transform_3D; `aligned` (B, n_points, 3), optional, receives
 *   rigid_align(pred, gt).  pred / gt (B, n_points, 3); n_points = 21 joints or 778 vertices.
 * ------------------------------------------------------------------------------------------------- */
int64_t hoisdf_obj_metrics_workspace_bytes(int64_t batch, int64_t n_verts);
int hoisdf_obj_metrics_fwd(const float* templates, const int64_t* obj_ids, int64_t n_templates, int64_t n_verts,
                           const float* rot_pred, const float* trans_pred, int64_t votes, const float* rot_gt,
                           const float* trans_gt, int64_t batch, float* adds, float* mme, float* mce, float* oce,
                           void* workspace, int64_t workspace_bytes, void* stream);
int hoisdf_mesh_metrics_fwd(const float* pred_meshes, const float* target_meshes, int64_t batch, int64_t n_verts,
                            float* adds, float* mme, float* mce, void* workspace, int64_t workspace_bytes,
                            void* stream);
int hoisdf_hand_joint_metrics_fwd(const float* pred, const float* gt, int64_t batch, int64_t n_points, float* mje,
                                  float* pamje, float* aligned, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Backward kernels of the training step (upstream main/train.py:104-140: `loss.backward()` through main/model.py:357-665
 * and `optimizer.step()`, common/base.py:64-70).  fp32 SIMT arithmetic (csrc/backward.cu); the large Linear gradients run on
 * hoisdf_linear_h3_fwd with operands from hoisdf_linear_bwd_prep (csrc/train_prep.cu).  Called by hoisdf_b200/autograd.py
 * (Model.forward(mode="train"), hoisdf_b200/train.py); checked against PyTorch autograd on the CPU thread emulator and on the
 * B200.
 *   hoisdf_gemm_f32: C (m, n; pitch ldc) = op(A) . op(B) (+ C when accumulate): trans_a: A is stored (k, m), else (m, k);
 *     trans_b: B is stored (n, k), else (k, n).  A Linear Y = X . W^T has dX = dZ . W (no transposes), dW = dZ^T . X
 *     (trans_a) and Y itself (trans_b).
 *   hoisdf_act_bias_bwd: dZ = dY * [Y > 0] in place when act == HOISDF_ACT_RELU (Y = the forward's output), and
 *     db[n] = sum_m dZ[m, n] (db may be NULL; deterministic fixed-order sums).
 *   hoisdf_weight_norm_bwd: W = g * v / |v| per row (nn.utils.weight_norm, dim 0; upstream sdf_net.py:57-62):
 *     dg (rows), dv (rows, cols) from dW (rows, cols; pitch lddw).
 *   hoisdf_gather_bwd: CONCAT-mode bilinear gather backward: grad->map[l] (NHWC, same geometry as the forward's
 *     pyramid) += scatter of dout (rows, ld_dout) with the forward's tap weights (atomic adds; the sampling grid is
 *     detached upstream, main/model.py:158,199, so the points receive no gradient).
 *   hoisdf_sdf_loss_bwd: dz_i = d/dz_i [ scale * mean_i | clamp(tanh(z_i), +-clamp) - clamp(gt_i, +-clamp) | ]
 *     (upstream SepSDFLoss, common/nets/loss.py:64-78, with the clamps of main/model.py:241,388-395).
 *   hoisdf_layernorm_bwd: y = LayerNorm(h) * gamma + beta over rows of d = 256 (eps 1e-5; upstream transformer.py:296-301,
 *     h = the residual sum the forward normalised): dh (rows, d), and -- when dgamma / dbeta are given -- the parameter
 *     gradients (deterministic column sums; `stats` = workspace of 2 * rows floats, required with them).
 *   hoisdf_softmax_rows_fwd / _bwd: p = softmax(s[:, :valid]) per row (0 beyond `valid`, and 0 where the optional bool
 *     attn_mask (mask_rows, cols) -- row r uses mask row r % mask_rows, non-zero = blocked -- blocks a column);
 *     ds = p * (dp - sum_j dp_j p_j) (ds may alias dp).  Together with hoisdf_gemm_f32 per (sample, head) these are the
 *     backward of nn.MultiheadAttention's core: dV = P^T dO, dP = dO V^T, dS = softmax'(dP), dQ = dS K / 8, dK = dS^T Q / 8.
 *   hoisdf_tokens_bwd: backward of hoisdf_tokens_fwd (upstream main/model.py:123-126,520-531) for one token group:
 *     d_tokens (B, s_total, 256) -> d_fea (B*P, ld_dfea >= 223) = d_tok[.., 33:] * sigmoid(sdf / beta) / beta, d_sdf (B*P,
 *     may be NULL: the selected points' SDF is detached upstream), d_beta (1 float; the learnable hand / obj_sigmoid_beta);
 *     workspace: hoisdf_tokens_bwd_workspace_bytes(batch, p) (per-block partials of d_beta, folded in a fixed order).
 *   hoisdf_vote_loss_bwd: JointvoteLoss (upstream common/nets/loss.py:22-61), batch-major like hoisdf_vote_joints_fwd:
 *     points (B,P,3) [m], off (L,B,P,60), cls (L,B,P,20), joint_gt (B,20,3) [mm]; d_off / d_cls = gradients of
 *     g_joint_3d * loss_joint_3d + g_cls * loss_joint_cls + g_all_joint_3d * loss_all_joint_3d (the points carry no
 *     gradient); npos_ws: one float of workspace (the number of positive point / joint pairs).
 *   hoisdf_adamw_step: one torch.optim.AdamW update (upstream common/base.py:68) of a flat buffer of n parameters,
 *     `step` = 1-based update count (bias correction); decoupled weight decay, PyTorch's operation order.
 * ------------------------------------------------------------------------------------------------- */
int hoisdf_gemm_f32(const float* a, int64_t lda, int32_t trans_a, const float* b, int64_t ldb, int32_t trans_b, float* c,
                    int64_t ldc, int64_t m, int64_t n, int64_t k, int32_t accumulate, void* stream);
/* batched form: batch_outer x batch_inner matrices (e.g. sample x head), per-operand strides in floats for both levels;
 * C = alpha * op(A) . op(B) (+ C).  The attention core's backward on the (B, S, 3 d) projections: head h of sample b
 * lives at b * S * ld + h * 64. */
int hoisdf_gemm_f32_batched(const float* a, int64_t lda, int32_t trans_a, int64_t a_outer, int64_t a_inner, const float* b,
                            int64_t ldb, int32_t trans_b, int64_t b_outer, int64_t b_inner, float* c, int64_t ldc,
                            int64_t c_outer, int64_t c_inner, int64_t m, int64_t n, int64_t k, float alpha, int32_t accumulate,
                            int64_t batch_outer, int64_t batch_inner, void* stream);
/* ---------------------------------------------------------------------------------------------------
 * Data feed, image warp (upstream data/ho3d.py:399-427 `data_crop`, :351-381 in `data_aug`,
 * data/dexycb.py likewise; data/dataset_util.py:44-51 `transform_img`): PIL `Image.transform((size, size), AFFINE, coef)` with
 * its default NEAREST resampling on a batch of 8-bit frames in device memory, bit-exact with Pillow 12.2.0 (both of its 8-bit
 * code paths: the scale-only table walk of the evaluation crop -- also what `Image.resize(..., NEAREST)` of the segmentation
 * masks runs -- and the 16.16 fixed-point affine of the rotation augmentation).
 *   src: (batch, src_h, src_w, channels) bytes, channels = 3 (RGB) or 1 (mode "L" masks), row pitch src_pitch bytes, frame
 *        pitch src_stride bytes;
 *   coef: 6 doubles per sample, PIL's `data` = the first two rows of the INVERSE affine (output pixel -> source pixel);
 *   mirror: NULL or one int32 per sample, != 0: warp the left-right mirrored frame (data/dexycb.py:427-430,479-481: left hands);
 *   out_f32 (batch, channels, size, size) = pixel / divisor in fp32 (divisor 255: upstream's ToTensor(...) / 255.0,
 *        ho3d.py:550,624; divisor 1: the masks' astype(float32), :551-552) and / or
 *   out_u8 (batch, size, size, channels) = the PIL image; pixels that map outside the source are 0;
 *   tables: 2 * size int32 per sample of scratch.
 * Transforms that Pillow evaluates with its floating-point loop (an output corner mapping to |coordinate| >= 32768) are not
 * restated: the caller checks (hoisdf_b200/feed.py raises).
 * ------------------------------------------------------------------------------------------------- */
int hoisdf_image_crop_fwd(const uint8_t* src, int64_t batch, int64_t src_h, int64_t src_w, int64_t channels, int64_t src_pitch,
                          int64_t src_stride, const double* coef, const int32_t* mirror, int64_t size, float divisor, float* out_f32,
                          uint8_t* out_u8, int32_t* tables, void* stream);

/* Data feed, SDF point sets (upstream data/ho3d.py:484-486 `sdf_data[all_idx]`, :333 the augmentation's
 * rotation, :524-548 normalisation, :561-579 the `inputs` / `targets` entries; data/dexycb.py:515-548 incl. the mirror flip):
 * from the packed rows [x, y, z, sdf_hand, sdf_obj, label] (tool/pre_process_sdf.py:140-147) of `batch` frames in device
 * memory -- rows (total, 6) float32, frame b = rows[row_offsets[b] .. row_offsets[b + 1]) -- and the caller's draws
 * index (batch, n_sel) int64 (frame-local, upstream's `all_idx`: [hand n_hand | object n_obj | hand_pre n_hand | obj_pre n_obj],
 * the last two only when n_sel = 2 * (n_hand + n_obj)) to
 *   hand_points / hand_pre (batch, n_hand, 3) = (R xyz - hand_root) * hand_scale,  hand_sdf (batch, n_hand) = row[3] * hand_scale,
 *   obj_points / obj_pre (batch, n_obj, 3) = (R xyz - obj_centre) * obj_scale,     obj_sdf (batch, n_obj) = row[4] * obj_scale
 * in numpy's float32 arithmetic (xyz . rot^T accumulated x, y, z with fused multiply-adds as sgemm does).
 *   rot (batch, 3, 3) or NULL (evaluation: no rotation); flip (batch) int32 or NULL (dexycb mirror: x -> -x before the rotation);
 *   status: one int32 the kernel sets to 1 when an index lies outside its frame's rows (the caller clears and reads it).
 * ------------------------------------------------------------------------------------------------- */
int hoisdf_sdf_rows_fwd(const float* rows, const int64_t* row_offsets, const int64_t* index, int64_t batch, int64_t n_sel,
                        int64_t n_hand, int64_t n_obj, const float* rot, const int32_t* flip, const float* hand_root,
                        const float* obj_centre, float hand_scale, float obj_scale, float* hand_points, float* obj_points,
                        float* hand_pre, float* obj_pre, float* hand_sdf, float* obj_sdf, int32_t* status, void* stream);

/* Data feed, photometric augmentation of the training sample (upstream data/ho3d.py:355-364, data/dexycb.py:310-321):
 * `img.filter(ImageFilter.GaussianBlur(radius))` on a batch of 8-bit images in device memory, bit-exact with Pillow 12.2.0
 * (libImaging/BoxBlur.c: three box blurs per axis with fixed-point weights, bytes rounded after every pass).
 *   hoisdf_gaussian_blur_params (HOST function, no GPU work): Pillow's `_gaussian_blur_radius` and box weights for one radius,
 *     out[3] = {n, ww, fw}; the caller uploads one triple per sample;
 *   hoisdf_gaussian_blur_u8: src / dst / scratch (batch, h, w, channels) bytes, packed; params (batch, 3) uint32 on the device;
 *     passes = 3 for GaussianBlur.  2 * w * channels <= 48 KB and 8 * h <= 48 KB (HOISDF_E_SHAPE otherwise).  src may equal dst.
 * ------------------------------------------------------------------------------------------------- */
int hoisdf_gaussian_blur_params(float radius, int32_t passes, uint32_t* out);
int hoisdf_gaussian_blur_u8(const uint8_t* src, uint8_t* dst, uint8_t* scratch, int64_t batch, int64_t h, int64_t w,
                            int64_t channels, const uint32_t* params, int32_t passes, void* stream);

/* `dataset_util.color_jitter` (upstream data/dataset_util.py:144-201, called from ho3d.py:358-364): up to four torchvision
 * adjustments of a PIL image in a shuffled order, on a batch of 8-bit RGB images in device memory, bit-exact with torchvision's
 * PIL branch over Pillow 12.2.0 (ImageEnhance = ImagingBlend with black / the grey version / the mean grey; adjust_hue =
 * Pillow's RGB -> HSV, H shifted with byte wrap-around, HSV -> RGB).
 *   src / dst (batch, h, w, 3) bytes, packed (src may equal dst); h * w <= 2^24;
 *   ops (batch, 4) int32: the adjustment of each of the four steps, 0 none, 1 brightness, 2 saturation, 3 hue, 4 contrast;
 *   factors (batch, 4) float: that step's factor as torchvision receives it (hue: the byte added to H,
 *     np.int32(hue_factor * 255).astype(np.uint8), as a float 0..255);
 *   sums: 4 * batch uint64 of scratch (the grey sums behind the contrast adjustment's mean; cleared by the call).
 * ------------------------------------------------------------------------------------------------- */
int hoisdf_color_jitter_u8(const uint8_t* src, uint8_t* dst, int64_t batch, int64_t h, int64_t w, const int32_t* ops,
                           const float* factors, uint64_t* sums, void* stream);

/* The whole training image in ONE launch (upstream data/ho3d.py:351-364,550: transform_img + crop, GaussianBlur, color_jitter,
 * ToTensor / 255): one CTA per frame, the warped res x res x 3 image resident in shared memory from the gather to the final
 * store (hoisdf_train_image_smem_bytes(res): 193 KB at res = 256, of the SM's 227 KB) -- the same bytes as
 * hoisdf_image_crop_fwd -> hoisdf_gaussian_blur_u8 -> hoisdf_color_jitter_u8 -> the float conversion, in 1 launch instead of 13.
 *   src / coef / mirror as hoisdf_image_crop_fwd (RGB); blur (batch, 3) uint32 = hoisdf_gaussian_blur_params triples, box radius
 *   n = 0 only (Gaussian radius < ~1.4: every radius upstream draws; the caller checks, hoisdf_b200/feed.py falls back);
 *   ops / factors as hoisdf_color_jitter_u8; out_f32 (batch, 3, res, res) and / or out_u8 (batch, res, res, 3).
 *   HOISDF_E_UNSUPPORTED when the image does not fit one SM's shared memory (res > 256).
 * ------------------------------------------------------------------------------------------------- */
int64_t hoisdf_train_image_smem_bytes(int64_t res);
int hoisdf_train_image_fwd(const uint8_t* src, int64_t batch, int64_t src_h, int64_t src_w, int64_t src_pitch, int64_t src_stride,
                           const double* coef, const int32_t* mirror, const uint32_t* blur, const int32_t* ops,
                           const float* factors, int64_t res, float* out_f32, uint8_t* out_u8, void* stream);

/* The segmentation masks of a sample in ONE launch (upstream data/ho3d.py:366-381,551-552; data/dexycb.py:323-336,389-402):
 * `transform_img(mask)` to (res, res), `.resize((out_res, out_res), Image.NEAREST)`, `.astype(np.float32)` -- one CTA per mask,
 * Pillow's tables built in shared memory, one byte gathered from the frame per output pixel; the same values as two
 * hoisdf_image_crop_fwd calls (channels = 1, divisor 1) per mask.
 *   src (batch, src_h, src_w) bytes (mode "L"), row pitch src_pitch, frame pitch src_stride; coef / mirror as
 *   hoisdf_image_crop_fwd; out (batch, out_res, out_res) float32; out_res <= res <= 4096.
 * ------------------------------------------------------------------------------------------------- */
int hoisdf_mask_crop_fwd(const uint8_t* src, int64_t batch, int64_t src_h, int64_t src_w, int64_t src_pitch, int64_t src_stride,
                         const double* coef, const int32_t* mirror, int64_t res, int64_t out_res, float* out, void* stream);

/* One Linear of the training step per call (what hoisdf_b200/autograd.py:LinearFn runs; upstream main/train.py:108-131 through
 * every nn.Linear of the hot path), fp32 in / fp32 out on the FP16x3 tensor-core GEMM, caller-owned workspace of
 * hoisdf_linear_train_workspace_bytes(m, n, k) bytes (16-byte aligned; HOISDF_E_WORKSPACE when too small):
 *   _fwd: y (m, n) = act(x (m, k) . w (n, k)^T + bias)              = hoisdf_split_rows + hoisdf_pack_h3 + hoisdf_linear_h3_fwd
 *   _bwd: dZ = dy * [y > 0] (act == RELU), db (n) = column sums of dZ (may be NULL),
 *         dx (m, k) = dZ . w (may be NULL), dwt (k, n) = x^T . dZ = dW TRANSPOSED (may be NULL; split-K, see
 *         hoisdf_linear_h3_args.split_k)                           = hoisdf_absmax + hoisdf_linear_bwd_prep + two GEMMs
 * Pitches ldy / lddx / lddwt that are multiples of 4 floats (and 16-byte aligned bases) take the TMA-store epilogue. */
int64_t hoisdf_linear_train_workspace_bytes(int64_t m, int64_t n, int64_t k);
int hoisdf_linear_train_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, int64_t m, int64_t n,
                            int64_t k, int32_t act, float* y, int64_t ldy, void* workspace, int64_t workspace_bytes,
                            void* stream);
int hoisdf_linear_train_bwd(const float* dy, int64_t lddy, const float* y, int64_t ldy, const float* x, int64_t ldx,
                            const float* w, int64_t ldw, int64_t m, int64_t n, int64_t k, int32_t act, float* dx, int64_t lddx,
                            float* dwt, int64_t lddwt, float* db, void* workspace, int64_t workspace_bytes, void* stream);
/* Linear layers with n <= 16 output features over many rows (the SDF value / class / offset heads: upstream
 * common/nets/sdf_net.py:53-64, main/model.py:82-91), fp32 FMA, one streaming pass over x:
 *   hoisdf_thin_linear_fwd: y (m, n; pitch ldy) = act(x (m, k) . w (n, k)^T + bias) (bias may be NULL);
 *   hoisdf_thin_linear_dw : dw (n, k; pitch lddw) = dz (m, n)^T . x (m, k) (+ dw when accumulate) -- the weight gradient,
 *                           without transposed copies of x or dz; partial sums meet in atomicAdd.
 * n > 16: HOISDF_E_UNSUPPORTED (use the GEMMs). */
int hoisdf_thin_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, int64_t m, int64_t k,
                           int64_t n, int32_t act, float* y, int64_t ldy, void* stream);
int hoisdf_thin_linear_dw(const float* x, int64_t ldx, const float* dz, int64_t lddz, int64_t m, int64_t k, int64_t n,
                          float* dw, int64_t lddw, int32_t accumulate, void* stream);
int hoisdf_act_bias_bwd(float* dy, int64_t lddy, const float* y, int64_t ldy, int64_t m, int64_t n, int32_t act, float* db,
                        int32_t accumulate, void* stream);
int hoisdf_weight_norm_bwd(const float* g, const float* v, const float* dw, int64_t lddw, int64_t rows, int64_t cols,
                           float* dg, float* dv, int32_t accumulate, void* stream);
int hoisdf_gather_bwd(const hoisdf_pyramid* grad, const float* uv, int64_t rows, const int64_t* row_offsets, int64_t batch,
                      int64_t rows_per_sample, const float* dout, int64_t ld_dout, void* stream);
int hoisdf_sdf_loss_bwd(const float* z, const float* sdf_gt, int64_t n, float clamp, float scale, float* dz, void* stream);
int hoisdf_layernorm_bwd(const float* h, const float* gamma, const float* dy, int64_t rows, int64_t d, float* dh,
                         float* dgamma, float* dbeta, float* stats, int32_t accumulate, void* stream);
int hoisdf_softmax_rows_fwd(const float* s, int64_t lds, int64_t rows, int64_t cols, int64_t valid, const uint8_t* mask,
                            int64_t mask_rows, float* p, int64_t ldp, void* stream);
int hoisdf_softmax_rows_bwd(const float* p, int64_t ldp, const float* dp, int64_t lddp, int64_t rows, int64_t cols, float* ds,
                            int64_t ldds, void* stream);
/* Operand preparation for the tensor-core backward of a Linear (csrc/train_prep.cu): dX = dZ . W and dW^T = X^T . dZ run on
 * hoisdf_linear_h3_fwd with
 *   hoisdf_absmax: out[0] = max |x| over a (rows, cols; pitch ld) matrix (device scalar);
 *   hoisdf_linear_bwd_prep: dZ = dY * [Y > 0] (act == HOISDF_ACT_RELU), s = 2^(ceil(log2(amax)) - 3) computed on the device
 *     (written to scale_out), then in ONE pass dZ / s in split-half format (dz_hi / dz_lo, pitch ld_dz: the x operand of the dX
 *     GEMM), (dZ / s)^T as the three hoisdf_pack_h3 planes (dzt_*, (n, ld_dzt >= m): the w operand of the dW GEMM) and
 *     db[c] = sum_r dZ[r, c] (may be NULL; atomic adds);
 *   hoisdf_split_rows_t: X (m, k) fp32 -> X^T in split-half format, planes (k, ldh >= m). */
int hoisdf_absmax(const float* x, int64_t rows, int64_t cols, int64_t ld, float* out, void* stream);
int hoisdf_linear_bwd_prep(const float* dy, int64_t lddy, const float* y, int64_t ldy, int64_t m, int64_t n, int32_t act,
                           const float* amax, uint16_t* dz_hi, uint16_t* dz_lo, int64_t ld_dz, uint16_t* dzt_a, uint16_t* dzt_b,
                           uint16_t* dzt_c, int64_t ld_dzt, float* db, float* scale_out, void* stream);
int hoisdf_split_rows_t(const float* x, int64_t m, int64_t k, int64_t ldx, uint16_t* hi, uint16_t* lo, int64_t ldh, void* stream);
/* Row softmax with nn.MultiheadAttention's dropout on the probabilities (p_drop in [0, 1)) and its backward, without a stored
 * mask: keep(r, c) is a counter-based hash of (seed, row r, column c) -- the decisions of hoisdf_attention_train_fwd / _bwd for
 * row = (sample * heads + head) * lq + query.  Forward: p (may be NULL) as hoisdf_softmax_rows_fwd,
 * pd = keep ? p / (1 - p_drop) : 0 (s may alias p or pd).  Backward: ds = p * (g - sum_j g_j p_j) with
 * g = keep ? dpd / (1 - p_drop) : 0 (ds may alias dpd). */
int hoisdf_softmax_dropout_rows_fwd(const float* s, int64_t lds, int64_t rows, int64_t cols, int64_t valid, const uint8_t* mask,
                                    int64_t mask_rows, float* p, int64_t ldp, float* pd, int64_t ldpd, float p_drop, uint64_t seed,
                                    void* stream);
int hoisdf_softmax_dropout_rows_bwd(const float* p, int64_t ldp, const float* dpd, int64_t lddp, int64_t rows, int64_t cols,
                                    float* ds, int64_t ldds, float p_drop, uint64_t seed, void* stream);
int64_t hoisdf_tokens_bwd_workspace_bytes(int64_t batch, int64_t p);
int hoisdf_tokens_bwd(const float* d_tokens, int64_t s_total, int64_t t0, const float* fea, int64_t ld_fea, const float* sdf,
                      const float* beta, int64_t batch, int64_t p, float* d_fea, int64_t ld_dfea, float* d_sdf, float* d_beta,
                      int32_t accumulate_beta, void* workspace, int64_t workspace_bytes, void* stream);
int hoisdf_vote_loss_bwd(const float* points, const float* off, const float* cls, const float* joint_gt, int64_t layers,
                         int64_t batch, int64_t p, float cls_dist, float g_joint_3d, float g_cls, float g_all_joint_3d,
                         float* d_off, float* d_cls, float* npos_ws, void* stream);
int hoisdf_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                      float beta2, float eps, float weight_decay, int64_t step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HOISDF_B200_H */
