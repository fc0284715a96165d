// Backward kernels of the training step (upstream main/train.py:104-140 back-propagates the weighted
// loss sum through main/model.py:357-665).  fp32 SIMT arithmetic: the element-wise / reduction / scatter pieces of the
// backward and the small or batched products that are not tensor-core shapes; the large Linear gradients (dX, dW) run on the
// FP16x3 tcgen05 GEMM with operands prepared by csrc/train_prep.cu.  Called through hoisdf_b200/autograd.py; every kernel is
// checked against PyTorch autograd on the CPU thread emulator (tests/test_kernel_emulation.py) and on the B200
// (tests/test_gpu_zz_backward.py, tests/test_gpu_zy_train.py).
//   hoisdf_gemm_f32          C (M,N) = op(A) . op(B) (+ C): the three contractions of a Linear's backward
//                            (dX = dZ . W, dW = dZ^T . X) and its forward (Y = X . W^T) on one tiled fp32 FMA kernel
//   hoisdf_act_bias_bwd      dZ = dY * relu'(Y) in place, db = column sums of dZ
//   hoisdf_weight_norm_bwd   gradients of nn.utils.weight_norm (dim 0): W = g * v / |v|  ->  dg, dv
//   hoisdf_gather_bwd        bilinear gather backward: scatter-add of row gradients into the NHWC pyramid gradient
//                            (the sampling grid is detached upstream, main/model.py:158,199: no gradient to the points)
//   hoisdf_sdf_loss_bwd      clamp + L1 mean of SepSDFLoss (common/nets/loss.py:64-78) and tanh': dLoss / d(pre-tanh)
//   hoisdf_layernorm_bwd     nn.LayerNorm(256) backward: dh, dgamma, dbeta (transformer.py:296-301)
//   hoisdf_softmax_rows_fwd / _bwd   row softmax and its backward: with hoisdf_gemm_f32 per head, the attention core's
//                            backward (dV = P^T dO, dP = dO V^T, dS = softmax', dQ = dS K, dK = dS^T Q)
//   hoisdf_adamw_step        torch.optim.AdamW over a flat parameter buffer (upstream common/base.py:68)
//   hoisdf_tokens_bwd        token assembly with the SDF activation sigmoid(sdf / beta) / beta (model.py:123-126,520-531):
//                            gradients of the point features, the SDF values and beta
//   hoisdf_vote_loss_bwd     JointvoteLoss (common/nets/loss.py:22-61): gradients of the three losses w.r.t. the vote
//                            offsets and the class logits
#include <cmath>

#include "common.cuh"

namespace hoisdf {
namespace {

constexpr int GB = 64;      // C tile edge
constexpr int GK = 16;      // K step

struct GemmBatch {          // strides (in floats) of the two batch levels of hoisdf_gemm_f32_batched; all 0 / inner 1 = one matrix
  int64_t a_outer, a_inner, b_outer, b_inner, c_outer, c_inner;
  int inner;
  float alpha;
};

// C[m, n] = sum_k a(m, k) * b(k, n); TA: A is stored (K, M) (a(m,k) = A[k*lda + m]), else (M, K);
//                                    TB: B is stored (N, K) (b(k,n) = B[n*ldb + k]), else (K, N).  256 threads, 4 x 4 each.
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb, float* __restrict__ Cm,
                int64_t ldc, int64_t M, int64_t N, int64_t K, int accumulate, GemmBatch gb) {
  __shared__ float As[GK][GB + 4];
  __shared__ float Bs[GK][GB + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  {   // batched form: blockIdx.z = outer * inner_count + inner (e.g. sample x head), per-operand strides in floats
    const int64_t zo = blockIdx.z / gb.inner, zi = blockIdx.z % gb.inner;
    A += zo * gb.a_outer + zi * gb.a_inner;
    B += zo * gb.b_outer + zi * gb.b_inner;
    Cm += zo * gb.c_outer + zi * gb.c_inner;
  }
  const int64_t m0 = static_cast<int64_t>(blockIdx.y) * GB, n0 = static_cast<int64_t>(blockIdx.x) * GB;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t k0 = 0; k0 < K; k0 += GK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int f = tid + 256 * e;                    // 1024 elements per operand tile
      // pick the fast-running index along the operand's contiguous dimension
      const int ka = TA ? f / GB : f % GK, ma = TA ? f % GB : f / GK;
      const int64_t m = m0 + ma, k = k0 + ka;
      As[ka][ma] = (m < M && k < K) ? (TA ? A[k * lda + m] : A[m * lda + k]) : 0.f;
      const int kb = TB ? f % GK : f / GB, nb = TB ? f / GK : f % GB;
      const int64_t n = n0 + nb, kk = k0 + kb;
      Bs[kb][nb] = (n < N && kk < K) ? (TB ? B[n * ldb + kk] : B[kk * ldb + n]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n < N) Cm[m * ldc + n] = accumulate ? Cm[m * ldc + n] + gb.alpha * acc[i][j] : gb.alpha * acc[i][j];
    }
  }
}

// dZ = dY * (Y > 0) in place (act == ReLU; the forward stored Y = relu(Z)), then db[n] = sum_m dZ[m, n].
// grid = (ceil(N / 32), row chunks) blocks of 256 threads: 32 columns x 8 row lanes, fixed-order tree.  One row chunk
// (M <= ACT_BWD_CHUNK): deterministic sums; more: the chunks' partial sums meet in db through atomicAdd (db zeroed by the
// host entry unless accumulating), so large-M launches fill the GPU instead of ceil(N / 32) SMs.
constexpr int64_t ACT_BWD_CHUNK = 2048;
__global__ void __launch_bounds__(256)
act_bias_bwd_kernel(float* __restrict__ dy, int64_t lddy, const float* __restrict__ y, int64_t ldy, int64_t M, int64_t N,
                    int act, float* __restrict__ db, int accumulate) {
  __shared__ float part[8][33];
  const int c = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int64_t n = static_cast<int64_t>(blockIdx.x) * 32 + c;
  const int64_t m_lo = gridDim.y > 1 ? static_cast<int64_t>(blockIdx.y) * ACT_BWD_CHUNK : 0;
  const int64_t m_hi = gridDim.y > 1 ? (m_lo + ACT_BWD_CHUNK < M ? m_lo + ACT_BWD_CHUNK : M) : M;
  float s = 0.f;
  if (n < N) {
    for (int64_t m = m_lo + rl; m < m_hi; m += 8) {
      float g = dy[m * lddy + n];
      if (act == HOISDF_ACT_RELU && !(y[m * ldy + n] > 0.f)) {
        g = 0.f;
        dy[m * lddy + n] = 0.f;
      }
      s += g;
    }
  }
  part[rl][c] = s;
  __syncthreads();
  if (rl == 0 && n < N && db != nullptr) {
    float t = part[0][c];
#pragma unroll
    for (int i = 1; i < 8; ++i) t += part[i][c];
    if (gridDim.y > 1) atomicAdd(db + n, t);
    else db[n] = accumulate ? db[n] + t : t;
  }
}

// W[r, :] = g[r] * v[r, :] / |v[r, :]|:  dg[r] = <dW[r], v[r]> / |v[r]|,  dv[r] = g[r] / |v[r]| * (dW[r] - dg[r] * v[r] / |v[r]|)
__global__ void __launch_bounds__(256)
weight_norm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ v, const float* __restrict__ dw,
                       int64_t lddw, int64_t cols, float* __restrict__ dg, float* __restrict__ dv, int accumulate) {
  __shared__ float red[2][8];
  const int64_t r = blockIdx.x;
  const float* vr = v + r * cols;
  const float* dr = dw + r * lddw;
  float ss = 0.f, dot = 0.f;
  for (int64_t c = threadIdx.x; c < cols; c += 256) {
    ss = fmaf(vr[c], vr[c], ss);
    dot = fmaf(dr[c], vr[c], dot);
  }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ss; red[1][threadIdx.x >> 5] = dot; }
  __syncthreads();
  float tss = 0.f, tdot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { tss += red[0][i]; tdot += red[1][i]; }
  const float norm = sqrtf(tss);
  const float dgr = tdot / norm;
  const float scale = g[r] / norm;
  if (threadIdx.x == 0) dg[r] = accumulate ? dg[r] + dgr : dgr;
  for (int64_t c = threadIdx.x; c < cols; c += 256) {
    const float val = scale * (dr[c] - dgr * vr[c] / norm);
    dv[r * cols + c] = accumulate ? dv[r * cols + c] + val : val;
  }
}

struct GatherBwdParams {
  hoisdf_pyramid grad;            // NHWC gradient maps (zero-initialised or accumulating), same geometry as the forward
  const float* __restrict__ uv;
  const int64_t* __restrict__ row_offsets;
  const float* __restrict__ dout; // (rows, ld) gradient of the CONCAT output
  int64_t rows, batch, rows_per_sample, ld;
};

__device__ __forceinline__ void atomic_add_f32(float* p, float v) {
#ifdef HOISDF_EMULATE
  unsigned old = __atomic_load_n(reinterpret_cast<unsigned*>(p), __ATOMIC_RELAXED), neu;
  do {
    neu = __float_as_uint(__uint_as_float(old) + v);
  } while (!__atomic_compare_exchange_n(reinterpret_cast<unsigned*>(p), &old, neu, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
#else
  atomicAdd(p, v);
#endif
}

// one warp per row, the forward's tap arithmetic (ATen grid_sampler_2d, align_corners=True, border padding)
__global__ void __launch_bounds__(256) gather_concat_bwd_kernel(const GatherBwdParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= p.rows) return;
  int64_t b;
  if (p.row_offsets == nullptr) {
    b = r / p.rows_per_sample;
  } else {
    int64_t lo = 0, hi = p.batch;
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (p.row_offsets[mid] <= r) lo = mid; else hi = mid;
    }
    b = lo;
  }
  const float u = p.uv[r * 2 + 0], v = p.uv[r * 2 + 1];
  int off = 0;
  for (int l = 0; l < p.grad.levels; ++l) {
    const int C = p.grad.c[l], W = p.grad.w[l], H = p.grad.h[l];
    const float nx = static_cast<float>(p.grad.img_w - 1) / 2.0f, ny = static_cast<float>(p.grad.img_h - 1) / 2.0f;
    const float gx = __fdiv_rn(__fsub_rn(u, nx), nx), gy = __fdiv_rn(__fsub_rn(v, ny), ny);
    float x = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), static_cast<float>(W - 1));
    float y = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), static_cast<float>(H - 1));
    x = fminf(static_cast<float>(W - 1), fmaxf(x, 0.f));
    y = fminf(static_cast<float>(H - 1), fmaxf(y, 0.f));
    const float x0f = floorf(x), y0f = floorf(y);
    const float tx = x - x0f, ty = y - y0f;
    const int x0 = static_cast<int>(x0f), y0 = static_cast<int>(y0f);
    const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
    const float ax = __fsub_rn(x0f + 1.f, x), ay = __fsub_rn(y0f + 1.f, y);
    float w00 = ax * ay, w01 = tx * ay, w10 = ax * ty, w11 = tx * ty;
    if (x0 + 1 > W - 1) { w01 = 0.f; w11 = 0.f; }
    if (y0 + 1 > H - 1) { w10 = 0.f; w11 = 0.f; }
    const int64_t base = b * static_cast<int64_t>(H) * W;
    float* m = const_cast<float*>(p.grad.map[l]);
    float* t00 = m + (base + static_cast<int64_t>(y0) * W + x0) * C;
    float* t01 = m + (base + static_cast<int64_t>(y0) * W + x1) * C;
    float* t10 = m + (base + static_cast<int64_t>(y1) * W + x0) * C;
    float* t11 = m + (base + static_cast<int64_t>(y1) * W + x1) * C;
    const float* d = p.dout + r * p.ld + off;
    for (int c = lane; c < C; c += 32) {
      const float gval = d[c];
      if (w00 != 0.f) atomic_add_f32(t00 + c, gval * w00);
      if (w01 != 0.f) atomic_add_f32(t01 + c, gval * w01);
      if (w10 != 0.f) atomic_add_f32(t10 + c, gval * w10);
      if (w11 != 0.f) atomic_add_f32(t11 + c, gval * w11);
    }
    off += C;
  }
}

// SDFLoss on the decoder's pre-activation z (sdf = tanh(z)):
//   loss = mean_i | clamp(tanh(z_i), -c, c) - clamp(gt_i, -c, c) |   (upstream common/nets/loss.py:64-78 SepSDFLoss = L1,
//   reduction mean; the clamps to ClampingDistance are main/model.py:241 and :388-395)
//   dz_i = scale / n * sign(pred_i - gt_i) * [|tanh(z_i)| < c] * (1 - tanh(z_i)^2)
__global__ void sdf_loss_bwd_kernel(const float* __restrict__ z, const float* __restrict__ gt, int64_t n, float clamp,
                                    float scale, float* __restrict__ dz) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float t = tanhf(z[i]);
  const float pred = fminf(fmaxf(t, -clamp), clamp), tgt = fminf(fmaxf(gt[i], -clamp), clamp);
  const float diff = pred - tgt;
  const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
  const bool pass = t >= -clamp && t <= clamp;        // torch.clamp passes the gradient inside (and at) the bounds
  dz[i] = pass ? scale / static_cast<float>(n) * sgn * (1.f - t * t) : 0.f;
}


// ---- transformer side: LayerNorm and row softmax (the attention core's backward is these + hoisdf_gemm_f32 per head)

// y = (h - mean) * rstd * gamma + beta over rows of D = 256 (eps 1e-5, biased variance).  One warp per row:
//   g = dy * gamma;  dh = rstd * (g - mean(g) - xhat * mean(g * xhat));  stats[r] = (mean, rstd) for the column pass
template <int D>
__global__ void __launch_bounds__(256)
layernorm_bwd_rows_kernel(const float* __restrict__ h, const float* __restrict__ gamma, const float* __restrict__ dy,
                          int64_t rows, float* __restrict__ dh, float* __restrict__ stats) {
  constexpr int Q = D / 32;
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float v[Q], g[Q];
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < Q; ++q) { v[q] = h[r * D + q * 32 + lane]; s += v[q]; }
  const float mean = warp_sum(s) * (1.0f / D);
  float ss = 0.f;
#pragma unroll
  for (int q = 0; q < Q; ++q) { const float d = v[q] - mean; ss = fmaf(d, d, ss); }
  const float rstd = 1.0f / sqrtf(warp_sum(ss) * (1.0f / D) + 1e-5f);
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int c = q * 32 + lane;
    v[q] = (v[q] - mean) * rstd;                        // xhat
    g[q] = dy[r * D + c] * gamma[c];
    sg += g[q];
    sgx = fmaf(g[q], v[q], sgx);
  }
  const float mg = warp_sum(sg) * (1.0f / D), mgx = warp_sum(sgx) * (1.0f / D);
#pragma unroll
  for (int q = 0; q < Q; ++q) dh[r * D + q * 32 + lane] = rstd * (g[q] - mg - v[q] * mgx);
  if (lane == 0 && stats != nullptr) { stats[r * 2] = mean; stats[r * 2 + 1] = rstd; }
}

// dgamma[c] = sum_r dy[r, c] * xhat[r, c], dbeta[c] = sum_r dy[r, c]; 32 columns x 8 row lanes per block, fixed-order sums.
// grid.y > 1 (many rows): every block owns LN_COLS_CHUNK rows and the chunks' partial sums meet through atomicAdd (the host
// entry zeroes dgamma / dbeta first unless accumulating) -- 8 column blocks alone leave 140 SMs idle for 1.3 ms per call at
// 51 200 rows
constexpr int64_t LN_COLS_CHUNK = 1024;
__global__ void __launch_bounds__(256)
layernorm_bwd_cols_kernel(const float* __restrict__ h, const float* __restrict__ dy, const float* __restrict__ stats,
                          int64_t rows, int D, float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
  __shared__ float pg[8][33], pb[8][33];
  const int c = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + c;
  const int64_t r_lo = gridDim.y > 1 ? static_cast<int64_t>(blockIdx.y) * LN_COLS_CHUNK : 0;
  const int64_t r_hi = gridDim.y > 1 ? (r_lo + LN_COLS_CHUNK < rows ? r_lo + LN_COLS_CHUNK : rows) : rows;
  float sg = 0.f, sb = 0.f;
  if (n < D) {
    for (int64_t r = r_lo + rl; r < r_hi; r += 8) {
      const float d = dy[r * D + n];
      sg = fmaf(d, (h[r * D + n] - stats[r * 2]) * stats[r * 2 + 1], sg);
      sb += d;
    }
  }
  pg[rl][c] = sg; pb[rl][c] = sb;
  __syncthreads();
  if (rl == 0 && n < D) {
    float tg = pg[0][c], tb = pb[0][c];
#pragma unroll
    for (int i = 1; i < 8; ++i) { tg += pg[i][c]; tb += pb[i][c]; }
    if (gridDim.y > 1) {
      atomicAdd(dgamma + n, tg);
      atomicAdd(dbeta + n, tb);
    } else {
      dgamma[n] = accumulate ? dgamma[n] + tg : tg;
      dbeta[n] = accumulate ? dbeta[n] + tb : tb;
    }
  }
}

// p[r, :] = softmax over the columns c < valid that are not blocked by mask[r % mask_rows, c] (bool attn_mask of
// nn.MultiheadAttention: non-zero = blocked); blocked / invalid columns get exactly 0.  One warp per row.
__global__ void __launch_bounds__(256)
softmax_rows_fwd_kernel(const float* __restrict__ s, int64_t lds, int64_t rows, int cols, int valid,
                        const uint8_t* __restrict__ mask, int64_t mask_rows, float* __restrict__ p, int64_t ldp) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const uint8_t* mrow = mask != nullptr ? mask + (r % mask_rows) * cols : nullptr;
  float mx = -3.402823466e+38f;
  for (int c = lane; c < valid; c += 32)
    if (mrow == nullptr || mrow[c] == 0) mx = fmaxf(mx, s[r * lds + c]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < valid; c += 32)
    if (mrow == nullptr || mrow[c] == 0) sum += expf(s[r * lds + c] - mx);
  sum = warp_sum(sum);
  for (int c = lane; c < cols; c += 32) {
    const bool on = c < valid && (mrow == nullptr || mrow[c] == 0);
    p[r * ldp + c] = on ? expf(s[r * lds + c] - mx) / sum : 0.f;
  }
}

// ds[r, :] = p[r, :] * (dp[r, :] - sum_j dp[r, j] * p[r, j]), one warp per row (in place over dp when ds == dp)
__global__ void __launch_bounds__(256)
softmax_rows_bwd_kernel(const float* __restrict__ p, int64_t ldp, const float* dp, int64_t lddp, int64_t rows, int cols,
                        float* ds, int64_t ldds) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float dot = 0.f;
  for (int c = lane; c < cols; c += 32) dot = fmaf(dp[r * lddp + c], p[r * ldp + c], dot);
  dot = warp_sum(dot);
  for (int c = lane; c < cols; c += 32) ds[r * ldds + c] = p[r * ldp + c] * (dp[r * lddp + c] - dot);
}


// ---- nn.MultiheadAttention's dropout on the attention probabilities (upstream cfg.dropout = 0.1, transformer.py layers)
// without materialising a mask: keep(r, c) is a counter-based hash of (seed, r, c) (common.cuh), so the backward regenerates
// the very same decisions from the seed.  The forward emits the dropped, rescaled probabilities pd = keep ? p / (1 - q) : 0 (and
// optionally p itself); the backward folds the mask into the softmax derivative:
//   g = keep ? dpd / (1 - q) : 0;   ds = p * (g - sum_j g_j p_j).
__global__ void __launch_bounds__(256)
softmax_dropout_rows_fwd_kernel(const float* s, int64_t lds, int64_t rows, int cols, int valid,
                                const uint8_t* __restrict__ mask, int64_t mask_rows, float* p, int64_t ldp, float* pd,
                                int64_t ldpd, float p_drop, uint64_t seed) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const uint8_t* mrow = mask != nullptr ? mask + (r % mask_rows) * cols : nullptr;
  float mx = -3.402823466e+38f;
  for (int c = lane; c < valid; c += 32)
    if (mrow == nullptr || mrow[c] == 0) mx = fmaxf(mx, s[r * lds + c]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < valid; c += 32)
    if (mrow == nullptr || mrow[c] == 0) sum += expf(s[r * lds + c] - mx);
  sum = warp_sum(sum);
  const float keep_scale = 1.0f / (1.0f - p_drop);
  const uint32_t key = dropout_row_key(seed, static_cast<uint64_t>(r)), thr = dropout_threshold(p_drop);
  for (int c = lane; c < cols; c += 32) {
    const bool on = c < valid && (mrow == nullptr || mrow[c] == 0);
    const float pv = on ? expf(s[r * lds + c] - mx) / sum : 0.f;      // (s may alias p or pd: read before either write)
    const bool keep = dropout_keep(key, static_cast<uint32_t>(c), thr);
    if (p != nullptr) p[r * ldp + c] = pv;
    pd[r * ldpd + c] = keep ? pv * keep_scale : 0.f;
  }
}

__global__ void __launch_bounds__(256)
softmax_dropout_rows_bwd_kernel(const float* __restrict__ p, int64_t ldp, const float* dpd, int64_t lddp, int64_t rows, int cols,
                                float* ds, int64_t ldds, float p_drop, uint64_t seed) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float keep_scale = 1.0f / (1.0f - p_drop);
  const uint32_t key = dropout_row_key(seed, static_cast<uint64_t>(r)), thr = dropout_threshold(p_drop);
  float dot = 0.f;
  for (int c = lane; c < cols; c += 32) {
    const float g = dropout_keep(key, static_cast<uint32_t>(c), thr) ? dpd[r * lddp + c] * keep_scale : 0.f;
    dot = fmaf(g, p[r * ldp + c], dot);
  }
  dot = warp_sum(dot);
  for (int c = lane; c < cols; c += 32) {
    const float g = dropout_keep(key, static_cast<uint32_t>(c), thr) ? dpd[r * lddp + c] * keep_scale : 0.f;
    ds[r * ldds + c] = p[r * ldp + c] * (g - dot);
  }
}

// ---- Linear layers with a handful of output features (n <= 16: the SDF value, class and offset heads; upstream
// common/nets/sdf_net.py:53-64 `linh4`, main/model.py:82-91) over tens of thousands of rows.  A 64-wide GEMM tile wastes most of
// its work there and the weight gradient is a tall reduction (k = rows): both are one streaming pass over X instead.
//   thin_linear_fwd   y (m, n) = act(x (m, k) . w (n, k)^T + b): one warp per row, lanes stride over k
//   thin_linear_dw    dw (n, k) = dz (m, n)^T . x (m, k): CTA = 256 columns of x times a chunk of rows, dz rows staged in
//                     shared memory, per-thread partial sums, one atomicAdd per (output feature, column) and CTA
constexpr int THIN_MAX_N = 16;

__global__ void __launch_bounds__(256)
thin_linear_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w, int64_t ldw,
                       const float* __restrict__ bias, int64_t m, int k, int n, int act, float* __restrict__ y, int64_t ldy) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp; r < m; r += nwarps) {
    float acc[THIN_MAX_N];
#pragma unroll
    for (int j = 0; j < THIN_MAX_N; ++j) acc[j] = 0.f;
    for (int c = lane; c < k; c += 32) {
      const float xv = x[r * ldx + c];
#pragma unroll
      for (int j = 0; j < THIN_MAX_N; ++j)
        if (j < n) acc[j] = fmaf(xv, w[j * ldw + c], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < THIN_MAX_N; ++j) {
      if (j < n) {                                     // (n is warp-uniform)
        float v = warp_sum(acc[j]);
        if (lane == 0) {
          if (bias != nullptr) v += bias[j];
          if (act == HOISDF_ACT_RELU) v = fmaxf(v, 0.f);
          y[r * ldy + j] = v;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256)
thin_linear_dw_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ dz, int64_t lddz, int64_t m, int k,
                      int n, int64_t rows_per_cta, float* __restrict__ dw, int64_t lddw) {
  __shared__ float dzs[64][THIN_MAX_N];
  const int c = static_cast<int>(blockIdx.x) * 256 + static_cast<int>(threadIdx.x);
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < m ? r0 + rows_per_cta : m;
  float acc[THIN_MAX_N];
#pragma unroll
  for (int j = 0; j < THIN_MAX_N; ++j) acc[j] = 0.f;
  for (int64_t rb = r0; rb < r1; rb += 64) {
    const int cnt = static_cast<int>(r1 - rb < 64 ? r1 - rb : 64);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * n; e += 256) {
      const int rr = e / n, j = e - rr * n;
      dzs[rr][j] = dz[(rb + rr) * lddz + j];
    }
    __syncthreads();
    if (c < k) {
      for (int rr = 0; rr < cnt; ++rr) {
        const float xv = x[(rb + rr) * ldx + c];
#pragma unroll
        for (int j = 0; j < THIN_MAX_N; ++j)
          if (j < n) acc[j] = fmaf(xv, dzs[rr][j], acc[j]);
      }
    }
  }
  if (c < k) {
#pragma unroll
    for (int j = 0; j < THIN_MAX_N; ++j)
      if (j < n) atomicAdd(dw + j * lddw + c, acc[j]);
  }
}

// torch.optim.AdamW (upstream common/base.py:68: lr 1e-4, default betas / eps / weight_decay 0.01), one fused pass over a
// flat parameter buffer, the arithmetic in the order PyTorch's single-tensor implementation applies it
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, float step_size,
                             float bias2_sqrt) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float grad = g[i];
  float w = p[i] * (1.f - lr * weight_decay);                  // param.mul_(1 - lr * weight_decay)
  const float mi = m[i] + (grad - m[i]) * (1.f - beta1);       // exp_avg.lerp_(grad, 1 - beta1)
  const float vi = fmaf(grad * grad, 1.f - beta2, v[i] * beta2);   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(vi) / bias2_sqrt + eps;
  w -= step_size * (mi / denom);                               // param.addcdiv_(exp_avg, denom, value=-step_size)
  p[i] = w; m[i] = mi; v[i] = vi;
}


// ---- JointvoteLoss backward (upstream common/nets/loss.py:22-61), batch-major like the forward vote kernel:
//   points (B,P,3) [m], off (L,B,P,60), cls (L,B,P,20), joint_gt (B,20,3) [mm]
//   vote = point + off;  mask[b,p,j] = |point - gt_j / 1000| < cls_dist;  n_pos = sum(mask)
//   loss_joint_3d     = sum_{l,b,p,j,c} smooth_l1(1000 vote - gt) * mask / (3 L n_pos)
//   loss_joint_cls    = mean_{l,b,p,j} BCEWithLogits(cls, mask)
//   loss_all_joint_3d = mean_{l,b,j,c} smooth_l1(1000 * sum_p softmax_p(cls) vote - gt)
// given the three upstream gradients g1, g2, g3 (the training loop sums the losses: all 1).  The points carry no gradient
// (selected lattice points / jittered pre-points).
constexpr int kVoteJ = 20;

__global__ void __launch_bounds__(256)
vote_count_pos_kernel(const float* __restrict__ points, const float* __restrict__ gt, int64_t batch, int64_t p, float dist,
                      float* __restrict__ npos) {
  __shared__ float red[8];
  float n = 0.f;
  const int64_t total = batch * p * kVoteJ;
  for (int64_t i = threadIdx.x; i < total; i += 256) {
    const int j = static_cast<int>(i % kVoteJ);
    const int64_t bp = i / kVoteJ, b = bp / p;
    float d2 = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = points[bp * 3 + c] - gt[(b * kVoteJ + j) * 3 + c] / 1000.f;
      d2 += d * d;
    }
    n += sqrtf(d2) < dist ? 1.f : 0.f;
  }
  n = warp_sum(n);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    npos[0] = t;
  }
}

__device__ __forceinline__ float smooth_l1_grad(float x) { return fminf(fmaxf(x, -1.f), 1.f); }   // beta = 1

// one CTA per (layer, sample): warps 0..7 own joints j = w, w + 8, w + 16 for the softmax statistics, then all threads
// walk the (point, joint) pairs
__global__ void __launch_bounds__(256)
vote_loss_bwd_kernel(const float* __restrict__ points, const float* __restrict__ off, const float* __restrict__ cls,
                     const float* __restrict__ gt, int64_t layers, int64_t batch, int64_t p, float dist, float g1, float g2,
                     float g3, const float* __restrict__ npos, float* __restrict__ d_off, float* __restrict__ d_cls) {
  __shared__ float s_max[kVoteJ], s_sum[kVoteJ], s_joint[kVoteJ][3];
  const int64_t lb = blockIdx.x, b = lb % batch;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* pts = points + b * p * 3;
  const float* o = off + lb * p * 60;
  const float* c = cls + lb * p * kVoteJ;
  const float* g = gt + b * kVoteJ * 3;
  for (int j = warp; j < kVoteJ; j += 8) {
    float mx = -3.402823466e+38f;
    for (int64_t i = lane; i < p; i += 32) mx = fmaxf(mx, c[i * kVoteJ + j]);
    mx = warp_max(mx);
    float se = 0.f, sv[3] = {0.f, 0.f, 0.f};
    for (int64_t i = lane; i < p; i += 32) {
      const float e = expf(c[i * kVoteJ + j] - mx);
      se += e;
#pragma unroll
      for (int k = 0; k < 3; ++k) sv[k] = fmaf(e, pts[i * 3 + k] + o[i * 60 + j * 3 + k], sv[k]);
    }
    se = warp_sum(se);
#pragma unroll
    for (int k = 0; k < 3; ++k) sv[k] = warp_sum(sv[k]);
    if (lane == 0) {
      s_max[j] = mx; s_sum[j] = se;
#pragma unroll
      for (int k = 0; k < 3; ++k) s_joint[j][k] = sv[k] / se;
    }
  }
  __syncthreads();
  const float k1 = g1 * 1000.f / (3.f * static_cast<float>(layers) * npos[0]);
  const float k2 = g2 / static_cast<float>(layers * batch * p * kVoteJ);
  const float k3 = g3 * 1000.f / static_cast<float>(layers * batch * kVoteJ * 3);
  for (int64_t e = threadIdx.x; e < p * kVoteJ; e += 256) {
    const int64_t i = e / kVoteJ;
    const int j = static_cast<int>(e - i * kVoteJ);
    float d2 = 0.f, vote[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float d = pts[i * 3 + k] - g[j * 3 + k] / 1000.f;
      d2 += d * d;
      vote[k] = pts[i * 3 + k] + o[i * 60 + j * 3 + k];
    }
    const float mask = sqrtf(d2) < dist ? 1.f : 0.f;
    const float logit = c[i * kVoteJ + j];
    const float w = expf(logit - s_max[j]) / s_sum[j];
    float dc = k2 * (1.f / (1.f + expf(-logit)) - mask);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float dj = k3 * smooth_l1_grad(1000.f * s_joint[j][k] - g[j * 3 + k]);      // d loss / d joints[j][k]
      d_off[lb * p * 60 + i * 60 + j * 3 + k] = k1 * mask * smooth_l1_grad(1000.f * vote[k] - g[j * 3 + k]) + dj * w;
      dc = fmaf(dj * w, vote[k] - s_joint[j][k], dc);
    }
    d_cls[lb * p * kVoteJ + e] = dc;
  }
}


// ---- token assembly backward (upstream main/model.py:123-126,520-531): tokens[.., 33 + c] = fea[c] * sig,
// sig = sigmoid(sdf / beta) / beta.  d_fea = d_tok * sig;  d_sdf = <d_tok, fea> * s (1 - s) / beta^2;
// d_beta = sum over points of <d_tok, fea> * (-s / beta^2 - s (1 - s) sdf / beta^3)   (s = sigmoid(sdf / beta)).
// One warp per point; d_beta partials per block (fixed order), folded by the second kernel.
__global__ void __launch_bounds__(256)
tokens_bwd_kernel(const float* __restrict__ d_tok, int64_t s_total, int64_t t0, const float* __restrict__ fea, int64_t ld_fea,
                  const float* __restrict__ sdf, const float* __restrict__ beta, int64_t batch, int64_t p,
                  float* __restrict__ d_fea, int64_t ld_dfea, float* __restrict__ d_sdf, float* __restrict__ partial) {
  __shared__ float red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t bt = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  float db = 0.f;
  if (bt < batch * p) {
    const int64_t b = bt / p, t = bt - b * p;
    const float be = beta[0];
    const float z = __fdiv_rn(sdf[bt], be);
    const float sg = __fdiv_rn(1.f, 1.f + expf(-z));
    const float sig = __fdiv_rn(sg, be);
    const float* dt = d_tok + (b * s_total + t0 + t) * 256 + 33;
    float dot = 0.f;
    for (int c = lane; c < 223; c += 32) {
      const float g = dt[c];
      d_fea[bt * ld_dfea + c] = g * sig;
      dot = fmaf(g, fea[bt * ld_fea + c], dot);
    }
    dot = warp_sum(dot);
    const float ds = sg * (1.f - sg);
    if (lane == 0 && d_sdf != nullptr) d_sdf[bt] = dot * ds / (be * be);
    db = dot * (-sg / (be * be) - ds * sdf[bt] / (be * be * be));
  }
  if (lane == 0) red[warp] = db;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tsum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tsum += red[i];
    partial[blockIdx.x] = tsum;
  }
}

__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ partial, int64_t n, float* __restrict__ out,
                                                           int accumulate) {
  __shared__ float red[8];
  float sacc = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 256) sacc += partial[i];
  sacc = warp_sum(sacc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tsum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tsum += red[i];
    out[0] = accumulate ? out[0] + tsum : tsum;
  }
}

}  // namespace
}  // namespace hoisdf

using namespace hoisdf;

static int gemm_f32_launch(const float* a, int64_t lda, int32_t trans_a, const float* b, int64_t ldb, int32_t trans_b, float* c,
                           int64_t ldc, int64_t m, int64_t n, int64_t k, int32_t accumulate, const GemmBatch& gb, int64_t batches,
                           void* stream) {
  if (a == nullptr || b == nullptr || c == nullptr) return HOISDF_E_NULL;
  if (m <= 0 || n <= 0 || k <= 0 || ldc < n || batches <= 0 || batches > 65535) return HOISDF_E_SHAPE;
  if (lda < (trans_a ? m : k) || ldb < (trans_b ? k : n)) return HOISDF_E_SHAPE;
  const int64_t gx = ceil_div(n, GB), gy = ceil_div(m, GB);
  if (gy > 65535 || gx > 0x7fffffffLL) return HOISDF_E_SHAPE;
  const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(gy), static_cast<unsigned>(batches));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int acc = accumulate ? 1 : 0;
  if (trans_a && trans_b) HOISDF_LAUNCH((gemm_f32_kernel<true, true>), grid, 256, s, a, lda, b, ldb, c, ldc, m, n, k, acc, gb);
  else if (trans_a) HOISDF_LAUNCH((gemm_f32_kernel<true, false>), grid, 256, s, a, lda, b, ldb, c, ldc, m, n, k, acc, gb);
  else if (trans_b) HOISDF_LAUNCH((gemm_f32_kernel<false, true>), grid, 256, s, a, lda, b, ldb, c, ldc, m, n, k, acc, gb);
  else HOISDF_LAUNCH((gemm_f32_kernel<false, false>), grid, 256, s, a, lda, b, ldb, c, ldc, m, n, k, acc, gb);
  return launch_status();
}

HOISDF_API int hoisdf_gemm_f32(const float* a, int64_t lda, int32_t trans_a, const float* b, int64_t ldb, int32_t trans_b,
                               float* c, int64_t ldc, int64_t m, int64_t n, int64_t k, int32_t accumulate, void* stream) {
  GemmBatch gb{0, 0, 0, 0, 0, 0, 1, 1.0f};
  return gemm_f32_launch(a, lda, trans_a, b, ldb, trans_b, c, ldc, m, n, k, accumulate, gb, 1, stream);
}

HOISDF_API int hoisdf_gemm_f32_batched(const float* a, int64_t lda, int32_t trans_a, int64_t a_outer, int64_t a_inner,
                                       const float* b, int64_t ldb, int32_t trans_b, int64_t b_outer, int64_t b_inner, float* c,
                                       int64_t ldc, int64_t c_outer, int64_t c_inner, int64_t m, int64_t n, int64_t k, float alpha,
                                       int32_t accumulate, int64_t batch_outer, int64_t batch_inner, void* stream) {
  if (batch_outer <= 0 || batch_inner <= 0 || batch_inner > 65535) return HOISDF_E_SHAPE;
  GemmBatch gb{a_outer, a_inner, b_outer, b_inner, c_outer, c_inner, static_cast<int>(batch_inner), alpha};
  return gemm_f32_launch(a, lda, trans_a, b, ldb, trans_b, c, ldc, m, n, k, accumulate, gb, batch_outer * batch_inner, stream);
}

HOISDF_API int hoisdf_act_bias_bwd(float* dy, int64_t lddy, const float* y, int64_t ldy, int64_t m, int64_t n, int32_t act,
                                   float* db, int32_t accumulate, void* stream) {
  if (dy == nullptr || (act == HOISDF_ACT_RELU && y == nullptr)) return HOISDF_E_NULL;
  if (m <= 0 || n <= 0 || lddy < n || (y != nullptr && ldy < n)) return HOISDF_E_SHAPE;
  if (act != HOISDF_ACT_NONE && act != HOISDF_ACT_RELU) return HOISDF_E_UNSUPPORTED;
  const int64_t chunks = m > 4 * ACT_BWD_CHUNK ? ceil_div(m, ACT_BWD_CHUNK) : 1;
  if (chunks > 65535) return HOISDF_E_SHAPE;
  if (chunks > 1 && db != nullptr && !accumulate) {
    const cudaError_t e = cudaMemsetAsync(db, 0, sizeof(float) * n, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  const dim3 grid(static_cast<unsigned>(ceil_div(n, 32)), static_cast<unsigned>(chunks));
  HOISDF_LAUNCH(act_bias_bwd_kernel, grid, 256, static_cast<cudaStream_t>(stream), dy, lddy, y, ldy, m, n, act, db,
                accumulate ? 1 : 0);
  return launch_status();
}

HOISDF_API int hoisdf_thin_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, int64_t m,
                                      int64_t k, int64_t n, int32_t act, float* y, int64_t ldy, void* stream) {
  if (x == nullptr || w == nullptr || y == nullptr) return HOISDF_E_NULL;
  if (m <= 0 || k <= 0 || k > 0x7fffffffLL || n <= 0 || ldx < k || ldw < k || ldy < n) return HOISDF_E_SHAPE;
  if (n > THIN_MAX_N || (act != HOISDF_ACT_NONE && act != HOISDF_ACT_RELU)) return HOISDF_E_UNSUPPORTED;
  const int64_t want = ceil_div(m, 8);
  const unsigned blocks = static_cast<unsigned>(want < kNumSMs * 16 ? want : kNumSMs * 16);
  HOISDF_LAUNCH(thin_linear_fwd_kernel, blocks, 256, static_cast<cudaStream_t>(stream), x, ldx, w, ldw, bias, m,
                static_cast<int>(k), static_cast<int>(n), act, y, ldy);
  return launch_status();
}

HOISDF_API int hoisdf_thin_linear_dw(const float* x, int64_t ldx, const float* dz, int64_t lddz, int64_t m, int64_t k,
                                     int64_t n, float* dw, int64_t lddw, int32_t accumulate, void* stream) {
  if (x == nullptr || dz == nullptr || dw == nullptr) return HOISDF_E_NULL;
  if (m <= 0 || k <= 0 || k > 0x7fffffffLL || n <= 0 || ldx < k || lddz < n || lddw < k) return HOISDF_E_SHAPE;
  if (n > THIN_MAX_N) return HOISDF_E_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!accumulate) {
    for (int64_t j = 0; j < (lddw == k ? 1 : n); ++j) {
      const cudaError_t e = cudaMemsetAsync(dw + j * lddw, 0, sizeof(float) * (lddw == k ? n * k : k), s);
      if (e != cudaSuccess) return static_cast<int>(e);
    }
  }
  const int64_t gx = ceil_div(k, 256);
  int64_t chunks = (kNumSMs * 4) / gx;                       // ~4 CTAs per SM in total
  if (chunks < 1) chunks = 1;
  int64_t rows_per_cta = ceil_div(ceil_div(m, chunks), 64) * 64;
  chunks = ceil_div(m, rows_per_cta);
  if (chunks > 65535) return HOISDF_E_SHAPE;
  const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(chunks));
  HOISDF_LAUNCH(thin_linear_dw_kernel, grid, 256, s, x, ldx, dz, lddz, m, static_cast<int>(k), static_cast<int>(n),
                rows_per_cta, dw, lddw);
  return launch_status();
}

HOISDF_API int hoisdf_weight_norm_bwd(const float* g, const float* v, const float* dw, int64_t lddw, int64_t rows,
                                      int64_t cols, float* dg, float* dv, int32_t accumulate, void* stream) {
  if (g == nullptr || v == nullptr || dw == nullptr || dg == nullptr || dv == nullptr) return HOISDF_E_NULL;
  if (rows <= 0 || rows > 0x7fffffffLL || cols <= 0 || lddw < cols) return HOISDF_E_SHAPE;
  HOISDF_LAUNCH(weight_norm_bwd_kernel, static_cast<unsigned>(rows), 256, static_cast<cudaStream_t>(stream), g, v, dw, lddw,
                cols, dg, dv, accumulate ? 1 : 0);
  return launch_status();
}

HOISDF_API int hoisdf_gather_bwd(const hoisdf_pyramid* grad, const float* uv, int64_t rows, const int64_t* row_offsets,
                                 int64_t batch, int64_t rows_per_sample, const float* dout, int64_t ld_dout, void* stream) {
  if (grad == nullptr || uv == nullptr || dout == nullptr) return HOISDF_E_NULL;
  if (rows == 0) return HOISDF_OK;
  if (rows < 0 || batch <= 0 || grad->levels < 1 || grad->levels > 5) return HOISDF_E_SHAPE;
  if (row_offsets == nullptr && rows_per_sample <= 0) return HOISDF_E_SHAPE;
  int ctot = 0;
  for (int l = 0; l < grad->levels; ++l) {
    if (grad->map[l] == nullptr) return HOISDF_E_NULL;
    if (grad->c[l] <= 0 || grad->h[l] <= 0 || grad->w[l] <= 0) return HOISDF_E_SHAPE;
    ctot += grad->c[l];
  }
  if (ld_dout < ctot) return HOISDF_E_SHAPE;
  GatherBwdParams p;
  p.grad = *grad; p.uv = uv; p.row_offsets = row_offsets; p.dout = dout;
  p.rows = rows; p.batch = batch; p.rows_per_sample = rows_per_sample; p.ld = ld_dout;
  HOISDF_LAUNCH(gather_concat_bwd_kernel, static_cast<unsigned>(ceil_div(rows, 8)), 256, static_cast<cudaStream_t>(stream), p);
  return launch_status();
}

HOISDF_API int hoisdf_sdf_loss_bwd(const float* z, const float* sdf_gt, int64_t n, float clamp, float scale, float* dz,
                                   void* stream) {
  if (z == nullptr || sdf_gt == nullptr || dz == nullptr) return HOISDF_E_NULL;
  if (n <= 0 || !(clamp > 0.f)) return HOISDF_E_SHAPE;
  HOISDF_LAUNCH(sdf_loss_bwd_kernel, static_cast<unsigned>(ceil_div(n, 256)), 256, static_cast<cudaStream_t>(stream), z, sdf_gt,
                n, clamp, scale, dz);
  return launch_status();
}

HOISDF_API int hoisdf_layernorm_bwd(const float* h, const float* gamma, const float* dy, int64_t rows, int64_t d, float* dh,
                                    float* dgamma, float* dbeta, float* stats, int32_t accumulate, void* stream) {
  if (h == nullptr || gamma == nullptr || dy == nullptr || dh == nullptr) return HOISDF_E_NULL;
  if ((dgamma == nullptr) != (dbeta == nullptr) || (dgamma != nullptr && stats == nullptr)) return HOISDF_E_NULL;
  if (rows <= 0) return HOISDF_E_SHAPE;
  if (d != 256) return HOISDF_E_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  HOISDF_LAUNCH(layernorm_bwd_rows_kernel<256>, static_cast<unsigned>(ceil_div(rows, 8)), 256, s, h, gamma, dy, rows, dh, stats);
  if (dgamma != nullptr) {
    const int64_t chunks = rows > 4 * LN_COLS_CHUNK ? ceil_div(rows, LN_COLS_CHUNK) : 1;
    if (chunks > 65535) return HOISDF_E_SHAPE;
    if (chunks > 1 && !accumulate) {
      if (cudaMemsetAsync(dgamma, 0, sizeof(float) * d, s) != cudaSuccess || cudaMemsetAsync(dbeta, 0, sizeof(float) * d, s) != cudaSuccess)
        return HOISDF_E_SHAPE;
    }
    const dim3 grid(static_cast<unsigned>(ceil_div(d, 32)), static_cast<unsigned>(chunks));
    HOISDF_LAUNCH(layernorm_bwd_cols_kernel, grid, 256, s, h, dy, static_cast<const float*>(stats), rows, static_cast<int>(d),
                  dgamma, dbeta, accumulate ? 1 : 0);
  }
  return launch_status();
}

HOISDF_API int hoisdf_softmax_rows_fwd(const float* s, int64_t lds, int64_t rows, int64_t cols, int64_t valid,
                                       const uint8_t* mask, int64_t mask_rows, float* p, int64_t ldp, void* stream) {
  if (s == nullptr || p == nullptr) return HOISDF_E_NULL;
  if (rows <= 0 || cols <= 0 || cols > 0x7fffffffLL || valid <= 0 || valid > cols || lds < cols || ldp < cols)
    return HOISDF_E_SHAPE;
  if (mask != nullptr && mask_rows <= 0) return HOISDF_E_SHAPE;
  HOISDF_LAUNCH(softmax_rows_fwd_kernel, static_cast<unsigned>(ceil_div(rows, 8)), 256, static_cast<cudaStream_t>(stream), s,
                lds, rows, static_cast<int>(cols), static_cast<int>(valid), mask, mask_rows, p, ldp);
  return launch_status();
}

HOISDF_API int hoisdf_softmax_rows_bwd(const float* p, int64_t ldp, const float* dp, int64_t lddp, int64_t rows, int64_t cols,
                                       float* ds, int64_t ldds, void* stream) {
  if (p == nullptr || dp == nullptr || ds == nullptr) return HOISDF_E_NULL;
  if (rows <= 0 || cols <= 0 || cols > 0x7fffffffLL || ldp < cols || lddp < cols || ldds < cols) return HOISDF_E_SHAPE;
  HOISDF_LAUNCH(softmax_rows_bwd_kernel, static_cast<unsigned>(ceil_div(rows, 8)), 256, static_cast<cudaStream_t>(stream), p,
                ldp, dp, lddp, rows, static_cast<int>(cols), ds, ldds);
  return launch_status();
}

HOISDF_API int hoisdf_softmax_dropout_rows_fwd(const float* s, int64_t lds, int64_t rows, int64_t cols, int64_t valid,
                                               const uint8_t* mask, int64_t mask_rows, float* p, int64_t ldp, float* pd,
                                               int64_t ldpd, float p_drop, uint64_t seed, void* stream) {
  if (s == nullptr || pd == nullptr) return HOISDF_E_NULL;
  if (rows <= 0 || cols <= 0 || cols > 0x7fffffffLL || valid <= 0 || valid > cols || lds < cols || ldpd < cols ||
      (p != nullptr && ldp < cols) || !(p_drop >= 0.f && p_drop < 1.f))
    return HOISDF_E_SHAPE;
  if (mask != nullptr && mask_rows <= 0) return HOISDF_E_SHAPE;
  HOISDF_LAUNCH(softmax_dropout_rows_fwd_kernel, static_cast<unsigned>(ceil_div(rows, 8)), 256, static_cast<cudaStream_t>(stream),
                s, lds, rows, static_cast<int>(cols), static_cast<int>(valid), mask, mask_rows, p, ldp, pd, ldpd, p_drop, seed);
  return launch_status();
}

HOISDF_API int hoisdf_softmax_dropout_rows_bwd(const float* p, int64_t ldp, const float* dpd, int64_t lddp, int64_t rows,
                                               int64_t cols, float* ds, int64_t ldds, float p_drop, uint64_t seed, void* stream) {
  if (p == nullptr || dpd == nullptr || ds == nullptr) return HOISDF_E_NULL;
  if (rows <= 0 || cols <= 0 || cols > 0x7fffffffLL || ldp < cols || lddp < cols || ldds < cols || !(p_drop >= 0.f && p_drop < 1.f))
    return HOISDF_E_SHAPE;
  HOISDF_LAUNCH(softmax_dropout_rows_bwd_kernel, static_cast<unsigned>(ceil_div(rows, 8)), 256, static_cast<cudaStream_t>(stream),
                p, ldp, dpd, lddp, rows, static_cast<int>(cols), ds, ldds, p_drop, seed);
  return launch_status();
}

HOISDF_API int hoisdf_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, int64_t step, void* stream) {
  if (param == nullptr || grad == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr) return HOISDF_E_NULL;
  if (n <= 0 || step < 1 || !(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f)) return HOISDF_E_SHAPE;
  const double b1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
  const double b2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
  HOISDF_LAUNCH(adamw_kernel, static_cast<unsigned>(ceil_div(n, 256)), 256, static_cast<cudaStream_t>(stream), param, grad,
                exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, static_cast<float>(lr / b1),
                static_cast<float>(sqrt(b2)));
  return launch_status();
}

HOISDF_API int hoisdf_vote_loss_bwd(const float* points, const float* off, const float* cls, const float* joint_gt,
                                    int64_t layers, int64_t batch, int64_t p, float cls_dist, float g_joint_3d, float g_cls,
                                    float g_all_joint_3d, float* d_off, float* d_cls, float* npos_ws, void* stream) {
  if (points == nullptr || off == nullptr || cls == nullptr || joint_gt == nullptr || d_off == nullptr || d_cls == nullptr ||
      npos_ws == nullptr)
    return HOISDF_E_NULL;
  if (layers <= 0 || batch <= 0 || p <= 0 || layers * batch > 0x7fffffffLL) return HOISDF_E_SHAPE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  HOISDF_LAUNCH(vote_count_pos_kernel, 1, 256, s, points, joint_gt, batch, p, cls_dist, npos_ws);
  HOISDF_LAUNCH(vote_loss_bwd_kernel, static_cast<unsigned>(layers * batch), 256, s, points, off, cls, joint_gt, layers, batch,
                p, cls_dist, g_joint_3d, g_cls, g_all_joint_3d, static_cast<const float*>(npos_ws), d_off, d_cls);
  return launch_status();
}

HOISDF_API int64_t hoisdf_tokens_bwd_workspace_bytes(int64_t batch, int64_t p) {
  if (batch <= 0 || p <= 0) return 0;
  return ceil_div(batch * p, 8) * static_cast<int64_t>(sizeof(float));
}

HOISDF_API int hoisdf_tokens_bwd(const float* d_tokens, int64_t s_total, int64_t t0, const float* fea, int64_t ld_fea,
                                 const float* sdf, const float* beta, int64_t batch, int64_t p, float* d_fea, int64_t ld_dfea,
                                 float* d_sdf, float* d_beta, int32_t accumulate_beta, void* workspace, int64_t workspace_bytes,
                                 void* stream) {
  if (d_tokens == nullptr || fea == nullptr || sdf == nullptr || beta == nullptr || d_fea == nullptr || d_beta == nullptr ||
      workspace == nullptr)
    return HOISDF_E_NULL;
  if (batch <= 0 || p <= 0 || t0 < 0 || t0 + p > s_total || ld_fea < 223 || ld_dfea < 223) return HOISDF_E_SHAPE;
  if (workspace_bytes < hoisdf_tokens_bwd_workspace_bytes(batch, p)) return HOISDF_E_SHAPE;
  const int64_t blocks = ceil_div(batch * p, 8);
  if (blocks > 0x7fffffffLL) return HOISDF_E_SHAPE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(workspace);
  HOISDF_LAUNCH(tokens_bwd_kernel, static_cast<unsigned>(blocks), 256, s, d_tokens, s_total, t0, fea, ld_fea, sdf, beta, batch, p,
                d_fea, ld_dfea, d_sdf, partial);
  HOISDF_LAUNCH(sum_partials_kernel, 1, 256, s, static_cast<const float*>(partial), blocks, d_beta, accumulate_beta ? 1 : 0);
  return launch_status();
}
