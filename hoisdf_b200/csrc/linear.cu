// nn.Linear (+bias, ReLU, residual) over rows: Y = act(X . W^T + b) -- fp32 FMA path.
//
// Replaces every cuBLAS sgemm call site of the upstream hot path (common/nets/layer.py:192-201,
// common/nets/sdf_net.py:95-107, common/nets/transformer.py:294-299).  This is the bit-faithful fp32
// formulation (no reduced-precision operand rounding): the candidate SDF path needs fp32-grade values
// because the near-surface selection (main/model.py:345) is a top-k on them.
//
// Tiling: CTA tile BM x BN, K step 16, 256 threads, each thread an (8 x TN) register tile split into
// 4-wide quads so every shared-memory read is one conflict-free LDS.128; operands are staged through
// registers (global float4 along K -> transposed k-major smem), double buffered, one __syncthreads per
// K step.  Both operands are K-contiguous ("NT" GEMM), so a warp reads 8 rows x 64 contiguous bytes.
#include <cstdlib>

#include "common.cuh"

namespace hoisdf {

struct RowAddr {
  int64_t ld, rows_per_batch, batch_stride;
  __device__ __forceinline__ int64_t offset(int64_t r) const {
    if (rows_per_batch <= 0) return r * ld;
    const int64_t b = r / rows_per_batch;
    return b * batch_stride + (r - b * rows_per_batch) * ld;
  }
};

struct LinearParams {
  const float* __restrict__ x;
  const float* __restrict__ w;
  const float* __restrict__ bias;
  const float* __restrict__ residual;
  float* __restrict__ y;
  RowAddr xa, ya;
  int64_t ldw;
  int64_t m;
  int n, k, act;
};

constexpr int BK = 16;

template <int BM, int BN>
__global__ void __launch_bounds__(256, 2) linear_fp32_kernel(const LinearParams p) {
  constexpr int TM = 8;
  constexpr int TN = BN / 16;       // 8 (BN=128) or 4 (BN=64)
  constexpr int LDA = BM + 4;
  constexpr int LDB = BN + 4;
  constexpr int A_VEC = BM * BK / 4 / 256;  // float4 per thread per stage
  constexpr int B_VEC = (BN * BK / 4 + 255) / 256;
  static_assert(BM == 128, "row tile is fixed at 128");

  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][LDB];

  const int tid = threadIdx.x;
  // 1-D grid, N tile fastest: the CTAs that share one 128-row X tile are co-resident, so X streams from
  // HBM once and the re-reads hit L2 (W is small and always L2-resident).
  const int n_tiles = (p.n + BN - 1) / BN;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x / n_tiles) * BM;
  const int col0 = static_cast<int>(blockIdx.x % n_tiles) * BN;

  // global -> register staging map: float4 f = tid + 256*j ; kq = f % 4 (k quad), r = f / 4 (tile row)
  const float* a_ptr[A_VEC];
  bool a_ok[A_VEC];
#pragma unroll
  for (int j = 0; j < A_VEC; ++j) {
    const int f = tid + 256 * j;
    const int64_t r = row0 + (f >> 2);
    a_ok[j] = r < p.m;
    a_ptr[j] = p.x + (a_ok[j] ? p.xa.offset(r) : 0) + (f & 3) * 4;
  }
  const float* b_ptr[B_VEC];
  bool b_ok[B_VEC];
#pragma unroll
  for (int j = 0; j < B_VEC; ++j) {
    const int f = tid + 256 * j;
    const int r = col0 + (f >> 2);
    b_ok[j] = (f < BN * BK / 4) && (r < p.n);
    b_ptr[j] = p.w + (b_ok[j] ? static_cast<int64_t>(r) * p.ldw : 0) + (f & 3) * 4;
  }

  float4 a_reg[A_VEC], b_reg[B_VEC];
  const int ktiles = (p.k + BK - 1) / BK;

  auto load_global = [&](int kt) {
    const int kbase = kt * BK;
#pragma unroll
    for (int j = 0; j < A_VEC; ++j) {
      const int kk = kbase + ((tid + 256 * j) & 3) * 4;
      a_reg[j] = (a_ok[j] && kk < p.k) ? __ldg(reinterpret_cast<const float4*>(a_ptr[j] + kbase))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < B_VEC; ++j) {
      const int kk = kbase + ((tid + 256 * j) & 3) * 4;
      b_reg[j] = (b_ok[j] && kk < p.k) ? __ldg(reinterpret_cast<const float4*>(b_ptr[j] + kbase))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int j = 0; j < A_VEC; ++j) {
      const int f = tid + 256 * j;
      const int r = f >> 2, kq = (f & 3) * 4;
      As[buf][kq + 0][r] = a_reg[j].x;
      As[buf][kq + 1][r] = a_reg[j].y;
      As[buf][kq + 2][r] = a_reg[j].z;
      As[buf][kq + 3][r] = a_reg[j].w;
    }
#pragma unroll
    for (int j = 0; j < B_VEC; ++j) {
      const int f = tid + 256 * j;
      if (f < BN * BK / 4) {
        const int r = f >> 2, kq = (f & 3) * 4;
        Bs[buf][kq + 0][r] = b_reg[j].x;
        Bs[buf][kq + 1][r] = b_reg[j].y;
        Bs[buf][kq + 2][r] = b_reg[j].z;
        Bs[buf][kq + 3][r] = b_reg[j].w;
      }
    }
  };

  // compute map: ty = tid / 16 owns rows {ty*4..+3, 64+ty*4..+3}; tx = tid % 16 owns column quads
  const int tx = tid & 15, ty = tid >> 4;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load_global(0);
  store_smem(0);
  __syncthreads();

  for (int kt = 0; kt < ktiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ktiles) load_global(kt + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int q = 0; q < TN / 4; ++q) {
        const float4 bq = *reinterpret_cast<const float4*>(&Bs[buf][kk][q * 64 + tx * 4]);
        b[q * 4 + 0] = bq.x; b[q * 4 + 1] = bq.y; b[q * 4 + 2] = bq.z; b[q * 4 + 3] = bq.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < ktiles) {
      store_smem(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: bias, residual, activation
  const bool vec_ok = ((p.ya.ld & 3) == 0) && ((p.ya.batch_stride & 3) == 0) && aligned16(p.y) &&
                      (p.residual == nullptr || aligned16(p.residual));
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t r = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= p.m) continue;
    const int64_t yo = p.ya.offset(r);
#pragma unroll
    for (int q = 0; q < TN / 4; ++q) {
      const int c = col0 + q * 64 + tx * 4;
      if (c >= p.n) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = acc[i][q * 4 + j];
        if (p.bias != nullptr && c + j < p.n) v[j] += __ldg(p.bias + c + j);
      }
      if (vec_ok && c + 3 < p.n) {
        if (p.residual != nullptr) {
          const float4 rr = *reinterpret_cast<const float4*>(p.residual + yo + c);
          v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
        }
        if (p.act == HOISDF_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        *reinterpret_cast<float4*>(p.y + yo + c) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (c + j < p.n) {
            float o = v[j];
            if (p.residual != nullptr) o += p.residual[yo + c + j];
            if (p.act == HOISDF_ACT_RELU) o = fmaxf(o, 0.f);
            p.y[yo + c + j] = o;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// weight-norm fold / column permutation
// ---------------------------------------------------------------------------------------------------
__global__ void fold_weight_norm_kernel(const float* __restrict__ g, const float* __restrict__ v, int64_t rows,
                                        int64_t cols, float* __restrict__ out, int64_t ld_out,
                                        const int32_t* __restrict__ src_col, int64_t cols_out) {
  const int64_t r = blockIdx.x;
  if (r >= rows) return;
  const float* vr = v + r * cols;
  float scale = 1.f;
  if (g != nullptr) {
    // ||v||_2 of the row, pairwise (warp tree) summation in fp32 like ATen's norm kernel up to ordering
    float ss = 0.f;
    for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) ss = fmaf(vr[c], vr[c], ss);
    __shared__ float red[32];
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
      t = warp_sum(t);
      if (threadIdx.x == 0) red[0] = t;
    }
    __syncthreads();
    scale = __fdiv_rn(g[r], sqrtf(red[0]));
  }
  for (int64_t c = threadIdx.x; c < cols_out; c += blockDim.x) {
    const int64_t s = src_col ? src_col[c] : c;
    out[r * ld_out + c] = (s >= 0 && s < cols) ? vr[s] * scale : 0.f;
  }
}

int launch_linear_tf32x3(const hoisdf_linear_args* a, cudaStream_t s);      // linear_tc.cu  (1-SM MMA)
int launch_linear_tf32x3_2sm(const hoisdf_linear_args* a, cudaStream_t s);  // linear_tc2.cu (2-SM MMA)

}  // namespace hoisdf

using namespace hoisdf;

HOISDF_API int hoisdf_linear_fwd(const hoisdf_linear_args* a, void* stream) {
  if (a == nullptr || a->x == nullptr || a->w == nullptr || a->y == nullptr) return HOISDF_E_NULL;
  if (a->m < 0 || a->n <= 0 || a->k <= 0 || a->n > (1 << 20) || a->k > (1 << 20)) return HOISDF_E_SHAPE;
  if (a->m == 0) return HOISDF_OK;
  if ((a->k & 3) || (a->ldx & 3) || (a->ldw & 3) || (a->x_batch_stride & 3)) return HOISDF_E_ALIGN;
  if (!aligned16(a->x) || !aligned16(a->w)) return HOISDF_E_ALIGN;
  if (a->ldx < a->k || a->ldw < a->k || a->ldy < a->n) return HOISDF_E_SHAPE;
  if (a->w_lo != nullptr && a->y_rows_per_batch <= 0) {
    if (!aligned16(a->w_lo)) return HOISDF_E_ALIGN;
    // the 2-SM (cta_group::2) variant is correct but measured slower (131 vs 200 TFLOP/s at K = 512): the 1-SM
    // mainloop already runs at ~75 % of the TF32 pipe, so halving operand traffic buys nothing; opt-in for experiments
    static const bool use_2sm = [] { const char* e = getenv("HOISDF_TC_2SM"); return e != nullptr && e[0] == '1'; }();
    return use_2sm ? launch_linear_tf32x3_2sm(a, static_cast<cudaStream_t>(stream))
                   : launch_linear_tf32x3(a, static_cast<cudaStream_t>(stream));
  }
  LinearParams p;
  p.x = a->x; p.w = a->w; p.bias = a->bias; p.residual = a->residual; p.y = a->y;
  p.xa = {a->ldx, a->x_rows_per_batch, a->x_batch_stride};
  p.ya = {a->ldy, a->y_rows_per_batch, a->y_batch_stride};
  p.ldw = a->ldw; p.m = a->m; p.n = static_cast<int>(a->n); p.k = static_cast<int>(a->k); p.act = a->act;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t mt = ceil_div(a->m, 128);
  const int64_t nt = a->n > 64 ? ceil_div(a->n, 128) : 1;
  if (mt * nt > 0x7fffffffLL) return HOISDF_E_SHAPE;
  if (a->n > 64) {
    HOISDF_LAUNCH((linear_fp32_kernel<128, 128>), static_cast<unsigned>(mt * nt), 256, s, p);
  } else {
    HOISDF_LAUNCH((linear_fp32_kernel<128, 64>), static_cast<unsigned>(mt * nt), 256, s, p);
  }
  return launch_status();
}

HOISDF_API int hoisdf_fold_weight_norm(const float* g, const float* v, int64_t rows, int64_t cols, float* out,
                                       int64_t ld_out, const int32_t* src_col, int64_t cols_out, void* stream) {
  if (v == nullptr || out == nullptr) return HOISDF_E_NULL;
  if (rows <= 0 || cols <= 0 || cols_out <= 0 || ld_out < cols_out) return HOISDF_E_SHAPE;
  HOISDF_LAUNCH(fold_weight_norm_kernel, static_cast<unsigned>(rows), 256, static_cast<cudaStream_t>(stream), g, v, rows,
                cols, out, ld_out, src_col, cols_out);
  return launch_status();
}
