// Multi-head attention core (upstream nn.MultiheadAttention inside common/nets/transformer.py:294,378,383),
// head_dim = 64, fp32, streaming softmax: the Lq x Lk score matrix upstream materialises (need_weights=True
// forces the unfused math path there) never exists here.
//
// flash kernel (encoder self-attention, Lq = Lk = S up to a few thousand):
//   CTA = 128 queries x one head x one sample, 256 threads, KV tiles of 64 keys.
//   thread (ty = tid/16, tx = tid%16) owns query rows {ty + 16 i, i<8}; for S = Q.K^T it owns keys {tx + 16 j, j<4},
//   for O = P.V it owns head dims {4 tx .. 4 tx + 3}.  All shared-memory reads are LDS.128 along the
//   contiguous (d or key) axis of row-major tiles with a 68-float pitch -> conflict-free, and the tile
//   loads are straight coalesced 128-bit copies.
// small kernel (decoder: 17 queries): one CTA per (query, head, sample), scores staged in shared memory,
//   supports the dense boolean mask (common/utils/misc.py:11-31) and the "keys >= kv_valid" memory mask
//   (misc.py:42-47).
#include "common.cuh"

namespace hoisdf {

constexpr int HD = 64;       // head dim
constexpr int BQ = 128;      // queries per CTA
constexpr int BKV = 64;      // keys per tile
constexpr int PITCH = 68;    // smem row pitch (floats)
constexpr int kFlashSmem = (BQ * PITCH * 2 + BKV * PITCH * 2) * 4;

struct AttnParams {
  const float* __restrict__ q;
  const float* __restrict__ k;
  const float* __restrict__ v;
  float* __restrict__ out;
  const uint8_t* __restrict__ mask;
  int64_t ldq, ldk, ldo;
  int lq, lk, kv_valid;
};

__global__ void __launch_bounds__(256, 2) attention_flash_kernel(const AttnParams p) {
  HOISDF_DYNAMIC_SMEM(float, smem);
  float* Qs = smem;                    // [BQ][PITCH]
  float* Ps = Qs + BQ * PITCH;         // [BQ][PITCH]
  float* Ks = Ps + BQ * PITCH;         // [BKV][PITCH]
  float* Vs = Ks + BKV * PITCH;        // [BKV][PITCH]

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * BQ;
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z;
  const float* qg = p.q + (b * p.lq) * p.ldq + h * HD;
  const float* kg = p.k + (b * p.lk) * p.ldk + h * HD;
  const float* vg = p.v + (b * p.lk) * p.ldk + h * HD;

  // Q tile (pre-scaled by 1/sqrt(64) = 0.125, exact) : 128 x 64 floats = 2048 float4, 8 per thread
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int f = tid + 256 * j;
    const int r = f >> 4, c4 = (f & 15) * 4;
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < p.lq) val = __ldg(reinterpret_cast<const float4*>(qg + static_cast<int64_t>(q0 + r) * p.ldq + c4));
    val.x *= 0.125f; val.y *= 0.125f; val.z *= 0.125f; val.w *= 0.125f;
    *reinterpret_cast<float4*>(Qs + r * PITCH + c4) = val;
  }

  float o[8][4];
  float m_run[8], l_run[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }

  const int kend = min(p.lk, p.kv_valid);
  const int ntiles = (kend + BKV - 1) / BKV;
  for (int t = 0; t < ntiles; ++t) {
    const int k0 = t * BKV;
    __syncthreads();  // previous tile's Ks/Vs/Ps fully consumed (also orders the Q stores on t == 0)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int f = tid + 256 * j;
      const int r = f >> 4, c4 = (f & 15) * 4;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + r < kend) {
        kv = __ldg(reinterpret_cast<const float4*>(kg + static_cast<int64_t>(k0 + r) * p.ldk + c4));
        vv = __ldg(reinterpret_cast<const float4*>(vg + static_cast<int64_t>(k0 + r) * p.ldk + c4));
      }
      *reinterpret_cast<float4*>(Ks + r * PITCH + c4) = kv;
      *reinterpret_cast<float4*>(Vs + r * PITCH + c4) = vv;
    }
    __syncthreads();

    // S = Q . K^T for rows {ty + 16 i} x keys {tx + 16 j}
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < HD; d += 4) {
      float4 kk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) kk[j] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * PITCH + d);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 qq = *reinterpret_cast<const float4*>(Qs + (ty + 16 * i) * PITCH + d);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[i][j] = fmaf(qq.x, kk[j].x, s[i][j]);
          s[i][j] = fmaf(qq.y, kk[j].y, s[i][j]);
          s[i][j] = fmaf(qq.z, kk[j].z, s[i][j]);
          s[i][j] = fmaf(qq.w, kk[j].w, s[i][j]);
        }
      }
    }
    // mask the key tail
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (k0 + tx + 16 * j >= kend) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i][j] = -INFINITY;
      }
    }
    // online softmax; the 16 lanes that share ty hold one full row of the tile
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float mx = fmaxf(fmaxf(s[i][0], s[i][1]), fmaxf(s[i][2], s[i][3]));
#pragma unroll
      for (int w = 8; w > 0; w >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, w));
      const float m_new = fmaxf(m_run[i], mx);  // finite: every tile holds at least one valid key
      const float corr = expf(m_run[i] - m_new);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float e = expf(s[i][j] - m_new);
        s[i][j] = e;
        rs += e;
      }
#pragma unroll
      for (int w = 8; w > 0; w >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, w);
      l_run[i] = l_run[i] * corr + rs;
      m_run[i] = m_new;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[i][j] *= corr;
        Ps[(ty + 16 * i) * PITCH + tx + 16 * j] = s[i][j];
      }
    }
    __syncthreads();
    // O += P . V for rows {ty + 16 i} x dims {4 tx ..}
#pragma unroll 4
    for (int kk = 0; kk < BKV; kk += 4) {
      float4 vv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) vv[j] = *reinterpret_cast<const float4*>(Vs + (kk + j) * PITCH + tx * 4);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 pp = *reinterpret_cast<const float4*>(Ps + (ty + 16 * i) * PITCH + kk);
        o[i][0] = fmaf(pp.x, vv[0].x, o[i][0]); o[i][1] = fmaf(pp.x, vv[0].y, o[i][1]);
        o[i][2] = fmaf(pp.x, vv[0].z, o[i][2]); o[i][3] = fmaf(pp.x, vv[0].w, o[i][3]);
        o[i][0] = fmaf(pp.y, vv[1].x, o[i][0]); o[i][1] = fmaf(pp.y, vv[1].y, o[i][1]);
        o[i][2] = fmaf(pp.y, vv[1].z, o[i][2]); o[i][3] = fmaf(pp.y, vv[1].w, o[i][3]);
        o[i][0] = fmaf(pp.z, vv[2].x, o[i][0]); o[i][1] = fmaf(pp.z, vv[2].y, o[i][1]);
        o[i][2] = fmaf(pp.z, vv[2].z, o[i][2]); o[i][3] = fmaf(pp.z, vv[2].w, o[i][3]);
        o[i][0] = fmaf(pp.w, vv[3].x, o[i][0]); o[i][1] = fmaf(pp.w, vv[3].y, o[i][1]);
        o[i][2] = fmaf(pp.w, vv[3].z, o[i][2]); o[i][3] = fmaf(pp.w, vv[3].w, o[i][3]);
      }
    }
  }

  float* og = p.out + (b * p.lq) * p.ldo + h * HD;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = q0 + ty + 16 * i;
    if (r < p.lq) {
      const float inv = __fdiv_rn(1.f, l_run[i]);
      *reinterpret_cast<float4*>(og + static_cast<int64_t>(r) * p.ldo + tx * 4) =
          make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
    }
  }
}

// one CTA (128 threads) per (query, head, sample); scores in dynamic smem (lk floats)
__global__ void __launch_bounds__(128) attention_small_kernel(const AttnParams p) {
  HOISDF_DYNAMIC_SMEM(float, sc);
  __shared__ float qs[HD];
  __shared__ float red[4];
  __shared__ float part[2][HD];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int qi = blockIdx.x, h = blockIdx.y;
  const int64_t b = blockIdx.z;
  const float* qg = p.q + (b * p.lq + qi) * p.ldq + h * HD;
  const float* kg = p.k + (b * p.lk) * p.ldk + h * HD;
  const float* vg = p.v + (b * p.lk) * p.ldk + h * HD;
  if (tid < HD) qs[tid] = qg[tid] * 0.125f;
  __syncthreads();
  const int kend = min(p.lk, p.kv_valid);
  float mx = -INFINITY;
  for (int j = tid; j < kend; j += 128) {
    float s;
    if (p.mask != nullptr && p.mask[static_cast<int64_t>(qi) * p.lk + j]) {
      s = -INFINITY;
    } else {
      const float* kr = kg + static_cast<int64_t>(j) * p.ldk;
      s = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 kk = __ldg(reinterpret_cast<const float4*>(kr + d));
        s = fmaf(qs[d], kk.x, s); s = fmaf(qs[d + 1], kk.y, s);
        s = fmaf(qs[d + 2], kk.z, s); s = fmaf(qs[d + 3], kk.w, s);
      }
    }
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if (lane == 0) red[wid] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int j = tid; j < kend; j += 128) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[wid] = sum;
  __syncthreads();
  sum = red[0] + red[1] + red[2] + red[3];
  // out[d] = sum_j p_j V[j][d]; two halves of the key range, 64 dims each
  const int d = tid & 63, half = tid >> 6;
  float acc = 0.f;
  for (int j = half; j < kend; j += 2) acc = fmaf(sc[j], __ldg(vg + static_cast<int64_t>(j) * p.ldk + d), acc);
  part[half][d] = acc;
  __syncthreads();
  if (tid < HD) {
    p.out[(b * p.lq + qi) * p.ldo + h * HD + tid] = __fdiv_rn(part[0][tid] + part[1][tid], sum);
  }
}

int64_t attention_tc_workspace_bytes(int64_t batch, int64_t heads, int64_t lq, int64_t lk);   // attention_tc.cu
int launch_attention_tc(const float* q, int64_t ldq, const float* k, const float* v, int64_t ldk, float* out,
                        int64_t ldo, int64_t batch, int64_t heads, int64_t lq, int64_t lk, int64_t kv_valid,
                        void* workspace, cudaStream_t s, uint16_t* out_hi = nullptr, uint16_t* out_lo = nullptr,
                        float p_drop = 0.f, uint64_t seed = 0, float* lse = nullptr);

int64_t attention_bwd_tc_workspace_bytes(int64_t batch, int64_t heads, int64_t lq, int64_t lk);
int launch_attention_bwd_tc(const float* q, int64_t ldq, const float* k, const float* v, int64_t ldk, const float* out,
                            const float* dout, int64_t ldo, const float* lse, float* dq, float* dk, float* dv, int64_t ldg,
                            int64_t batch, int64_t heads, int64_t lq, int64_t lk, int64_t kv_valid, float p_drop,
                            uint64_t seed, void* workspace, cudaStream_t s);

}  // namespace hoisdf

using namespace hoisdf;

HOISDF_API int64_t hoisdf_attention_workspace_bytes(int64_t batch, int64_t heads, int64_t lq, int64_t lk) {
  if (batch <= 0 || heads <= 0 || lq <= 0 || lk <= 0) return 0;
  return attention_tc_workspace_bytes(batch, heads, lq, lk);
}

HOISDF_API int hoisdf_attention_fwd(const float* q, int64_t ldq, const float* k, const float* v, int64_t ldk,
                                    float* out, int64_t ldo, int64_t batch, int64_t heads, int64_t lq, int64_t lk,
                                    int64_t kv_valid, const uint8_t* mask, void* workspace, int64_t workspace_bytes,
                                    void* stream) {
  if (q == nullptr || k == nullptr || v == nullptr || out == nullptr) return HOISDF_E_NULL;
  if (batch <= 0 || batch > 65535 || heads <= 0 || heads > 65535 || lq <= 0 || lk <= 0 || kv_valid <= 0 ||
      lq > (1 << 24) || lk > (1 << 24))
    return HOISDF_E_SHAPE;
  if (ldq < heads * HD || ldk < heads * HD || ldo < heads * HD) return HOISDF_E_SHAPE;
  if ((ldq & 3) || (ldk & 3) || (ldo & 3) || !aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(out))
    return HOISDF_E_ALIGN;
  AttnParams p;
  p.q = q; p.k = k; p.v = v; p.out = out; p.mask = mask;
  p.ldq = ldq; p.ldk = ldk; p.ldo = ldo;
  p.lq = static_cast<int>(lq); p.lk = static_cast<int>(lk);
  p.kv_valid = static_cast<int>(kv_valid < lk ? kv_valid : lk);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (workspace != nullptr && mask == nullptr) {
    if (workspace_bytes < attention_tc_workspace_bytes(batch, heads, lq, lk)) return HOISDF_E_SHAPE;
    if (!aligned16(workspace)) return HOISDF_E_ALIGN;
    return launch_attention_tc(q, ldq, k, v, ldk, out, ldo, batch, heads, lq, lk, p.kv_valid, workspace, s);
  }
  if (mask != nullptr || lq <= 32) {
    if (lq > 65535 * 32 || lk > 12000) return HOISDF_E_UNSUPPORTED;  // scores must fit in shared memory
    dim3 grid(static_cast<unsigned>(lq), static_cast<unsigned>(heads), static_cast<unsigned>(batch));
    HOISDF_LAUNCH_SMEM(attention_small_kernel, grid, 128, static_cast<size_t>(lk) * sizeof(float), s, p);
  } else {
#ifndef HOISDF_EMULATE
    {   // per call: the attribute belongs to the CURRENT device's instance of the kernel (a process may drive several GPUs)
      cudaError_t e = cudaFuncSetAttribute(attention_flash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kFlashSmem);
      if (e != cudaSuccess) return static_cast<int>(e);
    }
#endif
    dim3 grid(static_cast<unsigned>(ceil_div(lq, BQ)), static_cast<unsigned>(heads), static_cast<unsigned>(batch));
    HOISDF_LAUNCH_SMEM(attention_flash_kernel, grid, 256, kFlashSmem, s, p);
  }
  return launch_status();
}

HOISDF_API int hoisdf_attention_train_fwd(const float* q, int64_t ldq, const float* k, const float* v, int64_t ldk,
                                          float* out, int64_t ldo, float* lse, int64_t batch, int64_t heads, int64_t lq,
                                          int64_t lk, int64_t kv_valid, float p_drop, uint64_t seed, void* workspace,
                                          int64_t workspace_bytes, void* stream) {
  if (q == nullptr || k == nullptr || v == nullptr || out == nullptr || workspace == nullptr) return HOISDF_E_NULL;
  if (batch <= 0 || batch > 65535 || heads <= 0 || heads > 65535 || lq <= 0 || lk <= 0 || kv_valid <= 0 ||
      lq > (1 << 24) || lk > (1 << 24) || !(p_drop >= 0.f && p_drop < 1.f))
    return HOISDF_E_SHAPE;
  if (ldq < heads * HD || ldk < heads * HD || ldo < heads * HD) return HOISDF_E_SHAPE;
  if ((ldq & 3) || (ldk & 3) || (ldo & 3) || !aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(out) ||
      !aligned16(workspace))
    return HOISDF_E_ALIGN;
  if (workspace_bytes < attention_tc_workspace_bytes(batch, heads, lq, lk)) return HOISDF_E_SHAPE;
  if ((kv_valid < lk ? kv_valid : lk) < 128) return HOISDF_E_UNSUPPORTED;       // the 128-key tensor-core kernel only
  return launch_attention_tc(q, ldq, k, v, ldk, out, ldo, batch, heads, lq, lk, kv_valid < lk ? kv_valid : lk, workspace,
                             static_cast<cudaStream_t>(stream), nullptr, nullptr, p_drop, seed, lse);
}

HOISDF_API int64_t hoisdf_attention_bwd_workspace_bytes(int64_t batch, int64_t heads, int64_t lq, int64_t lk) {
  if (batch <= 0 || heads <= 0 || lq <= 0 || lk <= 0) return 0;
  return attention_bwd_tc_workspace_bytes(batch, heads, lq, lk);
}

HOISDF_API int hoisdf_attention_bwd(const float* q, int64_t ldq, const float* k, const float* v, int64_t ldk, const float* out,
                                    const float* dout, int64_t ldo, const float* lse, float* dq, float* dk, float* dv,
                                    int64_t ldg, int64_t batch, int64_t heads, int64_t lq, int64_t lk, int64_t kv_valid,
                                    float p_drop, uint64_t seed, void* workspace, int64_t workspace_bytes, void* stream) {
  if (q == nullptr || k == nullptr || v == nullptr || out == nullptr || dout == nullptr || lse == nullptr || dq == nullptr ||
      dk == nullptr || dv == nullptr || workspace == nullptr)
    return HOISDF_E_NULL;
  if (batch <= 0 || batch > 65535 || heads <= 0 || heads > 65535 || lq <= 0 || lk <= 0 || kv_valid <= 0 ||
      lq > (1 << 24) || lk > (1 << 24) || !(p_drop >= 0.f && p_drop < 1.f))
    return HOISDF_E_SHAPE;
  if (ldq < heads * HD || ldk < heads * HD || ldo < heads * HD || ldg < heads * HD) return HOISDF_E_SHAPE;
  if ((ldq & 3) || (ldk & 3) || (ldo & 3) || (ldg & 3) || !aligned16(q) || !aligned16(k) || !aligned16(v) ||
      !aligned16(out) || !aligned16(dout) || !aligned16(dq) || !aligned16(dk) || !aligned16(dv) || !aligned16(workspace))
    return HOISDF_E_ALIGN;
  if (workspace_bytes < attention_bwd_tc_workspace_bytes(batch, heads, lq, lk)) return HOISDF_E_SHAPE;
  return launch_attention_bwd_tc(q, ldq, k, v, ldk, out, dout, ldo, lse, dq, dk, dv, ldg, batch, heads, lq, lk,
                                 kv_valid < lk ? kv_valid : lk, p_drop, seed, workspace, static_cast<cudaStream_t>(stream));
}

HOISDF_API int hoisdf_attention_split_fwd(const float* q, int64_t ldq, const float* k, const float* v, int64_t ldk,
                                          uint16_t* out_hi, uint16_t* out_lo, int64_t ldo, int64_t batch, int64_t heads,
                                          int64_t lq, int64_t lk, int64_t kv_valid, void* workspace,
                                          int64_t workspace_bytes, void* stream) {
  if (q == nullptr || k == nullptr || v == nullptr || out_hi == nullptr || out_lo == nullptr || workspace == nullptr)
    return HOISDF_E_NULL;
  if (batch <= 0 || batch > 65535 || heads <= 0 || heads > 65535 || lq <= 0 || lk <= 0 || kv_valid <= 0 ||
      lq > (1 << 24) || lk > (1 << 24))
    return HOISDF_E_SHAPE;
  if (ldq < heads * HD || ldk < heads * HD || ldo < heads * HD) return HOISDF_E_SHAPE;
  if ((ldq & 3) || (ldk & 3) || (ldo & 7) || !aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(out_hi) ||
      !aligned16(out_lo) || !aligned16(workspace))
    return HOISDF_E_ALIGN;
  if (workspace_bytes < attention_tc_workspace_bytes(batch, heads, lq, lk)) return HOISDF_E_SHAPE;
  return launch_attention_tc(q, ldq, k, v, ldk, nullptr, ldo, batch, heads, lq, lk, kv_valid < lk ? kv_valid : lk,
                             workspace, static_cast<cudaStream_t>(stream), out_hi, out_lo);
}
