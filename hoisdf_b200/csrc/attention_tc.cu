// Encoder self-attention on the 5th-generation tensor cores (head_dim 64), streaming softmax, BF16x3 split.
//
// Replaces the S x S `nn.MultiheadAttention` math path of upstream common/nets/transformer.py:294 for the long
// sequences (S = P_h + P_o = 800 ... 4096 tokens).  fp32 operands are split into two bf16 terms each
// (x = hi + lo, |lo| <= 2^-9 |x|) and every product is evaluated as hi.hi + lo.hi + hi.lo with fp32 accumulation
// in TMEM: relative error ~2^-16 -- 30x better than one TF32 pass at the same shared-memory footprint -- which keeps
// the transformer outputs inside the 1e-3 parity bar with a wide margin (the top-k-critical SDF path does not use
// this kernel).
//
// Two kernels:
//   attn_split_kernel   fp32 q|k|v rows -> bf16 hi/lo arrays laid out per (sample, head): Q, K as (B*H*L, 64),
//                       V transposed as (B*H*64, Lk_pad) so that it is a K-major B operand.  q is pre-scaled by log2(e)/8 (scores in log2 units).
//   attention_tc_kernel one CTA per (128 queries, head, sample), 192 threads:
//       warp 0      TMA producer (Q once; K / V^T tiles of 64 keys through a 3-stage ring, 128B swizzle)
//       warp 1      tcgen05.mma issuer: S_t = Q.K_t^T (M128 N64 K16 x 12) into one of two TMEM score buffers,
//                   O += P_t.V_t (x 12) into the TMEM output accumulator
//       warps 2..9  softmax (two groups of four, one 128-query tile each): thread = query row (TMEM lane):
//                   tcgen05.ld the score row, online max / sum in fp32 (log2 domain), rescale O in TMEM
//                   (tcgen05.ld/st) only when the running max moved, write P as packed bf16 hi / lo pairs back into
//                   TMEM (tcgen05.st) where the P.V MMA reads it as its A operand -- P never touches shared memory,
//                   whose bandwidth is what bounds this kernel; epilogue O / l.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"

namespace hoisdf {

constexpr int AT_BQ = 128;                 // queries per softmax group (one MMA M tile)
constexpr int AT_GROUPS = 2;               // query tiles per CTA: two softmax warpgroups ping-pong on the tensor core
constexpr int AT_BK = 64;                  // keys per tile
constexpr int AT_D = 64;                   // head dim
constexpr int AT_STAGES = 4;
constexpr float AT_RESCALE_LOG2 = 8.f;     // lazy softmax rescaling threshold (log2 units), see the softmax warps
constexpr int AT_Q_BYTES = AT_BQ * AT_D * 2;      // 16 KB (one bf16 term)
constexpr int AT_K_BYTES = AT_BK * AT_D * 2;      // 8 KB
constexpr int AT_P_BYTES = AT_BQ * AT_BK * 2;     // 16 KB
constexpr int AT_STAGE_BYTES = 4 * AT_K_BYTES;    // K_hi K_lo Vt_hi Vt_lo
constexpr int AT_SMEM_BYTES = AT_GROUPS * 2 * AT_Q_BYTES + AT_STAGES * AT_STAGE_BYTES + 1024 + 256;
constexpr int AT_THREADS = 64 + AT_GROUPS * 128;
constexpr uint32_t AT_TMEM_COLS = 512;     // group g: S0 [128g, +64) S1 [128g + 64, +64); O_g [256 + 64g, +64);
                                           // P_g (bf16 pairs, A operand of P.V): hi [384 + 64g, +32), lo [+32, +64)
constexpr uint32_t kAtSpinLimit = 1u << 27;

struct AttnTcParams {
  float* __restrict__ out;           // fp32 output rows, or nullptr when the split-half planes are given
  uint16_t* __restrict__ out_hi;     // split-half output (hi = fp16(x), lo = fp16((x - hi) * 2^11)): what the FP16x3
  uint16_t* __restrict__ out_lo;     // out-projection GEMM reads, saving a conversion pass
  int64_t ldo;                       // row pitch of whichever output is used (floats / halfs)
  int lq, lk, kv_valid, heads;
  // training only (attention_tc128_kernel<true>): dropout on the probabilities, decisions hashed from (seed, row, key)
  uint64_t seed;
  uint32_t drop_threshold;
  float keep_scale;                  // 1 / (1 - p_drop)
  float* __restrict__ lse;           // (B*H*lq) log2-sum-exp of the scaled score rows (log2 units), or nullptr: what the
                                     // tensor-core backward recomputes P = 2^(s - lse) from
};

__device__ __forceinline__ uint32_t at_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void at_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void at_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void at_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void at_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > kAtSpinLimit) __trap();
  }
}
__device__ __forceinline__ void at_tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// K-major, 128-byte swizzle (rows of 64 bf16): 8-row groups 1024 B apart, descriptor version 1, layout type 2
__device__ __forceinline__ uint64_t at_desc_sw128(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// One lane of a converged warp: the TMA / MMA issuing roles run on all 32 lanes and predicate only the issuing instructions
// with this, so that their operands stay in uniform registers (see tc::elect_one in tc_common.cuh)
__device__ __forceinline__ bool at_elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\t@p mov.u32 %0, 1;\n\t}" : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void at_umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
// A operand in tensor memory (lane = row, each 32-bit column = two consecutive K elements), B from shared memory
__device__ __forceinline__ void at_umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void at_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void at_tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void at_tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// two values at once, packed (element 0 in the low half): hi = bf16x2(x), lo = bf16x2(x - hi); 6 instructions per pair
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi2, uint32_t& lo2) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(x1), "f"(x0));
  const float h0 = __uint_as_float(hi2 << 16), h1 = __uint_as_float(hi2 & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(x1 - h1), "f"(x0 - h0));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return static_cast<uint32_t>(__bfloat16_as_ushort(a)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16);
}

// ---------------------------------------------------------------------------------------------------
// split kernels
// ---------------------------------------------------------------------------------------------------
// rows of q or k: src (B*L, ld) fp32 with heads as 64-wide column slices -> dst_hi/lo ((b*H+h)*L + i, 64) bf16
__global__ void attn_split_rows_kernel(const float* __restrict__ src, int64_t ld, int64_t batch, int heads, int64_t len,
                                       float scale, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // one float4 each
  const int64_t per_row = heads * 16;
  if (i >= batch * len * per_row) return;
  const int64_t row = i / per_row;            // b*len + t
  const int c4 = static_cast<int>(i - row * per_row);
  const int h = c4 >> 4, d4 = (c4 & 15) * 4;
  const int64_t b = row / len, t = row - b * len;
  const float4 v = __ldg(reinterpret_cast<const float4*>(src + row * ld + h * 64 + d4));
  __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
  split_bf16(v.x * scale, h0, l0); split_bf16(v.y * scale, h1, l1);
  split_bf16(v.z * scale, h2, l2); split_bf16(v.w * scale, h3, l3);
  const int64_t o = (((b * heads + h) * len + t) * 64 + d4);
  *reinterpret_cast<uint2*>(hi + o) = make_uint2(pack2(h0, h1), pack2(h2, h3));
  *reinterpret_cast<uint2*>(lo + o) = make_uint2(pack2(l0, l1), pack2(l2, l3));
}

// v: src (B*L, ld) -> dst ((b*H+h)*64 + d, len_pad) bf16 (transposed per head); grid (ceil(len/64), H, B), 256 threads
__global__ void __launch_bounds__(256) attn_split_vt_kernel(const float* __restrict__ src, int64_t ld, int heads,
                                                            int64_t len, int64_t len_pad,
                                                            __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  __shared__ float tile[64][65];
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z, t0 = static_cast<int64_t>(blockIdx.x) * 64;
  const int tid = threadIdx.x;
  for (int e = tid; e < 64 * 16; e += 256) {
    const int r = e >> 4, c4 = (e & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t0 + r < len) v = __ldg(reinterpret_cast<const float4*>(src + (b * len + t0 + r) * ld + h * 64 + c4));
    tile[r][c4] = v.x; tile[r][c4 + 1] = v.y; tile[r][c4 + 2] = v.z; tile[r][c4 + 3] = v.w;
  }
  __syncthreads();
  // each thread writes 2 consecutive keys of one d row: 64 d x 32 key pairs = 2048 items
  for (int e = tid; e < 64 * 32; e += 256) {
    const int d = e >> 5, kp = (e & 31) * 2;
    if (t0 + kp >= len_pad) continue;
    __nv_bfloat16 h0, l0, h1, l1;
    split_bf16(tile[kp][d], h0, l0);
    split_bf16(tile[kp + 1][d], h1, l1);
    const int64_t o = ((b * heads + h) * 64 + d) * len_pad + t0 + kp;
    *reinterpret_cast<uint32_t*>(hi + o) = pack2(h0, h1);
    *reinterpret_cast<uint32_t*>(lo + o) = pack2(l0, l1);
  }
}

// ---------------------------------------------------------------------------------------------------
// attention kernel
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_qhi, const __grid_constant__ CUtensorMap map_qlo,
                    const __grid_constant__ CUtensorMap map_khi, const __grid_constant__ CUtensorMap map_klo,
                    const __grid_constant__ CUtensorMap map_vhi, const __grid_constant__ CUtensorMap map_vlo,
                    const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = at_smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  // smem: Q_g hi | lo (g = 0, 1), K/V ring, P_g hi | lo, barriers
  auto q_hi = [&](int g) { return base + static_cast<uint32_t>(g) * 2 * AT_Q_BYTES; };
  auto q_lo = [&](int g) { return base + static_cast<uint32_t>(g) * 2 * AT_Q_BYTES + AT_Q_BYTES; };
  const uint32_t stages = base + AT_GROUPS * 2 * AT_Q_BYTES;
  const uint32_t bars = stages + AT_STAGES * AT_STAGE_BYTES;
  // barriers: q_full | kv_full[4] | kv_empty[4] | s_full[g][2] | s_empty[g][2] | p_full[g] | p_empty[g]
  const uint32_t bar_q = bars;
  auto bar_kvf = [&](int s) { return bars + 8u * (1 + s); };
  auto bar_kve = [&](int s) { return bars + 8u * (1 + AT_STAGES + s); };
  auto bar_sf = [&](int g, int s) { return bars + 8u * (1 + 2 * AT_STAGES + 2 * g + s); };
  auto bar_se = [&](int g, int s) { return bars + 8u * (5 + 2 * AT_STAGES + 2 * g + s); };
  auto bar_pf = [&](int g) { return bars + 8u * (9 + 2 * AT_STAGES + g); };
  auto bar_pe = [&](int g) { return bars + 8u * (11 + 2 * AT_STAGES + g); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (bars - base) + 8 * (14 + 2 * AT_STAGES));

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ * AT_GROUPS;
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z;
  const int64_t bh = b * p.heads + h;
  const int kend = min(p.lk, p.kv_valid);
  const int T = (kend + AT_BK - 1) / AT_BK;

  if (threadIdx.x == 0) {
    at_mbar_init(bar_q, 1);
    for (int s = 0; s < AT_STAGES; ++s) { at_mbar_init(bar_kvf(s), 1); at_mbar_init(bar_kve(s), 1); }
    for (int g = 0; g < AT_GROUPS; ++g) {
      for (int s = 0; s < 2; ++s) { at_mbar_init(bar_sf(g, s), 1); at_mbar_init(bar_se(g, s), 4); }
      at_mbar_init(bar_pf(g), 4);
      at_mbar_init(bar_pe(g), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(at_smem_u32(tmem_slot)),
                 "r"(AT_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  auto tmem_s = [&](int g, int buf) { return tmem + static_cast<uint32_t>(g * 128 + buf * AT_BK); };
  auto tmem_o = [&](int g) { return tmem + static_cast<uint32_t>(256 + g * AT_D); };
  auto tmem_p = [&](int g) { return tmem + static_cast<uint32_t>(384 + g * 64); };   // hi pairs; lo pairs at + 32

  if (warp == 0) {
    {
      const bool leader = at_elect_one();        // all 32 lanes walk the loops; this one issues
      if (leader) at_mbar_expect_tx(bar_q, AT_GROUPS * 2 * AT_Q_BYTES);
      for (int g = 0; g < AT_GROUPS; ++g) {
        // (a tile that starts beyond this (sample, head)'s queries reads the neighbour's rows or zeros: its
        // results are never stored)
        const int qrow = static_cast<int>(bh * p.lq + q0 + g * AT_BQ);
        if (leader) at_tma_2d(q_hi(g), &map_qhi, bar_q, 0, qrow);
        if (leader) at_tma_2d(q_lo(g), &map_qlo, bar_q, 0, qrow);
      }
      for (int t = 0; t < T; ++t) {
        const int s = t % AT_STAGES;
        at_mbar_wait(bar_kve(s), ((t / AT_STAGES) & 1) ^ 1u);
        const uint32_t st = stages + s * AT_STAGE_BYTES;
        if (leader) at_mbar_expect_tx(bar_kvf(s), AT_STAGE_BYTES);
        const int krow = static_cast<int>(bh * p.lk + t * AT_BK);
        if (leader) at_tma_2d(st, &map_khi, bar_kvf(s), 0, krow);
        if (leader) at_tma_2d(st + AT_K_BYTES, &map_klo, bar_kvf(s), 0, krow);
        const int vrow = static_cast<int>(bh * AT_D);
        if (leader) at_tma_2d(st + 2 * AT_K_BYTES, &map_vhi, bar_kvf(s), t * AT_BK, vrow);
        if (leader) at_tma_2d(st + 3 * AT_K_BYTES, &map_vlo, bar_kvf(s), t * AT_BK, vrow);
      }
    }
  } else if (warp == 1) {
    {
      const bool leader = at_elect_one();        // all 32 lanes walk the loops; this one issues
      // kind::f16: D = F32 (1<<4), A = B = BF16 (1<<7, 1<<10), K-major, N = 64 (8<<17), M = 128 (8<<24)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(AT_BK >> 3) << 17) |
                             (static_cast<uint32_t>(AT_BQ >> 4) << 24);
      auto issue_qk = [&](int g, int t) {
        const int s = t % AT_STAGES;
        at_mbar_wait(bar_kvf(s), (t / AT_STAGES) & 1);
        at_mbar_wait(bar_se(g, t & 1), ((t >> 1) & 1) ^ 1u);   // score buffer drained by the softmax warps
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = stages + s * AT_STAGE_BYTES;
        const uint64_t d_qhi = at_desc_sw128(q_hi(g)), d_qlo = at_desc_sw128(q_lo(g));
        const uint64_t d_khi = at_desc_sw128(st), d_klo = at_desc_sw128(st + AT_K_BYTES);
        const uint32_t d_s = tmem_s(g, t & 1);
#pragma unroll
        for (int kk = 0; kk < AT_D / 16; ++kk) {
          const uint64_t adv = static_cast<uint64_t>(kk * 2);   // 16 bf16 = 32 bytes
          if (leader) at_umma_bf16(d_s, d_qlo + adv, d_khi + adv, idesc, kk != 0 ? 1u : 0u);
          if (leader) at_umma_bf16(d_s, d_qhi + adv, d_klo + adv, idesc, 1u);
          if (leader) at_umma_bf16(d_s, d_qhi + adv, d_khi + adv, idesc, 1u);
        }
        if (leader) at_commit(bar_sf(g, t & 1));
      };
      at_mbar_wait(bar_q, 0);
      for (int g = 0; g < AT_GROUPS; ++g) issue_qk(g, 0);
      for (int t = 0; t < T; ++t) {
        if (t + 1 < T)
          for (int g = 0; g < AT_GROUPS; ++g) issue_qk(g, t + 1);
        const int s = t % AT_STAGES;
        const uint32_t st = stages + s * AT_STAGE_BYTES;
        const uint64_t d_vhi = at_desc_sw128(st + 2 * AT_K_BYTES), d_vlo = at_desc_sw128(st + 3 * AT_K_BYTES);
        for (int g = 0; g < AT_GROUPS; ++g) {
          at_mbar_wait(bar_pf(g), t & 1);                      // P_t of this group written, its O rescaled
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = tmem_p(g), a_lo = tmem_p(g) + 32;
          const uint32_t d_o = tmem_o(g);
#pragma unroll
          for (int kk = 0; kk < AT_BK / 16; ++kk) {
            const uint64_t adv = static_cast<uint64_t>(kk * 2);      // V^T: 16 keys = 32 bytes
            const uint32_t ka = static_cast<uint32_t>(kk * 8);        // P in TMEM: 16 bf16 = 8 columns
            if (leader) at_umma_bf16_ts(d_o, a_lo + ka, d_vhi + adv, idesc, (t | kk) != 0 ? 1u : 0u);
            if (leader) at_umma_bf16_ts(d_o, a_hi + ka, d_vlo + adv, idesc, 1u);
            if (leader) at_umma_bf16_ts(d_o, a_hi + ka, d_vhi + adv, idesc, 1u);
          }
          if (leader) at_commit(bar_pe(g));       // this group's P buffer free, its O updated
        }
        if (leader) at_commit(bar_kve(s));        // K/V stage free (both groups' P.V have read it)
      }
    }
  } else {
    // ------------------------------------------------ softmax warps: thread <-> query row
    const int g = (warp - 2) >> 2;                     // softmax group = query tile of this CTA
    const int q = warp & 3;
    const int r = q * 32 + lane;                       // row in the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    for (int t = 0; t < T; ++t) {
      at_mbar_wait(bar_sf(g, t & 1), (t >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t sr[64];
      at_tmem_ld32(tmem_s(g, t & 1) + lane_addr, sr);
      at_tmem_ld32(tmem_s(g, t & 1) + lane_addr + 32, sr + 32);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) at_mbar_arrive(bar_se(g, t & 1));
      const int valid = kend - t * AT_BK;              // keys of this tile that exist
      if (valid < AT_BK) {                             // ragged last tile only (warp-uniform)
#pragma unroll
        for (int j = 0; j < 64; ++j)
          if (j >= valid) sr[j] = __float_as_uint(-INFINITY);
      }
      // scores are in log2 units (q was pre-scaled by log2(e) / 8): p = 2^(s - m).  Four independent max / sum
      // chains: one warp per scheduler cannot hide a 64-long dependent chain.
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < 64; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(sr[j]));
      // Lazy rescaling: the running reference m_run only moves when the tile maximum exceeds it by more than 2^8
      // (scores are in log2 units), so p = 2^(s - m_run) stays below 256 -- harmless in the fp32 sums and in the
      // bf16 hi/lo split (relative precision) -- and O / l need rescaling only in the first few K/V tiles instead of
      // in most of them.  The O rescale is a 32 KB TMEM read + write per tile (TMEM reads run at 64 B/clk: 512 cycles
      // against 768 cycles of MMAs) and sits on the P.V critical path.  The final O / l is unchanged mathematically.
      const float m_tile = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_new = (m_tile - m_run > AT_RESCALE_LOG2) ? m_tile : m_run;      // first tile: m_run = -inf -> taken
      const float alpha = (m_new == m_run) ? 1.f : ex2_approx(m_run - m_new);
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const float e = ex2_approx(__uint_as_float(sr[j]) - m_new);
        sr[j] = __float_as_uint(e);
        rs4[j & 3] += e;
      }
      l_run = l_run * alpha + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
      m_run = m_new;
      // previous P.V must have completed before P is overwritten / O rescaled
      at_mbar_wait(bar_pe(g), (t & 1) ^ 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (t > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        uint32_t o[32];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          at_tmem_ld32(tmem_o(g) + lane_addr + half * 32, o);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
          at_tmem_st32(tmem_o(g) + lane_addr + half * 32, o);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      // P row -> packed bf16 pairs (element 2e in the low half), hi and lo, written to this lane's TMEM row: the
      // A operand of the P.V MMA
      {
        uint32_t ph[32], pl[32];
#pragma unroll
        for (int e = 0; e < 32; ++e)
          split_bf16x2(__uint_as_float(sr[2 * e]), __uint_as_float(sr[2 * e + 1]), ph[e], pl[e]);
        at_tmem_st32(tmem_p(g) + lane_addr, ph);
        at_tmem_st32(tmem_p(g) + lane_addr + 32, pl);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) at_mbar_arrive(bar_pf(g));
    }
    // epilogue: wait for the last P.V, then O / l
    at_mbar_wait(bar_pe(g), (T & 1) ^ 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = q0 + g * AT_BQ + r;
    const float inv = __fdiv_rn(1.f, l_run);
    const int64_t oo = (b * p.lq + row) * p.ldo + h * AT_D;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t o[32];
      at_tmem_ld32(tmem_o(g) + lane_addr + half * 32, o);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < p.lq) {
        if (p.out != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            *reinterpret_cast<float4*>(p.out + oo + half * 32 + j) =
                make_float4(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv,
                            __uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float x0 = __uint_as_float(o[j + 2 * e]) * inv, x1 = __uint_as_float(o[j + 2 * e + 1]) * inv;
              asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hw[e]) : "f"(x1), "f"(x0));
              const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
              asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lw[e]) : "f"((x1 - hf.y) * 2048.f), "f"((x0 - hf.x) * 2048.f));
            }
            *reinterpret_cast<uint4*>(p.out_hi + oo + half * 32 + j) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(p.out_lo + oo + half * 32 + j) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(AT_TMEM_COLS) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------
// attention kernel, 128-key tiles (the default for lk >= 128)
// ---------------------------------------------------------------------------------------------------
// Measured on B200 (scripts/umma_rate.cu, profiles/r01_umma_rate.txt): a tcgen05.mma M128 x N x K16 costs N/2 cycles only
// down to N = 128; below that there is a ~46-cycle floor (N = 64: 48 cycles), so the 64-key kernel above runs the
// tensor core at 2/3 rate at best and pays one softmax <-> tensor hand-over per 64 keys.  Here a K/V tile has 128 keys:
//   * S_t = Q.K_t^T is M128 x N128 (full rate); P_t.V_t stays N = 64 (the head dimension) with K = 128;
//   * TMEM per query tile g: one 128-column region that holds S_t (fp32) and is then OVERWRITTEN in place by P_t
//     (packed bf16 pairs: hi in columns [0, 64), lo in [64, 128)) + the 64-column O accumulator = 192 columns, 384 for
//     the two query tiles of a CTA.  tcgen05.mma instructions of one thread execute in order, so Q.K_{t+1}^T, issued
//     right after P_t.V_t, cannot overwrite P_t early, and "S_{t+1} complete" implies "P_t.V_t complete": the softmax
//     warps need no separate barrier before rescaling O or writing P_{t+1};
//   * K and V^T tiles travel through separate 2-deep rings (K_{t+2} is fetched as soon as both Q.K_t^T have
//     been issued and completed, V_{t+2} after both P_t.V_t), 32 KB per tile each.
constexpr int A2_BK = 128;
constexpr int A2_STAGES = 2;
constexpr int A2_K_BYTES = A2_BK * AT_D * 2;       // 16 KB per bf16 term
constexpr int A2_V_HALF = AT_D * 64 * 2;           // one 64-key box of V^T: 8 KB
constexpr int A2_KSTAGE = 2 * A2_K_BYTES;          // K_hi | K_lo
constexpr int A2_VSTAGE = 4 * A2_V_HALF;           // Vt_hi (2 boxes) | Vt_lo (2 boxes)
constexpr int A2_SMEM_BYTES = AT_GROUPS * 2 * AT_Q_BYTES + A2_STAGES * (A2_KSTAGE + A2_VSTAGE) + 1024 + 256;
constexpr uint32_t A2_GROUP_COLS = 192;            // S/P 128 + O 64

__device__ __forceinline__ void at_tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr)
      : "memory");
}

// DROP: nn.MultiheadAttention's dropout on the probabilities (training, upstream cfg.dropout = 0.1): the row sum l keeps
// every p, the P operand of P.V carries keep ? p / (1 - q) : 0 -- O / l is then dropout(softmax(S)) . V.  keep() is
// common.cuh's counter hash with row = (sample * heads + head) * lq + query and col = key, the indexing of
// hoisdf_softmax_dropout_rows_fwd / _bwd on the (B, H, Lq, Lk) probabilities, so the backward regenerates the decisions.
template <bool DROP>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc128_kernel(const __grid_constant__ CUtensorMap map_qhi, const __grid_constant__ CUtensorMap map_qlo,
                       const __grid_constant__ CUtensorMap map_khi, const __grid_constant__ CUtensorMap map_klo,
                       const __grid_constant__ CUtensorMap map_vhi, const __grid_constant__ CUtensorMap map_vlo,
                       const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = at_smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  auto q_hi = [&](int g) { return base + static_cast<uint32_t>(g) * 2 * AT_Q_BYTES; };
  auto q_lo = [&](int g) { return base + static_cast<uint32_t>(g) * 2 * AT_Q_BYTES + AT_Q_BYTES; };
  const uint32_t kring = base + AT_GROUPS * 2 * AT_Q_BYTES;
  const uint32_t vring = kring + A2_STAGES * A2_KSTAGE;
  const uint32_t bars = vring + A2_STAGES * A2_VSTAGE;
  // barriers: q | k_full[2] k_empty[2] v_full[2] v_empty[2] | s_full[g] p_full[g] o_full[g]
  const uint32_t bar_q = bars;
  auto bar_kf = [&](int s) { return bars + 8u * (1 + s); };
  auto bar_ke = [&](int s) { return bars + 8u * (3 + s); };
  auto bar_vf = [&](int s) { return bars + 8u * (5 + s); };
  auto bar_ve = [&](int s) { return bars + 8u * (7 + s); };
  auto bar_sf = [&](int g) { return bars + 8u * (9 + g); };
  auto bar_pf = [&](int g) { return bars + 8u * (11 + g); };
  auto bar_of = [&](int g) { return bars + 8u * (13 + g); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (bars - base) + 8 * 16);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ * AT_GROUPS;
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z;
  const int64_t bh = b * p.heads + h;
  const int kend = min(p.lk, p.kv_valid);
  const int T = (kend + A2_BK - 1) / A2_BK;

  if (threadIdx.x == 0) {
    at_mbar_init(bar_q, 1);
    for (int s = 0; s < A2_STAGES; ++s) {
      at_mbar_init(bar_kf(s), 1); at_mbar_init(bar_ke(s), 1);
      at_mbar_init(bar_vf(s), 1); at_mbar_init(bar_ve(s), 1);
    }
    for (int g = 0; g < AT_GROUPS; ++g) {
      at_mbar_init(bar_sf(g), 1);
      at_mbar_init(bar_pf(g), 4);
      at_mbar_init(bar_of(g), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(at_smem_u32(tmem_slot)),
                 "r"(AT_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  auto tmem_sp = [&](int g) { return tmem + static_cast<uint32_t>(g) * A2_GROUP_COLS; };          // S, then P in place
  auto tmem_o = [&](int g) { return tmem + static_cast<uint32_t>(g) * A2_GROUP_COLS + 128u; };

  if (warp == 0) {
    {
      const bool leader = at_elect_one();        // all 32 lanes walk the loops; this one issues
      if (leader) at_mbar_expect_tx(bar_q, AT_GROUPS * 2 * AT_Q_BYTES);
      for (int g = 0; g < AT_GROUPS; ++g) {
        const int qrow = static_cast<int>(bh * p.lq + q0 + g * AT_BQ);
        if (leader) at_tma_2d(q_hi(g), &map_qhi, bar_q, 0, qrow);
        if (leader) at_tma_2d(q_lo(g), &map_qlo, bar_q, 0, qrow);
      }
      const int vrow = static_cast<int>(bh * AT_D);
      for (int t = 0; t < T; ++t) {
        const int s = t % A2_STAGES;
        const uint32_t ph = ((t / A2_STAGES) & 1) ^ 1u;
        at_mbar_wait(bar_ke(s), ph);
        const uint32_t ks = kring + s * A2_KSTAGE;
        if (leader) at_mbar_expect_tx(bar_kf(s), A2_KSTAGE);
        const int krow = static_cast<int>(bh * p.lk + t * A2_BK);
        if (leader) at_tma_2d(ks, &map_khi, bar_kf(s), 0, krow);
        if (leader) at_tma_2d(ks + A2_K_BYTES, &map_klo, bar_kf(s), 0, krow);
        at_mbar_wait(bar_ve(s), ph);
        const uint32_t vs = vring + s * A2_VSTAGE;
        if (leader) at_mbar_expect_tx(bar_vf(s), A2_VSTAGE);
        if (leader) at_tma_2d(vs, &map_vhi, bar_vf(s), t * A2_BK, vrow);
        if (leader) at_tma_2d(vs + A2_V_HALF, &map_vhi, bar_vf(s), t * A2_BK + 64, vrow);
        if (leader) at_tma_2d(vs + 2 * A2_V_HALF, &map_vlo, bar_vf(s), t * A2_BK, vrow);
        if (leader) at_tma_2d(vs + 3 * A2_V_HALF, &map_vlo, bar_vf(s), t * A2_BK + 64, vrow);
      }
    }
  } else if (warp == 1) {
    {
      const bool leader = at_elect_one();        // all 32 lanes walk the loops; this one issues
      // kind::f16: D = F32 (1<<4), A = B = BF16 (1<<7, 1<<10), K-major, N >> 3 at bit 17, M = 128 (8<<24)
      const uint32_t idesc_common = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(AT_BQ >> 4) << 24);
      const uint32_t idesc_qk = idesc_common | (static_cast<uint32_t>(A2_BK >> 3) << 17);
      const uint32_t idesc_pv = idesc_common | (static_cast<uint32_t>(AT_D >> 3) << 17);
      auto issue_qk = [&](int g, int t) {
        const int s = t % A2_STAGES;
        const uint32_t ks = kring + s * A2_KSTAGE;
        const uint64_t d_qhi = at_desc_sw128(q_hi(g)), d_qlo = at_desc_sw128(q_lo(g));
        const uint64_t d_khi = at_desc_sw128(ks), d_klo = at_desc_sw128(ks + A2_K_BYTES);
        const uint32_t d_s = tmem_sp(g);
#pragma unroll
        for (int kk = 0; kk < AT_D / 16; ++kk) {
          const uint64_t adv = static_cast<uint64_t>(kk * 2);   // 16 bf16 = 32 bytes
          if (leader) at_umma_bf16(d_s, d_qlo + adv, d_khi + adv, idesc_qk, kk != 0 ? 1u : 0u);
          if (leader) at_umma_bf16(d_s, d_qhi + adv, d_klo + adv, idesc_qk, 1u);
          if (leader) at_umma_bf16(d_s, d_qhi + adv, d_khi + adv, idesc_qk, 1u);
        }
        if (leader) at_commit(bar_sf(g));
      };
      at_mbar_wait(bar_q, 0);
      at_mbar_wait(bar_kf(0), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int g = 0; g < AT_GROUPS; ++g) issue_qk(g, 0);
      if (leader) at_commit(bar_ke(0));
      for (int t = 0; t < T; ++t) {
        const int s = t % A2_STAGES;
        const uint32_t vs = vring + s * A2_VSTAGE;
        at_mbar_wait(bar_vf(s), (t / A2_STAGES) & 1);
        for (int g = 0; g < AT_GROUPS; ++g) {
          at_mbar_wait(bar_pf(g), t & 1);                        // P_t of this group written, its O rescaled
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = tmem_sp(g), a_lo = tmem_sp(g) + 64;
          const uint32_t d_o = tmem_o(g);
#pragma unroll
          for (int kk = 0; kk < A2_BK / 16; ++kk) {
            const uint32_t box = static_cast<uint32_t>(kk >> 2) * A2_V_HALF;          // which 64-key box of V^T
            const uint64_t adv = static_cast<uint64_t>((kk & 3) * 2);                 // 16 keys = 32 bytes inside the box
            const uint64_t d_vhi = at_desc_sw128(vs + box) + adv, d_vlo = at_desc_sw128(vs + 2 * A2_V_HALF + box) + adv;
            const uint32_t ka = static_cast<uint32_t>(kk * 8);                         // P in TMEM: 16 bf16 = 8 columns
            if (leader) at_umma_bf16_ts(d_o, a_lo + ka, d_vhi, idesc_pv, (t | kk) != 0 ? 1u : 0u);
            if (leader) at_umma_bf16_ts(d_o, a_hi + ka, d_vlo, idesc_pv, 1u);
            if (leader) at_umma_bf16_ts(d_o, a_hi + ka, d_vhi, idesc_pv, 1u);
          }
          if (t + 1 < T) {
            if (g == 0) {
              at_mbar_wait(bar_kf((t + 1) % A2_STAGES), ((t + 1) / A2_STAGES) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            issue_qk(g, t + 1);                                   // executes after P_t.V_t: S_{t+1} overwrites P_t
          } else {
            if (leader) at_commit(bar_of(g));                                 // last P.V of this group done: O final
          }
        }
        if (leader) at_commit(bar_ve(s));                                     // V_t consumed by both groups
        if (t + 1 < T) at_commit(bar_ke((t + 1) % A2_STAGES));    // K_{t+1} consumed by both groups
      }
    }
  } else {
    // ------------------------------------------------ softmax warps: thread <-> query row
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    uint32_t drop_key = 0;
    if (DROP) drop_key = dropout_row_key(p.seed, static_cast<uint64_t>(bh) * p.lq + static_cast<uint64_t>(q0 + g * AT_BQ + r));
    for (int t = 0; t < T; ++t) {
      at_mbar_wait(bar_sf(g), t & 1);          // S_t complete (and with it P_{t-1}.V_{t-1}: same issuing thread, in order)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t sr[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) at_tmem_ld32(tmem_sp(g) + lane_addr + c * 32, sr + c * 32);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int valid = kend - t * A2_BK;
      if (valid < A2_BK) {                               // ragged last tile only (warp-uniform)
#pragma unroll
        for (int j = 0; j < 128; ++j)
          if (j >= valid) sr[j] = __float_as_uint(-INFINITY);
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < 128; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(sr[j]));
      const float m_tile = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_new = (m_tile - m_run > AT_RESCALE_LOG2) ? m_tile : m_run;      // lazy rescaling, see above
      const float alpha = (m_new == m_run) ? 1.f : ex2_approx(m_run - m_new);
      if (t > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        uint32_t o[32];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          at_tmem_ld32(tmem_o(g) + lane_addr + half * 32, o);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
          at_tmem_st32(tmem_o(g) + lane_addr + half * 32, o);
        }
      }
      // p = 2^(s - m), row sum, and P as packed bf16 pairs (hi, lo) written over S, 32 keys at a time
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t ph[16], pl[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float e0 = ex2_approx(__uint_as_float(sr[c * 32 + 2 * e]) - m_new);
          float e1 = ex2_approx(__uint_as_float(sr[c * 32 + 2 * e + 1]) - m_new);
          rs4[e & 3] += e0 + e1;
          if (DROP) {
            const uint32_t key0 = static_cast<uint32_t>(t * A2_BK + c * 32 + 2 * e);
            e0 = dropout_keep(drop_key, key0, p.drop_threshold) ? e0 * p.keep_scale : 0.f;
            e1 = dropout_keep(drop_key, key0 + 1, p.drop_threshold) ? e1 * p.keep_scale : 0.f;
          }
          split_bf16x2(e0, e1, ph[e], pl[e]);
        }
        at_tmem_st16(tmem_sp(g) + lane_addr + c * 16, ph);
        at_tmem_st16(tmem_sp(g) + lane_addr + 64 + c * 16, pl);
      }
      l_run = l_run * alpha + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
      m_run = m_new;
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) at_mbar_arrive(bar_pf(g));
    }
    at_mbar_wait(bar_of(g), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = q0 + g * AT_BQ + r;
    const float inv = __fdiv_rn(1.f, l_run);
    if (p.lse != nullptr && row < p.lq) p.lse[bh * p.lq + row] = m_run + log2f(l_run);
    const int64_t oo = (b * p.lq + row) * p.ldo + h * AT_D;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t o[32];
      at_tmem_ld32(tmem_o(g) + lane_addr + half * 32, o);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < p.lq) {
        if (p.out != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            *reinterpret_cast<float4*>(p.out + oo + half * 32 + j) =
                make_float4(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv,
                            __uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float x0 = __uint_as_float(o[j + 2 * e]) * inv, x1 = __uint_as_float(o[j + 2 * e + 1]) * inv;
              asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hw[e]) : "f"(x1), "f"(x0));
              const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
              asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lw[e]) : "f"((x1 - hf.y) * 2048.f), "f"((x0 - hf.x) * 2048.f));
            }
            *reinterpret_cast<uint4*>(p.out_hi + oo + half * 32 + j) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(p.out_lo + oo + half * 32 + j) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(AT_TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// backward of the encoder self-attention on the tensor cores (training step)
// ---------------------------------------------------------------------------------------------------
// With the forward's log-sum-exp rows (lse, log2 units) and delta_i = sum_d dO_id O_id the probabilities are recomputed tile
// by tile, P_ij = 2^(s_ij - lse_i), and never stored:
//      g_ij  = keep_ij ? (dO V^T)_ij / (1 - q) : 0            (keep == 1 without dropout)
//      dS_ij = P_ij (g_ij - delta_i)
//      dQ = dS K / 8         dK = dS^T Q / 8         dV = Pd^T dO,   Pd_ij = keep_ij ? P_ij / (1 - q) : 0
// Two kernels, both built from the forward's pieces (BF16x3 products, S in TMEM, the softmax warps rewrite it in place as
// the bf16 hi / lo A operand of the next MMA, every shared-memory operand K-major through TMA with the 128-byte swizzle):
//   attention_bwd_dq_kernel   CTA = 128 queries of one (sample, head); TMEM lanes = queries.  Per 128-key tile:
//                             S = Q K_t^T and dP = dO V_t^T (M128 N128 K64), dS in place of S, dQ += dS K_t (B = K^T copy)
//   attention_bwd_dkv_kernel  CTA = 128 keys; TMEM lanes = keys.  Per 128-query tile: S^T = K Q_t^T, dP^T = V dO_t^T,
//                             Pd^T in place of S^T, dS^T in place of dP^T, dV += Pd^T dO_t (B = dO^T copy),
//                             dK += dS^T Q_t (B = Q^T copy)
// so every product has its contraction axis contiguous in both operands and no accumulator is shared between CTAs (no
// atomics; the price is that S and dP are formed twice, 2 x 24 of the 120 MMAs per tile pair).
// 320 threads: warp 0 TMA, warp 1 MMA issue, warps 2..9 the element-wise stage -- TMEM lane quarter = warp % 4, the two
// warps of a quarter split the 128 columns.  In-place operand layout, per chunk c of 32 columns: fp32 columns [32c, 32c+32)
// become bf16 pairs hi [32c, 32c+16) | lo [32c+16, 32c+32), so a warp only overwrites columns it has read itself.
constexpr int AB_THREADS = 64 + 256;
constexpr int AB_TILE = 128 * AT_D * 2;            // one bf16 term of a 128 x 64 row tile: 16 KB
constexpr int AB_TBOX = AT_D * 64 * 2;             // one 64-column box of a transposed (64 x L) operand: 8 KB
constexpr int AB_DQ_SMEM = 8 * AB_TILE + 4 * AB_TBOX + 1024 + 256;          // Q dO | K V | K^T
constexpr int AB_DKV_SMEM = 8 * AB_TILE + 8 * AB_TBOX + 1024 + 256;         // K V | Q dO | Q^T dO^T

struct AttnBwdParams {
  const float* __restrict__ lse;      // (B*H*lq), log2 units (attention_tc128_kernel)
  const float* __restrict__ delta;    // (B*H*lq)
  float* __restrict__ dq;             // (B*lq, ldg)
  float* __restrict__ dk;             // (B*lk, ldg)
  float* __restrict__ dv;
  int64_t ldg;
  int lq, lk, kv_valid, heads;
  uint64_t seed;
  uint32_t drop_threshold;
  float keep_scale;
};

// delta[(b*H + h)*lq + i] = sum_d dO[b*lq + i][64h + d] * O[...]: one warp per (row, head)
__global__ void attn_bwd_delta_kernel(const float* __restrict__ dout, const float* __restrict__ out, int64_t ldo, int64_t batch,
                                      int heads, int64_t lq, float* __restrict__ delta) {
  const int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= batch * lq * heads) return;
  const int64_t row = w / heads;
  const int h = static_cast<int>(w - row * heads);
  const int64_t b = row / lq, i = row - b * lq;
  const float2 a = __ldg(reinterpret_cast<const float2*>(dout + row * ldo + h * 64 + lane * 2));
  const float2 o = __ldg(reinterpret_cast<const float2*>(out + row * ldo + h * 64 + lane * 2));
  const float s = warp_sum(fmaf(a.x, o.x, a.y * o.y));
  if (lane == 0) delta[(b * heads + h) * lq + i] = s;
}

__device__ __forceinline__ void at_named_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <bool DROP>
__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_dq_kernel(const __grid_constant__ CUtensorMap map_qhi, const __grid_constant__ CUtensorMap map_qlo,
                        const __grid_constant__ CUtensorMap map_dohi, const __grid_constant__ CUtensorMap map_dolo,
                        const __grid_constant__ CUtensorMap map_khi, const __grid_constant__ CUtensorMap map_klo,
                        const __grid_constant__ CUtensorMap map_vhi, const __grid_constant__ CUtensorMap map_vlo,
                        const __grid_constant__ CUtensorMap map_kthi, const __grid_constant__ CUtensorMap map_ktlo,
                        const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = at_smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t s_q = base;                        // Q hi | Q lo | dO hi | dO lo   (resident)
  const uint32_t s_kv = base + 4 * AB_TILE;         // K hi | K lo | V hi | V lo     (key tile t)
  const uint32_t s_kt = base + 8 * AB_TILE;         // K^T hi (2 boxes) | K^T lo (2 boxes)
  const uint32_t bars = s_kt + 4 * AB_TBOX;
  const uint32_t bar_q = bars, bar_kvf = bars + 8, bar_kve = bars + 16, bar_ktf = bars + 24, bar_kte = bars + 32,
                 bar_sf = bars + 40, bar_pf = bars + 48, bar_of = bars + 56;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (bars - base) + 64);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z;
  const int64_t bh = b * p.heads + h;
  const int kend = min(p.lk, p.kv_valid);
  const int T = (kend + 127) / 128;

  if (threadIdx.x == 0) {
    at_mbar_init(bar_q, 1);
    at_mbar_init(bar_kvf, 1); at_mbar_init(bar_kve, 1);
    at_mbar_init(bar_ktf, 1); at_mbar_init(bar_kte, 1);
    at_mbar_init(bar_sf, 1); at_mbar_init(bar_pf, 8); at_mbar_init(bar_of, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(at_smem_u32(tmem_slot)),
                 "r"(AT_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t tmem_s = tmem, tmem_dp = tmem + 128u, tmem_acc = tmem + 256u;

  if (warp == 0) {
    const bool leader = at_elect_one();
    const int qrow = static_cast<int>(bh * p.lq + q0);
    if (leader) at_mbar_expect_tx(bar_q, 4 * AB_TILE);
    if (leader) at_tma_2d(s_q, &map_qhi, bar_q, 0, qrow);
    if (leader) at_tma_2d(s_q + AB_TILE, &map_qlo, bar_q, 0, qrow);
    if (leader) at_tma_2d(s_q + 2 * AB_TILE, &map_dohi, bar_q, 0, qrow);
    if (leader) at_tma_2d(s_q + 3 * AB_TILE, &map_dolo, bar_q, 0, qrow);
    const int trow = static_cast<int>(bh * AT_D);
    for (int t = 0; t < T; ++t) {
      const uint32_t ph = (t & 1) ^ 1u;
      const int krow = static_cast<int>(bh * p.lk + t * 128);
      at_mbar_wait(bar_kve, ph);
      if (leader) at_mbar_expect_tx(bar_kvf, 4 * AB_TILE);
      if (leader) at_tma_2d(s_kv, &map_khi, bar_kvf, 0, krow);
      if (leader) at_tma_2d(s_kv + AB_TILE, &map_klo, bar_kvf, 0, krow);
      if (leader) at_tma_2d(s_kv + 2 * AB_TILE, &map_vhi, bar_kvf, 0, krow);
      if (leader) at_tma_2d(s_kv + 3 * AB_TILE, &map_vlo, bar_kvf, 0, krow);
      at_mbar_wait(bar_kte, ph);
      if (leader) at_mbar_expect_tx(bar_ktf, 4 * AB_TBOX);
      if (leader) at_tma_2d(s_kt, &map_kthi, bar_ktf, t * 128, trow);
      if (leader) at_tma_2d(s_kt + AB_TBOX, &map_kthi, bar_ktf, t * 128 + 64, trow);
      if (leader) at_tma_2d(s_kt + 2 * AB_TBOX, &map_ktlo, bar_ktf, t * 128, trow);
      if (leader) at_tma_2d(s_kt + 3 * AB_TBOX, &map_ktlo, bar_ktf, t * 128 + 64, trow);
    }
  } else if (warp == 1) {
    const bool leader = at_elect_one();
    const uint32_t idesc_common = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(128 >> 4) << 24);
    const uint32_t idesc_n128 = idesc_common | (static_cast<uint32_t>(128 >> 3) << 17);
    const uint32_t idesc_n64 = idesc_common | (static_cast<uint32_t>(64 >> 3) << 17);
    const uint64_t d_qhi = at_desc_sw128(s_q), d_qlo = at_desc_sw128(s_q + AB_TILE);
    const uint64_t d_dohi = at_desc_sw128(s_q + 2 * AB_TILE), d_dolo = at_desc_sw128(s_q + 3 * AB_TILE);
    const uint64_t d_khi = at_desc_sw128(s_kv), d_klo = at_desc_sw128(s_kv + AB_TILE);
    const uint64_t d_vhi = at_desc_sw128(s_kv + 2 * AB_TILE), d_vlo = at_desc_sw128(s_kv + 3 * AB_TILE);
    at_mbar_wait(bar_q, 0);
    for (int t = 0; t < T; ++t) {
      at_mbar_wait(bar_kvf, t & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int kk = 0; kk < AT_D / 16; ++kk) {                   // S = Q K_t^T
        const uint64_t adv = static_cast<uint64_t>(kk * 2);
        if (leader) at_umma_bf16(tmem_s, d_qlo + adv, d_khi + adv, idesc_n128, kk != 0 ? 1u : 0u);
        if (leader) at_umma_bf16(tmem_s, d_qhi + adv, d_klo + adv, idesc_n128, 1u);
        if (leader) at_umma_bf16(tmem_s, d_qhi + adv, d_khi + adv, idesc_n128, 1u);
      }
#pragma unroll
      for (int kk = 0; kk < AT_D / 16; ++kk) {                   // dP = dO V_t^T
        const uint64_t adv = static_cast<uint64_t>(kk * 2);
        if (leader) at_umma_bf16(tmem_dp, d_dolo + adv, d_vhi + adv, idesc_n128, kk != 0 ? 1u : 0u);
        if (leader) at_umma_bf16(tmem_dp, d_dohi + adv, d_vlo + adv, idesc_n128, 1u);
        if (leader) at_umma_bf16(tmem_dp, d_dohi + adv, d_vhi + adv, idesc_n128, 1u);
      }
      if (leader) at_commit(bar_sf);
      if (leader) at_commit(bar_kve);                            // K / V row tiles consumed
      at_mbar_wait(bar_pf, t & 1);                               // dS_t written over S_t
      at_mbar_wait(bar_ktf, t & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {                           // dQ += dS K_t: 16 keys per step
        const uint32_t a_hi = tmem_s + static_cast<uint32_t>((kk >> 1) * 32 + (kk & 1) * 8), a_lo = a_hi + 16;
        const uint32_t box = static_cast<uint32_t>(kk >> 2) * AB_TBOX;
        const uint64_t adv = static_cast<uint64_t>((kk & 3) * 2);
        const uint64_t d_hi = at_desc_sw128(s_kt + box) + adv, d_lo = at_desc_sw128(s_kt + 2 * AB_TBOX + box) + adv;
        if (leader) at_umma_bf16_ts(tmem_acc, a_lo, d_hi, idesc_n64, (t | kk) != 0 ? 1u : 0u);
        if (leader) at_umma_bf16_ts(tmem_acc, a_hi, d_lo, idesc_n64, 1u);
        if (leader) at_umma_bf16_ts(tmem_acc, a_hi, d_hi, idesc_n64, 1u);
      }
      if (leader) at_commit(bar_kte);
      if (t + 1 == T && leader) at_commit(bar_of);
    }
  } else {
    const int qd = warp & 3;                           // TMEM lane quarter of this warp
    const int half = (warp - 2) >> 2;                  // which 64 of the 128 columns
    const int r = qd * 32 + lane;                      // query row of the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const int qi = q0 + r;
    const bool row_ok = qi < p.lq;
    // an absent row: lse = +inf makes every p exactly 0
    const float lse = row_ok ? __ldg(p.lse + bh * p.lq + qi) : INFINITY;
    const float delta = row_ok ? __ldg(p.delta + bh * p.lq + qi) : 0.f;
    uint32_t drop_key = 0;
    if (DROP) drop_key = dropout_row_key(p.seed, static_cast<uint64_t>(bh) * p.lq + static_cast<uint64_t>(qi));
    for (int t = 0; t < T; ++t) {
      at_mbar_wait(bar_sf, t & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int valid = kend - t * 128;                // keys of this tile that exist
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;
        uint32_t sv[32], dp[32], hi[16], lo[16];
        at_tmem_ld32(tmem_s + lane_addr + c * 32, sv);
        at_tmem_ld32(tmem_dp + lane_addr + c * 32, dp);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float ds[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int col = c * 32 + 2 * e + u;
            const float pv = ex2_approx(__uint_as_float(sv[2 * e + u]) - lse);
            float g = __uint_as_float(dp[2 * e + u]);
            if (DROP) g = dropout_keep(drop_key, static_cast<uint32_t>(t * 128 + col), p.drop_threshold) ? g * p.keep_scale : 0.f;
            ds[u] = col < valid ? pv * (g - delta) : 0.f;
          }
          split_bf16x2(ds[0], ds[1], hi[e], lo[e]);
        }
        at_tmem_st16(tmem_s + lane_addr + c * 32, hi);
        at_tmem_st16(tmem_s + lane_addr + c * 32 + 16, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) at_mbar_arrive(bar_pf);
    }
    at_mbar_wait(bar_of, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t o[32];
    at_tmem_ld32(tmem_acc + lane_addr + half * 32, o);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (row_ok) {
      float* dst = p.dq + (b * p.lq + qi) * p.ldg + h * AT_D + half * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(dst + j) =
            make_float4(__uint_as_float(o[j]) * 0.125f, __uint_as_float(o[j + 1]) * 0.125f,
                        __uint_as_float(o[j + 2]) * 0.125f, __uint_as_float(o[j + 3]) * 0.125f);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(AT_TMEM_COLS) : "memory");
  }
}

template <bool DROP>
__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_dkv_kernel(const __grid_constant__ CUtensorMap map_khi, const __grid_constant__ CUtensorMap map_klo,
                         const __grid_constant__ CUtensorMap map_vhi, const __grid_constant__ CUtensorMap map_vlo,
                         const __grid_constant__ CUtensorMap map_qhi, const __grid_constant__ CUtensorMap map_qlo,
                         const __grid_constant__ CUtensorMap map_dohi, const __grid_constant__ CUtensorMap map_dolo,
                         const __grid_constant__ CUtensorMap map_qthi, const __grid_constant__ CUtensorMap map_qtlo,
                         const __grid_constant__ CUtensorMap map_dothi, const __grid_constant__ CUtensorMap map_dotlo,
                         const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float lse_s[2][128];
  __shared__ float delta_s[2][128];
  __shared__ uint32_t key_s[2][128];
  const uint32_t raw = at_smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t s_kv = base;                       // K hi | K lo | V hi | V lo      (resident)
  const uint32_t s_qd = base + 4 * AB_TILE;         // Q hi | Q lo | dO hi | dO lo    (query tile t)
  const uint32_t s_t = base + 8 * AB_TILE;          // Q^T hi (2 boxes) | Q^T lo | dO^T hi | dO^T lo
  const uint32_t bars = s_t + 8 * AB_TBOX;
  const uint32_t bar_kv = bars, bar_rf = bars + 8, bar_re = bars + 16, bar_tf = bars + 24, bar_te = bars + 32,
                 bar_sf = bars + 40, bar_pf = bars + 48, bar_of = bars + 56;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (bars - base) + 64);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z;
  const int64_t bh = b * p.heads + h;
  const int kend = min(p.lk, p.kv_valid);
  const int T = (p.lq + 127) / 128;

  if (k0 >= kend) {                                  // keys no query attends to: zero gradients
    for (int e = threadIdx.x; e < 128 * 16; e += AB_THREADS) {
      const int r = e >> 4, c4 = (e & 15) * 4;
      if (k0 + r < p.lk) {
        const int64_t o = (b * p.lk + k0 + r) * p.ldg + h * AT_D + c4;
        *reinterpret_cast<float4*>(p.dk + o) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(p.dv + o) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    return;
  }

  if (threadIdx.x == 0) {
    at_mbar_init(bar_kv, 1);
    at_mbar_init(bar_rf, 1); at_mbar_init(bar_re, 1);
    at_mbar_init(bar_tf, 1); at_mbar_init(bar_te, 1);
    at_mbar_init(bar_sf, 1); at_mbar_init(bar_pf, 8); at_mbar_init(bar_of, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(at_smem_u32(tmem_slot)),
                 "r"(AT_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t tmem_s = tmem, tmem_dp = tmem + 128u, tmem_dv = tmem + 256u, tmem_dk = tmem + 320u;

  if (warp == 0) {
    const bool leader = at_elect_one();
    const int krow = static_cast<int>(bh * p.lk + k0);
    if (leader) at_mbar_expect_tx(bar_kv, 4 * AB_TILE);
    if (leader) at_tma_2d(s_kv, &map_khi, bar_kv, 0, krow);
    if (leader) at_tma_2d(s_kv + AB_TILE, &map_klo, bar_kv, 0, krow);
    if (leader) at_tma_2d(s_kv + 2 * AB_TILE, &map_vhi, bar_kv, 0, krow);
    if (leader) at_tma_2d(s_kv + 3 * AB_TILE, &map_vlo, bar_kv, 0, krow);
    const int trow = static_cast<int>(bh * AT_D);
    for (int t = 0; t < T; ++t) {
      const uint32_t ph = (t & 1) ^ 1u;
      const int qrow = static_cast<int>(bh * p.lq + t * 128);
      at_mbar_wait(bar_re, ph);
      if (leader) at_mbar_expect_tx(bar_rf, 4 * AB_TILE);
      if (leader) at_tma_2d(s_qd, &map_qhi, bar_rf, 0, qrow);
      if (leader) at_tma_2d(s_qd + AB_TILE, &map_qlo, bar_rf, 0, qrow);
      if (leader) at_tma_2d(s_qd + 2 * AB_TILE, &map_dohi, bar_rf, 0, qrow);
      if (leader) at_tma_2d(s_qd + 3 * AB_TILE, &map_dolo, bar_rf, 0, qrow);
      at_mbar_wait(bar_te, ph);
      if (leader) at_mbar_expect_tx(bar_tf, 8 * AB_TBOX);
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if (leader) at_tma_2d(s_t + x * AB_TBOX, &map_qthi, bar_tf, t * 128 + 64 * x, trow);
        if (leader) at_tma_2d(s_t + (2 + x) * AB_TBOX, &map_qtlo, bar_tf, t * 128 + 64 * x, trow);
        if (leader) at_tma_2d(s_t + (4 + x) * AB_TBOX, &map_dothi, bar_tf, t * 128 + 64 * x, trow);
        if (leader) at_tma_2d(s_t + (6 + x) * AB_TBOX, &map_dotlo, bar_tf, t * 128 + 64 * x, trow);
      }
    }
  } else if (warp == 1) {
    const bool leader = at_elect_one();
    const uint32_t idesc_common = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(128 >> 4) << 24);
    const uint32_t idesc_n128 = idesc_common | (static_cast<uint32_t>(128 >> 3) << 17);
    const uint32_t idesc_n64 = idesc_common | (static_cast<uint32_t>(64 >> 3) << 17);
    const uint64_t d_khi = at_desc_sw128(s_kv), d_klo = at_desc_sw128(s_kv + AB_TILE);
    const uint64_t d_vhi = at_desc_sw128(s_kv + 2 * AB_TILE), d_vlo = at_desc_sw128(s_kv + 3 * AB_TILE);
    const uint64_t d_qhi = at_desc_sw128(s_qd), d_qlo = at_desc_sw128(s_qd + AB_TILE);
    const uint64_t d_dohi = at_desc_sw128(s_qd + 2 * AB_TILE), d_dolo = at_desc_sw128(s_qd + 3 * AB_TILE);
    at_mbar_wait(bar_kv, 0);
    for (int t = 0; t < T; ++t) {
      at_mbar_wait(bar_rf, t & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int kk = 0; kk < AT_D / 16; ++kk) {                   // S^T = K Q_t^T
        const uint64_t adv = static_cast<uint64_t>(kk * 2);
        if (leader) at_umma_bf16(tmem_s, d_klo + adv, d_qhi + adv, idesc_n128, kk != 0 ? 1u : 0u);
        if (leader) at_umma_bf16(tmem_s, d_khi + adv, d_qlo + adv, idesc_n128, 1u);
        if (leader) at_umma_bf16(tmem_s, d_khi + adv, d_qhi + adv, idesc_n128, 1u);
      }
#pragma unroll
      for (int kk = 0; kk < AT_D / 16; ++kk) {                   // dP^T = V dO_t^T
        const uint64_t adv = static_cast<uint64_t>(kk * 2);
        if (leader) at_umma_bf16(tmem_dp, d_vlo + adv, d_dohi + adv, idesc_n128, kk != 0 ? 1u : 0u);
        if (leader) at_umma_bf16(tmem_dp, d_vhi + adv, d_dolo + adv, idesc_n128, 1u);
        if (leader) at_umma_bf16(tmem_dp, d_vhi + adv, d_dohi + adv, idesc_n128, 1u);
      }
      if (leader) at_commit(bar_sf);
      if (leader) at_commit(bar_re);                             // Q / dO row tiles consumed
      at_mbar_wait(bar_pf, t & 1);                               // Pd^T, dS^T written
      at_mbar_wait(bar_tf, t & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {                           // 16 queries per step
        const uint32_t off = static_cast<uint32_t>((kk >> 1) * 32 + (kk & 1) * 8);
        const uint32_t box = static_cast<uint32_t>(kk >> 2) * AB_TBOX;
        const uint64_t adv = static_cast<uint64_t>((kk & 3) * 2);
        const uint64_t d_qthi = at_desc_sw128(s_t + box) + adv, d_qtlo = at_desc_sw128(s_t + 2 * AB_TBOX + box) + adv;
        const uint64_t d_dthi = at_desc_sw128(s_t + 4 * AB_TBOX + box) + adv,
                       d_dtlo = at_desc_sw128(s_t + 6 * AB_TBOX + box) + adv;
        const uint32_t acc = (t | kk) != 0 ? 1u : 0u;
        // dV += Pd^T dO_t
        if (leader) at_umma_bf16_ts(tmem_dv, tmem_s + off + 16, d_dthi, idesc_n64, acc);
        if (leader) at_umma_bf16_ts(tmem_dv, tmem_s + off, d_dtlo, idesc_n64, 1u);
        if (leader) at_umma_bf16_ts(tmem_dv, tmem_s + off, d_dthi, idesc_n64, 1u);
        // dK += dS^T Q_t
        if (leader) at_umma_bf16_ts(tmem_dk, tmem_dp + off + 16, d_qthi, idesc_n64, acc);
        if (leader) at_umma_bf16_ts(tmem_dk, tmem_dp + off, d_qtlo, idesc_n64, 1u);
        if (leader) at_umma_bf16_ts(tmem_dk, tmem_dp + off, d_qthi, idesc_n64, 1u);
      }
      if (leader) at_commit(bar_te);
      if (t + 1 == T && leader) at_commit(bar_of);
    }
  } else {
    const int qd = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = qd * 32 + lane;                      // key row of the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const int kj = k0 + r;
    const bool key_ok = kj < kend;
    const int st = static_cast<int>(threadIdx.x) - 64;           // 0..255 among the element-wise warps
    for (int t = 0; t < T; ++t) {
      // per-query (column) terms of this tile; double-buffered: a warp is at most one tile ahead of the slowest one
      if (st < 128) {
        const int qi = t * 128 + st;
        const bool ok = qi < p.lq;
        lse_s[t & 1][st] = ok ? __ldg(p.lse + bh * p.lq + qi) : INFINITY;      // absent query: p = 0
        delta_s[t & 1][st] = ok ? __ldg(p.delta + bh * p.lq + qi) : 0.f;
        if (DROP) key_s[t & 1][st] = dropout_row_key(p.seed, static_cast<uint64_t>(bh) * p.lq + static_cast<uint64_t>(qi));
      }
      at_named_barrier(1, 256);
      at_mbar_wait(bar_sf, t & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;
        uint32_t sv[32], dp[32], phi[16], plo[16], dhi[16], dlo[16];
        at_tmem_ld32(tmem_s + lane_addr + c * 32, sv);
        at_tmem_ld32(tmem_dp + lane_addr + c * 32, dp);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float pd[2], ds[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int col = c * 32 + 2 * e + u;
            float pv = ex2_approx(__uint_as_float(sv[2 * e + u]) - lse_s[t & 1][col]);
            float g = __uint_as_float(dp[2 * e + u]);
            pd[u] = pv;
            if (DROP) {
              const bool keep = dropout_keep(key_s[t & 1][col], static_cast<uint32_t>(kj), p.drop_threshold);
              g = keep ? g * p.keep_scale : 0.f;
              pd[u] = keep ? pv * p.keep_scale : 0.f;
            }
            ds[u] = pv * (g - delta_s[t & 1][col]);
            if (!key_ok) { pd[u] = 0.f; ds[u] = 0.f; }
          }
          split_bf16x2(pd[0], pd[1], phi[e], plo[e]);
          split_bf16x2(ds[0], ds[1], dhi[e], dlo[e]);
        }
        at_tmem_st16(tmem_s + lane_addr + c * 32, phi);
        at_tmem_st16(tmem_s + lane_addr + c * 32 + 16, plo);
        at_tmem_st16(tmem_dp + lane_addr + c * 32, dhi);
        at_tmem_st16(tmem_dp + lane_addr + c * 32 + 16, dlo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) at_mbar_arrive(bar_pf);
    }
    at_mbar_wait(bar_of, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t ov[32], ok_[32];
    at_tmem_ld32(tmem_dv + lane_addr + half * 32, ov);
    at_tmem_ld32(tmem_dk + lane_addr + half * 32, ok_);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (kj < p.lk) {
      const int64_t o = (b * p.lk + kj) * p.ldg + h * AT_D + half * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        *reinterpret_cast<float4*>(p.dv + o + j) = make_float4(__uint_as_float(ov[j]), __uint_as_float(ov[j + 1]),
                                                               __uint_as_float(ov[j + 2]), __uint_as_float(ov[j + 3]));
        *reinterpret_cast<float4*>(p.dk + o + j) =
            make_float4(__uint_as_float(ok_[j]) * 0.125f, __uint_as_float(ok_[j + 1]) * 0.125f,
                        __uint_as_float(ok_[j + 2]) * 0.125f, __uint_as_float(ok_[j + 3]) * 0.125f);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(AT_TMEM_COLS) : "memory");
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 at_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// bf16 row-major (rows, cols) with pitch `ld` elements; box = 64 cols x box_rows, 128-byte swizzle
static bool at_make_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  auto enc = at_encode_fn();
  if (enc == nullptr) return false;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int64_t attention_tc_workspace_bytes(int64_t batch, int64_t heads, int64_t lq, int64_t lk) {
  const int64_t lk_pad = (lk + 7) / 8 * 8;
  const int64_t q = batch * heads * lq * AT_D * 2, k = batch * heads * lk * AT_D * 2,
                v = batch * heads * AT_D * lk_pad * 2;
  auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
  return 2 * (up(q) + up(k) + up(v));
}

int launch_attention_tc(const float* q, int64_t ldq, const float* k, const float* v, int64_t ldk, float* out,
                        int64_t ldo, int64_t batch, int64_t heads, int64_t lq, int64_t lk, int64_t kv_valid,
                        void* workspace, cudaStream_t s, uint16_t* out_hi, uint16_t* out_lo, float p_drop, uint64_t seed, float* lse) {
  if (batch * heads * (lq > lk ? lq : lk) > 0x7fffff00LL) return HOISDF_E_SHAPE;  // TMA row coordinates are int32
  const int64_t lk_pad = (lk + 7) / 8 * 8;
  auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
  uint8_t* w = static_cast<uint8_t*>(workspace);
  const int64_t qb = up(batch * heads * lq * AT_D * 2), kb = up(batch * heads * lk * AT_D * 2),
                vb = up(batch * heads * AT_D * lk_pad * 2);
  __nv_bfloat16 *qhi = reinterpret_cast<__nv_bfloat16*>(w), *qlo = reinterpret_cast<__nv_bfloat16*>(w + qb);
  __nv_bfloat16 *khi = reinterpret_cast<__nv_bfloat16*>(w + 2 * qb), *klo = reinterpret_cast<__nv_bfloat16*>(w + 2 * qb + kb);
  __nv_bfloat16 *vhi = reinterpret_cast<__nv_bfloat16*>(w + 2 * qb + 2 * kb),
                *vlo = reinterpret_cast<__nv_bfloat16*>(w + 2 * qb + 2 * kb + vb);
  {
    const int64_t nq = batch * lq * heads * 16, nk = batch * lk * heads * 16;
    attn_split_rows_kernel<<<static_cast<unsigned>(ceil_div(nq, 256)), 256, 0, s>>>(q, ldq, batch, static_cast<int>(heads),
                                                                                  lq, 0.125f * 1.4426950408889634f, qhi, qlo);   // scores in log2 units
    attn_split_rows_kernel<<<static_cast<unsigned>(ceil_div(nk, 256)), 256, 0, s>>>(k, ldk, batch, static_cast<int>(heads),
                                                                                  lk, 1.0f, khi, klo);
    dim3 g(static_cast<unsigned>(ceil_div(lk, 64)), static_cast<unsigned>(heads), static_cast<unsigned>(batch));
    attn_split_vt_kernel<<<g, 256, 0, s>>>(v, ldk, static_cast<int>(heads), lk, lk_pad, vhi, vlo);
  }
  CUtensorMap mqh, mql, mkh, mkl, mvh, mvl;
  const int64_t bh = batch * heads;
  // 128-key tiles (attention_tc128_kernel) for the long sequences; HOISDF_ATTN_BK=64 keeps the 64-key kernel
  static const bool allow_wide = [] { const char* e = getenv("HOISDF_ATTN_BK"); return e == nullptr || atoi(e) != 64; }();
  const bool wide = allow_wide && (kv_valid < lk ? kv_valid : lk) >= A2_BK;
  const int kbox = wide ? A2_BK : AT_BK;
  if (!at_make_map(&mqh, qhi, bh * lq, AT_D, AT_D, AT_BQ) || !at_make_map(&mql, qlo, bh * lq, AT_D, AT_D, AT_BQ) ||
      !at_make_map(&mkh, khi, bh * lk, AT_D, AT_D, kbox) || !at_make_map(&mkl, klo, bh * lk, AT_D, AT_D, kbox) ||
      !at_make_map(&mvh, vhi, bh * AT_D, lk_pad, lk_pad, AT_D) || !at_make_map(&mvl, vlo, bh * AT_D, lk_pad, lk_pad, AT_D))
    return HOISDF_E_UNSUPPORTED;
  if ((p_drop > 0.f || lse != nullptr) && !wide) return HOISDF_E_UNSUPPORTED;      // training: the 128-key kernel only
  AttnTcParams p{out, out_hi, out_lo, ldo, static_cast<int>(lq), static_cast<int>(lk), static_cast<int>(kv_valid),
                 static_cast<int>(heads), seed, dropout_threshold(p_drop), 1.0f / (1.0f - p_drop), lse};
  dim3 grid(static_cast<unsigned>(ceil_div(lq, AT_BQ * AT_GROUPS)), static_cast<unsigned>(heads), static_cast<unsigned>(batch));
  if (wide) {
    auto kernel = p_drop > 0.f ? attention_tc128_kernel<true> : attention_tc128_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A2_SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    kernel<<<grid, AT_THREADS, A2_SMEM_BYTES, s>>>(mqh, mql, mkh, mkl, mvh, mvl, p);
    return launch_status();
  }
  cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES);
  if (e != cudaSuccess) return static_cast<int>(e);
  attention_tc_kernel<<<grid, AT_THREADS, AT_SMEM_BYTES, s>>>(mqh, mql, mkh, mkl, mvh, mvl, p);
  return launch_status();
}


int64_t attention_bwd_tc_workspace_bytes(int64_t batch, int64_t heads, int64_t lq, int64_t lk) {
  const int64_t lq_pad = (lq + 7) / 8 * 8, lk_pad = (lk + 7) / 8 * 8, bh = batch * heads;
  auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
  const int64_t qb = up(bh * lq * AT_D * 2), kb = up(bh * lk * AT_D * 2), qtb = up(bh * AT_D * lq_pad * 2),
                ktb = up(bh * AT_D * lk_pad * 2);
  return 4 * qb + 4 * kb + 4 * qtb + 2 * ktb + up(bh * lq * 4);
}

int launch_attention_bwd_tc(const float* q, int64_t ldq, const float* k, const float* v, int64_t ldk, const float* out,
                            const float* dout, int64_t ldo, const float* lse, float* dq, float* dk, float* dv, int64_t ldg,
                            int64_t batch, int64_t heads, int64_t lq, int64_t lk, int64_t kv_valid, float p_drop,
                            uint64_t seed, void* workspace, cudaStream_t s) {
  if (batch * heads * (lq > lk ? lq : lk) > 0x7fffff00LL) return HOISDF_E_SHAPE;  // TMA row coordinates are int32
  const int64_t lq_pad = (lq + 7) / 8 * 8, lk_pad = (lk + 7) / 8 * 8, bh = batch * heads;
  auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
  const int64_t qb = up(bh * lq * AT_D * 2), kb = up(bh * lk * AT_D * 2), qtb = up(bh * AT_D * lq_pad * 2),
                ktb = up(bh * AT_D * lk_pad * 2);
  uint8_t* w = static_cast<uint8_t*>(workspace);
  auto take = [&](int64_t bytes) { auto* r = reinterpret_cast<__nv_bfloat16*>(w); w += bytes; return r; };
  __nv_bfloat16 *qhi = take(qb), *qlo = take(qb), *dohi = take(qb), *dolo = take(qb);
  __nv_bfloat16 *khi = take(kb), *klo = take(kb), *vhi = take(kb), *vlo = take(kb);
  __nv_bfloat16 *qthi = take(qtb), *qtlo = take(qtb), *dothi = take(qtb), *dotlo = take(qtb);
  __nv_bfloat16 *kthi = take(ktb), *ktlo = take(ktb);
  float* delta = reinterpret_cast<float*>(w);
  const int H = static_cast<int>(heads);
  {
    const int64_t nw = batch * lq * heads;                               // one warp each
    attn_bwd_delta_kernel<<<static_cast<unsigned>(ceil_div(nw * 32, 256)), 256, 0, s>>>(dout, out, ldo, batch, H, lq, delta);
    const int64_t nq = batch * lq * heads * 16, nk = batch * lk * heads * 16;
    const unsigned gq = static_cast<unsigned>(ceil_div(nq, 256)), gk = static_cast<unsigned>(ceil_div(nk, 256));
    attn_split_rows_kernel<<<gq, 256, 0, s>>>(q, ldq, batch, H, lq, 0.125f * 1.4426950408889634f, qhi, qlo);   // as the forward
    attn_split_rows_kernel<<<gq, 256, 0, s>>>(dout, ldo, batch, H, lq, 1.0f, dohi, dolo);
    attn_split_rows_kernel<<<gk, 256, 0, s>>>(k, ldk, batch, H, lk, 1.0f, khi, klo);
    attn_split_rows_kernel<<<gk, 256, 0, s>>>(v, ldk, batch, H, lk, 1.0f, vhi, vlo);
    const dim3 tq(static_cast<unsigned>(ceil_div(lq, 64)), static_cast<unsigned>(heads), static_cast<unsigned>(batch));
    const dim3 tk(static_cast<unsigned>(ceil_div(lk, 64)), static_cast<unsigned>(heads), static_cast<unsigned>(batch));
    attn_split_vt_kernel<<<tq, 256, 0, s>>>(q, ldq, H, lq, lq_pad, qthi, qtlo);
    attn_split_vt_kernel<<<tq, 256, 0, s>>>(dout, ldo, H, lq, lq_pad, dothi, dotlo);
    attn_split_vt_kernel<<<tk, 256, 0, s>>>(k, ldk, H, lk, lk_pad, kthi, ktlo);
  }
  CUtensorMap mqh, mql, mdh, mdl, mkh, mkl, mvh, mvl, mqth, mqtl, mdth, mdtl, mkth, mktl;
  if (!at_make_map(&mqh, qhi, bh * lq, AT_D, AT_D, 128) || !at_make_map(&mql, qlo, bh * lq, AT_D, AT_D, 128) ||
      !at_make_map(&mdh, dohi, bh * lq, AT_D, AT_D, 128) || !at_make_map(&mdl, dolo, bh * lq, AT_D, AT_D, 128) ||
      !at_make_map(&mkh, khi, bh * lk, AT_D, AT_D, 128) || !at_make_map(&mkl, klo, bh * lk, AT_D, AT_D, 128) ||
      !at_make_map(&mvh, vhi, bh * lk, AT_D, AT_D, 128) || !at_make_map(&mvl, vlo, bh * lk, AT_D, AT_D, 128) ||
      !at_make_map(&mqth, qthi, bh * AT_D, lq_pad, lq_pad, AT_D) || !at_make_map(&mqtl, qtlo, bh * AT_D, lq_pad, lq_pad, AT_D) ||
      !at_make_map(&mdth, dothi, bh * AT_D, lq_pad, lq_pad, AT_D) ||
      !at_make_map(&mdtl, dotlo, bh * AT_D, lq_pad, lq_pad, AT_D) ||
      !at_make_map(&mkth, kthi, bh * AT_D, lk_pad, lk_pad, AT_D) || !at_make_map(&mktl, ktlo, bh * AT_D, lk_pad, lk_pad, AT_D))
    return HOISDF_E_UNSUPPORTED;
  AttnBwdParams p{lse, delta, dq, dk, dv, ldg, static_cast<int>(lq), static_cast<int>(lk), static_cast<int>(kv_valid),
                  H, seed, dropout_threshold(p_drop), 1.0f / (1.0f - p_drop)};
  const bool drop = p_drop > 0.f;
  {
    auto kernel = drop ? attention_bwd_dq_kernel<true> : attention_bwd_dq_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_DQ_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    const dim3 grid(static_cast<unsigned>(ceil_div(lq, 128)), static_cast<unsigned>(heads), static_cast<unsigned>(batch));
    kernel<<<grid, AB_THREADS, AB_DQ_SMEM, s>>>(mqh, mql, mdh, mdl, mkh, mkl, mvh, mvl, mkth, mktl, p);
  }
  {
    auto kernel = drop ? attention_bwd_dkv_kernel<true> : attention_bwd_dkv_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_DKV_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    const dim3 grid(static_cast<unsigned>(ceil_div(lk, 128)), static_cast<unsigned>(heads), static_cast<unsigned>(batch));
    kernel<<<grid, AB_THREADS, AB_DKV_SMEM, s>>>(mkh, mkl, mvh, mvl, mqh, mql, mdh, mdl, mqth, mqtl, mdth, mdtl, p);
  }
  return launch_status();
}

}  // namespace hoisdf
