// nn.Linear on the 5th-generation tensor cores with fp32-grade accuracy: Y = act(X . W^T + b) via 3xTF32.
//
//   x = x_hi + x_lo  (x_hi = RNA-rounded TF32, x_lo = x - x_hi exactly), same for W (split once at pack time)
//   X.W^T ~= x_hi.w_hi + x_lo.w_hi + x_hi.w_lo     (the dropped x_lo.w_lo term is < 2^-22 relative)
// accumulated in fp32 in TMEM.  Measured error is indistinguishable from an fp32 FMA GEMM (SURVEY.md section 7
// precision probe), which the near-surface top-k selection needs; a single TF32 or BF16 pass is not enough.
//
// sm_100a structure (one CTA per 128 x 256 output tile, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D tiles (64B swizzle, 16 floats of K per stage) of X, W_hi,
//               W_lo -> 4-stage smem ring (the ring depth hides the ~1 us TMA latency)
//   warp 1      tcgen05.mma issuer (one thread): 3 MMAs (kind::tf32, M128 x N<=256 x K8) per 32-byte K slice.
//               Two accumulators in TMEM (all 512 columns): `main` takes x_hi.w_hi, `corr` takes the two small
//               cross terms.  The tensor core truncates when it adds into the accumulator, so keeping the small
//               terms out of the large sum cuts the truncation events on it 3x (measured: error drops ~3x).
//               tcgen05.commit releases smem stages / signals the epilogue.
//   warps 2..5  per stage: split the raw X tile in place into x_hi / x_lo (elementwise, so the swizzled
//               layout never has to be decoded), then the epilogue: tcgen05.ld 32 lanes x 32 columns per
//               warp of both accumulators, fused bias (staged in smem) / residual / ReLU, 128-bit global stores.
// Tails need no padding: TMA zero-fills rows >= M / >= N and columns >= K.
//
// Thread-block clusters: the kernel is bound by L2 -> SM traffic when every CTA streams the whole W tile (hi + lo =
// 2 x 256 rows) next to its 128 X rows (25.6 FLOP per fetched byte).  CTAs are therefore launched as clusters of 2
// that own two M tiles of the SAME N tile: each CTA fetches half of the W rows and TMA-multicasts them into both
// CTAs' shared memory, so W crosses the L2 -> SM fabric once per pair (42.7 FLOP/B).  A stage is released to the
// producers only when BOTH CTAs' tensor cores have consumed it (tcgen05.commit multicast to both `empty` barriers).
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"

namespace hoisdf {

constexpr int TC_BM = 128;
constexpr int TC_BN = 256;
constexpr int TC_BK = 16;                       // floats per stage row = one 64-byte swizzle row
constexpr int TC_STAGES = 4;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;   // 8 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;   // 16 KB
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;  // x_hi, x_lo, w_hi, w_lo = 48 KB
constexpr int TC_BAR_BYTES = 256;
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + TC_BAR_BYTES + TC_BN * 4 /*bias*/;
constexpr int TC_THREADS = 192;
constexpr uint32_t TC_TMEM_COLS = 512;          // main accumulator [0,256) + correction accumulator [256,512)
constexpr uint32_t kSpinLimit = 1u << 27;       // watchdog: a protocol bug traps instead of hanging the GPU
constexpr int TC_CLUSTER = 2;                   // CTAs per cluster (share the W tile through TMA multicast)
constexpr int TC_W_SLICE_ROWS = TC_BN / TC_CLUSTER;          // W rows each CTA fetches (and multicasts)
constexpr int TC_W_SLICE_BYTES = TC_W_SLICE_ROWS * TC_BK * 4;

struct TcParams {
  const float* __restrict__ bias;
  const float* __restrict__ residual;
  float* __restrict__ y;
  int64_t ldy;
  int64_t rows_per_batch;   // X rows form groups of this many rows (plain GEMM: = M, one group)
  int tiles_per_batch;      // ceil(rows_per_batch / 128)
  int m_tiles;              // groups * tiles_per_batch (CTAs beyond it are cluster padding)
  int n, k, act;
  int passes;               // 3 = 3xTF32 (fp32-grade), 1 = single TF32 pass (screening only)
  int tma_store;            // 1: epilogue stages 32x32 boxes in smem and writes them with TMA (no residual, dense rows)
  unsigned long long* trace; // developer timeline (globaltimer ns) of CTA `trace_cta`; NULL in production
  int trace_cta;
};

__device__ __forceinline__ void tc_trace(const TcParams& p, int slot) {
  if (p.trace != nullptr && static_cast<int>(blockIdx.x) == p.trace_cta) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.trace[slot] = t;
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
    if (++spins > kSpinLimit) __trap();
  }
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// K-major, 64-byte swizzle (rows of 16 floats): 8-row groups are 512 B apart (SBO), LBO unused (=1),
// descriptor version 1 (sm_100), layout type 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (static_cast<uint64_t>(512 >> 4) << 32) |
         (1ull << 46) | (4ull << 61);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(cta_mask)
      : "memory");
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
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

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
linear_tf32x3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_whi,
                     const __grid_constant__ CUtensorMap map_wlo, const __grid_constant__ CUtensorMap map_y,
                     const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;           // swizzled tiles: keep every tile 1024-byte aligned
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bars = base + TC_STAGES * TC_STAGE_BYTES;  // full[S] conv[S] empty[S] done | tmem ptr | bias
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + TC_STAGES * TC_STAGE_BYTES + 8 * (3 * TC_STAGES + 1));
  float* bias_s = reinterpret_cast<float*>(gen + TC_STAGES * TC_STAGE_BYTES + TC_BAR_BYTES);
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_conv = [&](int s) { return bars + 8u * (TC_STAGES + s); };
  auto bar_empty = [&](int s) { return bars + 8u * (2 * TC_STAGES + s); };
  const uint32_t bar_done = bars + 8u * (3 * TC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.n + TC_BN - 1) / TC_BN;
  uint32_t cta_rank;
  asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const int cid = static_cast<int>(blockIdx.x) / TC_CLUSTER;          // cluster id: (M-tile pair, N tile), N fastest
  const int m_tile = (cid / n_tiles) * TC_CLUSTER + static_cast<int>(cta_rank);
  // cluster padding (odd number of M tiles): group index = #groups, every X row out of bounds -> zero-filled
  const int grp = m_tile < p.m_tiles ? m_tile / p.tiles_per_batch : p.m_tiles / p.tiles_per_batch;
  const int m0 = m_tile < p.m_tiles ? (m_tile - grp * p.tiles_per_batch) * TC_BM : 0;
  const int n0 = (cid % n_tiles) * TC_BN;
  const int n_here = min(TC_BN, p.n - n0);
  const int n_inst = (n_here + 15) & ~15;                  // UMMA N (multiple of 16 for M = 128)
  const int num_kb = (p.k + TC_BK - 1) / TC_BK;

  if (threadIdx.x == 0) tc_trace(p, 0);
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_whi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wlo) : "memory");
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_conv(s), 4);
      mbar_init(bar_empty(s), TC_CLUSTER);   // both CTAs' tensor cores must have consumed the stage
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                        // the peer's barriers are initialised before any multicast lands there
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = *tmem_slot;
  constexpr uint16_t kAllCtas = (1u << TC_CLUSTER) - 1u;
  if (threadIdx.x == 0) tc_trace(p, 1);

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (kb / TC_STAGES) & 1;
        mbar_wait(bar_empty(s), ph ^ 1u);
        if (kb < 40) tc_trace(p, 8 + kb);                       // producer: stage free
        const uint32_t st = base + s * TC_STAGE_BYTES;
        mbar_expect_tx(bar_full(s), TC_A_BYTES + (p.passes == 3 ? 2 : 1) * TC_B_BYTES);
        tma_load_3d(st, &map_x, bar_full(s), kb * TC_BK, m0, grp);
        // this CTA's slice of the W rows, multicast to every CTA of the cluster (same smem offset, same barrier)
        const uint32_t wo = cta_rank * TC_W_SLICE_BYTES;
        const int wrow = n0 + static_cast<int>(cta_rank) * TC_W_SLICE_ROWS;
        tma_load_2d_mc(st + 2 * TC_A_BYTES + wo, &map_whi, bar_full(s), kb * TC_BK, wrow, kAllCtas);
        if (p.passes == 3)
          tma_load_2d_mc(st + 2 * TC_A_BYTES + TC_B_BYTES + wo, &map_wlo, bar_full(s), kb * TC_BK, wrow, kAllCtas);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n_inst >> 3) << 17) |
                             (static_cast<uint32_t>(TC_BM >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (kb / TC_STAGES) & 1;
        mbar_wait(bar_full(s), ph);
        if (kb < 40) tc_trace(p, 48 + kb);                      // MMA: TMA bytes landed
        if (p.passes == 3) mbar_wait(bar_conv(s), ph);
        if (kb < 40) tc_trace(p, 88 + kb);                      // MMA: x split done
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = base + s * TC_STAGE_BYTES;
        const uint64_t d_xhi = umma_desc_sw64(st), d_xlo = umma_desc_sw64(st + TC_A_BYTES);
        const uint64_t d_whi = umma_desc_sw64(st + 2 * TC_A_BYTES);
        const uint64_t d_wlo = umma_desc_sw64(st + 2 * TC_A_BYTES + TC_B_BYTES);
#pragma unroll
        for (int kk = 0; kk < TC_BK / 8; ++kk) {
          const uint64_t adv = static_cast<uint64_t>(kk * 2);  // 8 tf32 = 32 bytes = 2 x 16-byte units
          const uint32_t acc = (kb | kk) != 0 ? 1u : 0u;
          umma_tf32(tmem_acc, d_xhi + adv, d_whi + adv, idesc, acc);   // single pass: the hardware truncates raw x
          if (p.passes == 3) {
            umma_tf32(tmem_acc + TC_BN, d_xlo + adv, d_whi + adv, idesc, acc);
            umma_tf32(tmem_acc + TC_BN, d_xhi + adv, d_wlo + adv, idesc, 1u);
          }
        }
        umma_commit_mc(bar_empty(s), kAllCtas);   // stage may be refilled once BOTH CTAs' MMAs have read it
        if (kb < 40) tc_trace(p, 128 + kb);                     // MMA: issued
      }
      umma_commit(bar_done);         // accumulator complete
    }
  } else {
    // ------------------------------------------------ converter (mainloop), then epilogue
    const int t = threadIdx.x - 64;  // 0..127
    for (int c = t; c < TC_BN; c += 128) bias_s[c] = (p.bias != nullptr && c < n_here) ? __ldg(p.bias + n0 + c) : 0.f;
    for (int kb = 0; kb < (p.passes == 3 ? num_kb : 0); ++kb) {
      const int s = kb % TC_STAGES;
      const uint32_t ph = (kb / TC_STAGES) & 1;
      mbar_wait(bar_full(s), ph);
      float4* hi = reinterpret_cast<float4*>(gen + s * TC_STAGE_BYTES);
      float4* lo = reinterpret_cast<float4*>(gen + s * TC_STAGE_BYTES + TC_A_BYTES);
#pragma unroll
      for (int j = 0; j < TC_A_BYTES / 16 / 128; ++j) {
        const int i = t + 128 * j;
        const float4 v = hi[i];
        float4 h, l;
        h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
        l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
        hi[i] = h;
        lo[i] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_conv(s));
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");   // bias_s visible to all four epilogue warps
    mbar_wait(bar_done, 0);
    if (threadIdx.x == 64) tc_trace(p, 2);                       // accumulator complete
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;                        // TMEM lane quarter this warp may read
    const int64_t lrow = static_cast<int64_t>(m0) + q * 32 + lane;      // row inside the group
    const bool row_ok = lrow < p.rows_per_batch && m_tile < p.m_tiles;
    const int64_t row = static_cast<int64_t>(grp) * p.rows_per_batch + lrow;   // output rows are dense
    float* yrow = p.y + (row_ok ? row * p.ldy : 0) + n0;
    const float* rrow = p.residual ? p.residual + (row_ok ? row * p.ldy : 0) + n0 : nullptr;
    const bool vec = ((p.ldy & 3) == 0) && aligned16(p.y) && (p.residual == nullptr || aligned16(p.residual));
    if (p.tma_store) {
      // Row-strided 16-byte stores cost one LSU sector operation per lane (the old epilogue spent ~4 us per tile in
      // them).  Instead each warp stages its 32 rows x 32 columns as a 128B-swizzled box in the (now idle) operand
      // ring and lets the TMA write it: full-line writes, rows >= M and columns >= N clipped by the tensor map.
      uint8_t* stage_gen = gen + q * 32768;
      const uint32_t stage_sh = base + q * 32768;
      const int64_t row0 = static_cast<int64_t>(grp) * p.rows_per_batch + m0 + q * 32;
      for (int c0 = 0; c0 < n_inst; c0 += 32) {
        uint32_t r[32], rc[32];
        const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c0);
        tmem_ld32(taddr, r);
        if (p.passes == 3) {
          tmem_ld32(taddr + TC_BN, rc);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) rc[j] = 0u;
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint8_t* box = stage_gen + (c0 >> 5) * 4096;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[j] = (__uint_as_float(r[g * 4 + j]) + __uint_as_float(rc[g * 4 + j])) + bias_s[c0 + g * 4 + j];
            if (p.act == HOISDF_ACT_RELU) v[j] = fmaxf(v[j], 0.f);
          }
          *reinterpret_cast<float4*>(box + lane * 128 + ((g ^ (lane & 7)) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && m_tile < p.m_tiles && c0 < n_here) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(&map_y), "r"(stage_sh + (c0 >> 5) * 4096), "r"(n0 + c0), "r"(static_cast<int>(row0))
                       : "memory");
        }
      }
      if (lane == 0) {
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem source consumed; kernel end flushes the writes
      }
      __syncwarp();
    } else
    for (int c0 = 0; c0 < n_inst; c0 += 32) {
      uint32_t r[32], rc[32];
      const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c0);
      tmem_ld32(taddr, r);
      if (p.passes == 3) {
        tmem_ld32(taddr + TC_BN, rc);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) rc[j] = 0u;
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (!row_ok) continue;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int c = c0 + g * 4;
        if (c >= n_here) break;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j] = (__uint_as_float(r[g * 4 + j]) + __uint_as_float(rc[g * 4 + j])) + bias_s[c + j];
        }
        if (vec && c + 3 < n_here) {
          if (rrow != nullptr) {
            const float4 rr = *reinterpret_cast<const float4*>(rrow + c);
            v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
          }
          if (p.act == HOISDF_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          *reinterpret_cast<float4*>(yrow + c) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (c + j < n_here) {
              float o = v[j];
              if (rrow != nullptr) o += rrow[c + j];
              if (p.act == HOISDF_ACT_RELU) o = fmaxf(o, 0.f);
              yrow[c + j] = o;
            }
          }
        }
      }
    }
  }
  if (threadIdx.x == 64) tc_trace(p, 3);                         // epilogue stores issued
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                        // no CTA exits while the peer may still signal its barriers
  if (threadIdx.x == 0) tc_trace(p, 4);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(TC_TMEM_COLS) : "memory");
  }
}

// W -> (W_hi = RNA-rounded TF32, W_lo = W - W_hi), elementwise over a (rows, ld) matrix
__global__ void split_tf32_kernel(const float* __restrict__ w, int64_t n, float* __restrict__ hi,
                                  float* __restrict__ lo) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = w[i];
  const float h = tf32_rna(v);
  hi[i] = h;
  lo[i] = v - h;
}

// developer hook (not part of the public header): timeline of one CTA of the next launches
static unsigned long long* g_tc_trace = nullptr;
static int g_tc_trace_cta = 0;

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// 2-D fp32 row-major (rows, ld) view with `cols` valid columns; box = box_rows x 16 floats, 64-byte swizzle
static bool make_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  auto enc = encode_fn();
  if (enc == nullptr) return false;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {TC_BK, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// X as (groups, rows_per_group, K): box = 1 x 128 rows x 16 floats
static bool make_map_x(CUtensorMap* map, const float* ptr, int64_t groups, int64_t rows, int64_t cols, int64_t ld,
                       int64_t group_stride) {
  auto enc = encode_fn();
  if (enc == nullptr) return false;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(groups)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 4, static_cast<cuuint64_t>(group_stride) * 4};
  cuuint32_t box[3] = {TC_BK, TC_BM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Y as (rows, n) with pitch ldy: box = 32 rows x 32 floats, 128-byte swizzle (TMA-store epilogue)
static bool make_map_y(CUtensorMap* map, float* ptr, int64_t rows, int64_t cols, int64_t ld) {
  auto enc = encode_fn();
  if (enc == nullptr) return false;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// called by hoisdf_linear_fwd when args->w_lo is set and the OUTPUT rows are dense (input rows may be batched)
int launch_linear_tf32x3(const hoisdf_linear_args* a, cudaStream_t s) {
  CUtensorMap mx, mhi, mlo;
  int64_t groups = 1, rpb = a->m, gstride = a->m * a->ldx;
  if (a->x_rows_per_batch > 0) {
    if (a->m % a->x_rows_per_batch != 0) return HOISDF_E_SHAPE;
    rpb = a->x_rows_per_batch;
    groups = a->m / rpb;
    gstride = a->x_batch_stride;
  }
  if (groups > 1 && ((gstride * 4) % 16 != 0)) return HOISDF_E_ALIGN;
  if (groups == 1) gstride = rpb * a->ldx;   // unused by the hardware for a single group, but must be valid
  if (!make_map_x(&mx, a->x, groups, rpb, a->k, a->ldx, gstride)) return HOISDF_E_UNSUPPORTED;
  if (!make_map(&mhi, a->w, a->n, a->k, a->ldw, TC_W_SLICE_ROWS)) return HOISDF_E_UNSUPPORTED;
  if (!make_map(&mlo, a->w_lo, a->n, a->k, a->ldw, TC_W_SLICE_ROWS)) return HOISDF_E_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(linear_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  if (e != cudaSuccess) return static_cast<int>(e);
  // TMA-store epilogue: no residual, 16-byte aligned pitch, and tiles never straddle two row groups
  const bool tma_store = a->residual == nullptr && (a->ldy % 4 == 0) && aligned16(a->y) &&
                         (groups == 1 || rpb % TC_BM == 0) && a->m < 0x7fffffffLL;
  CUtensorMap my = mx;   // placeholder when the TMA-store epilogue is not used
  if (tma_store && !make_map_y(&my, a->y, a->m, a->n, a->ldy)) return HOISDF_E_UNSUPPORTED;
  const int64_t tpb = ceil_div(rpb, TC_BM);
  const int64_t m_tiles = groups * tpb;
  TcParams p{a->bias, a->residual, a->y, a->ldy, rpb, static_cast<int>(tpb), static_cast<int>(m_tiles),
             static_cast<int>(a->n), static_cast<int>(a->k), a->act, a->tf32_passes == 1 ? 1 : 3, tma_store ? 1 : 0,
             g_tc_trace, g_tc_trace_cta};
  const int64_t ctas = ceil_div(m_tiles, TC_CLUSTER) * TC_CLUSTER * ceil_div(a->n, TC_BN);
  if (ctas > 0x7fffffffLL || m_tiles > 0x3fffffffLL) return HOISDF_E_SHAPE;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(ctas));
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TC_SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = TC_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, linear_tf32x3_kernel, mx, mhi, mlo, my, p);
  if (e != cudaSuccess) return static_cast<int>(e);
  return launch_status();
}

}  // namespace hoisdf

using namespace hoisdf;

extern "C" __attribute__((visibility("default"))) void hoisdf_debug_tc_trace(unsigned long long* buf, int cta) {
  g_tc_trace = buf;
  g_tc_trace_cta = cta;
}

HOISDF_API int hoisdf_split_tf32(const float* w, int64_t count, float* w_hi, float* w_lo, void* stream) {
  if (w == nullptr || w_hi == nullptr || w_lo == nullptr) return HOISDF_E_NULL;
  if (count <= 0) return HOISDF_E_SHAPE;
  split_tf32_kernel<<<static_cast<unsigned>(ceil_div(count, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, count, w_hi, w_lo);
  return launch_status();
}
