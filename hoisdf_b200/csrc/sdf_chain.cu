// The candidate chain of `sdf_infer` (upstream main/model.py:316-346 + common/nets/sdf_net.py:87-122) as ONE persistent
// tcgen05 kernel: for a tile of 128 candidate rows
//
//   A0 (128 x 512)  relu(linear_sdfin.layers.0 of the gathered features)       [TMA from HBM, or gathered in-kernel]
//   s1   linear_sdfin.layers.1   512 -> 256, ReLU        -> SKIP[:, 0:256]      (shared memory)
//        NeRF embedding (30) + xyz (3) of the lattice point -> SKIP[:, 256:289]  (shared memory)
//   l0   linh0   289 -> 512, ReLU                        -> ACT (128 x 512 fp16, packed pairs, TENSOR MEMORY)
//   l1   linh1   512 -> 223, ReLU   (A operand read from TMEM)      -> H1 (shared memory)
//   l2   linh2   [input 289 | h1 223] -> 512, ReLU       -> ACT (tensor memory)
//   l3   linh3   512 -> 512, ReLU, fused with linh4 (512 -> 1) + tanh           -> out_sdf[row]  (4 bytes per row)
//
// Nothing but the 4-byte result ever goes back to HBM: the activation tile lives in shared memory (the 289-wide decoder
// input, which the skip connection needs twice, and the 223-wide linh1 output) and in tensor memory (the two 512-wide
// hidden activations, consumed as the A operand of the next MMA straight from TMEM, like P in the attention kernel).
// ONE tensor-core product per K step on fp16-rounded operands (fp32 accumulation in TMEM): this is the candidate
// SCREENING arithmetic of the verified coarse-to-fine cascade -- ~5e-5 absolute on the SDF;
// the survivors are re-ranked by the FP16x3 chain.
//
// Persistent, warp-specialised, one CTA per SM, clusters of CL CTAs that walk consecutive row tiles in lock step and
// share every weight stage (each CTA fetches 1/CL of it and multicasts it):
//   warp 0       weight producer: [128 rows of W x 64 of K] fp16 stages (16 KB, 128-byte swizzle) through a 5-deep ring;
//                1.9 MB of weights per tile stream from L2 (they are shared by every SM, hence always L2 hits)
//   warp 1       tcgen05.mma issuer (one thread), M128 x N128 x K16, N processed in 128-column "quarters" that alternate
//                between two TMEM accumulators D0 / D1, so the epilogue of one quarter overlaps the MMAs of the next
//   warps 2..9   epilogue: tcgen05.ld of a finished quarter (2 warps per TMEM lane quarter, 64 columns each), bias, ReLU,
//                fp16 -> shared memory (swizzled K-major operand layout) / tensor memory (tcgen05.st) / the linh4 dot
//   warp 10      row producer: TMA of the A0 K blocks (a 4-slot ring that re-uses the H1 region, dead during s1), or of
//                the decoder input rows in decoder-only mode
// TMEM: D0 [0,128) D1 [128,256) ACT [256,512) -- all 512 columns.
// SMEM: SKIP 5 x 16 KB | H1 / A0 ring 4 x 16 KB | W ring 5 x 16 KB | barriers = 226 KB.
#include <cstdlib>

#include "tc_common.cuh"

namespace hoisdf {
using namespace tc;

constexpr int CH_BM = 128;
constexpr int CH_BLK = CH_BM * 64 * 2;              // one 128 x 64 fp16 K block: 16 KB
constexpr int CH_SKIP_BLKS = 5;                     // decoder input: 256 fea + 33 (+ padding) = 5 K blocks
constexpr int CH_H1_BLKS = 4;                       // relu(linh1): 223 (+ padding); the A0 ring during s1
constexpr int CH_W_STAGES = 5;
constexpr int CH_OFF_H1 = CH_SKIP_BLKS * CH_BLK;
constexpr int CH_OFF_W = CH_OFF_H1 + CH_H1_BLKS * CH_BLK;
constexpr int CH_OFF_BAR = CH_OFF_W + CH_W_STAGES * CH_BLK;
constexpr int CH_BAR_BYTES = 1024;                  // 44 mbarriers + TMEM slot + the 128-float linh4 reduction scratch
constexpr int CH_SMEM_BYTES = CH_OFF_BAR + CH_BAR_BYTES + 1024 /*align*/;
static_assert(CH_SMEM_BYTES <= 232448, "shared memory budget of one CTA");
constexpr int CH_EPI_WARPS = 8;
constexpr int CH_THREADS = 32 * (2 + CH_EPI_WARPS + 1);   // 352: weights, MMA, 8 epilogue, 1 row producer (TMA)
constexpr int CH_GATHER_WARPS = 8;
constexpr int CH_THREADS_G = 32 * (2 + CH_EPI_WARPS + CH_GATHER_WARPS);   // 576: ... 8 gather warps instead (96 registers each)
constexpr uint32_t CH_TMEM_COLS = 512;
constexpr uint32_t CH_TM_ACT = 256;                 // first TMEM column of the packed fp16 activation tile

enum { CH_MODE_ROWS = 0 /* A0 rows by TMA */, CH_MODE_DECODER = 1 /* decoder input rows by TMA, no s1 */,
       CH_MODE_GATHER = 2 /* A0 rows gathered in the kernel from the fp16 projected maps */ };

// gather mode: the projected pyramid G_l = F_l . W0_l^T (linear_sdfin layer 0 applied per level) as NHWC fp16
struct ChainGather {
  const uint4* __restrict__ map[5];      // (B, H_l, W_l, 512) halfs, 8 per uint4
  int h[5], w[5];
  int levels;
  float nx, ny;                          // (img_w - 1) / 2, (img_h - 1) / 2
  const float* __restrict__ uv;          // (rows, 2) projected pixel of every candidate
  const int64_t* __restrict__ row_offsets;   // (batch + 1) or NULL
  const float* __restrict__ b0;          // linear_sdfin.layers.0.bias (512)
  int64_t batch, rows_per_sample;
};

struct ChainParams {
  const float* __restrict__ b_s1;
  const float* __restrict__ b[4];
  const float* __restrict__ w4;
  const float* __restrict__ b4;
  const int32_t* __restrict__ lattice_index;
  const float* __restrict__ points;
  float* __restrict__ out;
  int64_t rows;
  int tiles;          // ceil(rows / 128)
  int bins;
  int mode;
  float clamp;
  ChainGather g;
};

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// byte offset of 16-byte chunk `c` (8 halfs) of row `r` inside a 128 x 64 fp16 K block with 128-byte swizzle
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// developer profile of CTA 0 (HOISDF_CHAIN_PROF=1): cycles the MMA issuer spent waiting for [0] weights, [1] a free
// accumulator, [2] its A operand, [3] its total; [4] epilogue warp 2: waiting for accumulators, [5] its total;
// [6] weight producer: waiting for a free stage, [7] its total
__device__ long long g_chain_prof[16];   // [8..12] MMA issuer: cycles from 'stage ready' to 'commit issued', per layer

template <int CL, bool PROF, bool GATHER>
__global__ void __launch_bounds__(GATHER ? CH_THREADS_G : CH_THREADS, 1)
sdf_chain_kernel(const __grid_constant__ CUtensorMap map_rows, const __grid_constant__ CUtensorMap map_ws1,
                 const __grid_constant__ CUtensorMap map_w0, const __grid_constant__ CUtensorMap map_w1,
                 const __grid_constant__ CUtensorMap map_w2a, const __grid_constant__ CUtensorMap map_w2b,
                 const __grid_constant__ CUtensorMap map_w3, const ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bars = base + CH_OFF_BAR;
  // barrier map (8 bytes each)
  auto bar_wfull = [&](int s) { return bars + 8u * s; };                       // 0..4
  auto bar_wempty = [&](int s) { return bars + 8u * (5 + s); };                // 5..9
  auto bar_afull = [&](int kb) { return bars + 8u * (10 + kb); };              // 10..17  A0 K block landed
  auto bar_dfull = [&](int b) { return bars + 8u * (18 + b); };                // 18..19
  auto bar_dempty = [&](int b) { return bars + 8u * (20 + b); };               // 20..21
  auto bar_skip = [&](int kb) { return bars + 8u * (22 + kb); };               // 22..26  SKIP K block written
  auto bar_act = [&](int kb) { return bars + 8u * (27 + kb); };                // 27..34  ACT K block (64 columns) written
  auto bar_h1 = [&](int kb) { return bars + 8u * (35 + kb); };                 // 35..38  H1 K block written
  const uint32_t bar_rfree = bars + 8u * 39;                                   // SKIP / H1 region reusable (linh2 done)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + CH_OFF_BAR + 8 * 44);
  float* red = reinterpret_cast<float*>(gen + CH_OFF_BAR + 512);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
  const int cluster = static_cast<int>(blockIdx.x) / CL;
  const int nclusters = static_cast<int>(gridDim.x) / CL;
  const int npairs = (p.tiles + CL - 1) / CL;
  constexpr uint16_t kAllCtas = static_cast<uint16_t>((1u << CL) - 1u);
  const bool decoder_only = p.mode == CH_MODE_DECODER;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_rows) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ws1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w2a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w2b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w3) : "memory");
    for (int s = 0; s < CH_W_STAGES; ++s) {
      mbar_init(bar_wfull(s), 1);
      mbar_init(bar_wempty(s), CL);          // every CTA's tensor core must have consumed the stage
    }
    for (int kb = 0; kb < 8; ++kb) mbar_init(bar_afull(kb), GATHER ? CH_GATHER_WARPS : 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_dfull(b), 1);
      mbar_init(bar_dempty(b), CH_EPI_WARPS);
    }
    // SKIP blocks: written by 4 epilogue warps each (block 4, the embedding: by all 8) -- or by one TMA in decoder mode
    for (int kb = 0; kb < 5; ++kb) mbar_init(bar_skip(kb), decoder_only ? 1 : (kb < 4 ? 4 : CH_EPI_WARPS));
    for (int kb = 0; kb < 8; ++kb) mbar_init(bar_act(kb), 4);
    for (int kb = 0; kb < 4; ++kb) mbar_init(bar_h1(kb), 4);
    mbar_init(bar_rfree, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(CH_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();            // peers' barriers are initialised before any multicast lands there
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // The per-tile program is a list of "quarters" (layer, q): layer 0 = s1 (rows mode only), 1..4 = linh0..linh3.  Every
  // role walks the same list with compact, NON-unrolled loops: the MMA issuer is one thread executing ~500 instructions
  // per quarter, and an unrolled per-layer code path (17 k SASS instructions in the first version) made it stall on
  // instruction fetch for half of its time (ncu: stall_no_inst at every reconvergence point, tensor pipe 25 % active).
  const int layer0 = decoder_only ? 1 : 0;
  auto quarters_of = [](int layer) { return (layer == 0 || layer == 2) ? 2 : 4; };
  auto kblocks_of = [](int layer) { return layer == 1 ? 5 : (layer == 3 ? 9 : 8); };

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer (whole warp, one lane issues)
    {
      uint32_t ws = 0, wph = 1;           // stage, parity of the "stage free" phase to wait for
      const bool prof = PROF && blockIdx.x == 0;
      long long tw = 0;
      const long long tstart = PROF ? clock64() : 0;
      for (int pair = cluster; pair < npairs; pair += nclusters) {
#pragma unroll 1
        for (int layer = layer0; layer < 5; ++layer) {
          const CUtensorMap* m = layer == 0 ? &map_ws1 : layer == 1 ? &map_w0 : layer == 2 ? &map_w1 : layer == 3 ? &map_w2a : &map_w3;
          const int nq = quarters_of(layer), nkb = kblocks_of(layer);
#pragma unroll 1
          for (int q = 0; q < nq; ++q) {
#pragma unroll 1
            for (int kb = 0; kb < nkb; ++kb) {
              const uint32_t s = ws;
              const long long c0 = prof ? clock64() : 0;
              mbar_wait(bar_wempty(s), wph);
              if (prof) tw += clock64() - c0;
              if (++ws == CH_W_STAGES) { ws = 0; wph ^= 1u; }
              const uint32_t dst = base + CH_OFF_W + s * CH_BLK + rank * (CH_BLK / CL);
              const CUtensorMap* mm = (layer == 3 && kb >= 5) ? &map_w2b : m;
              const int col0 = 64 * ((layer == 3 && kb >= 5) ? kb - 5 : kb);
              const int row0 = 128 * q + static_cast<int>(rank) * (128 / CL);
              if (elect_one()) {
                mbar_expect_tx(bar_wfull(s), CH_BLK);
                if (CL > 1) tma_load_2d_mc(dst, mm, bar_wfull(s), col0, row0, kAllCtas);
                else tma_load_2d(dst, mm, bar_wfull(s), col0, row0);
              }
              __syncwarp();
            }
          }
        }
      }
      if (prof && lane == 0) {
        g_chain_prof[6] += tw;
        g_chain_prof[7] += clock64() - tstart;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: the whole warp walks the program
    // (uniform control flow); one elected lane issues the tcgen05 instructions.  tcgen05.mma issue BLOCKS while the tensor
    // core is busy (measured: 64 cycles per M128 x N128 x K16 MMA in the issuing thread), so every instruction of this
    // loop that is not an MMA is tensor-core idle time: stage / phase / descriptor bookkeeping is incremental.
    {
      uint32_t ws = 0, wph = 0, du0 = 0u, du1 = 0u, it = 0;    // weight stage + phase; uses of accumulator D0 / D1 so far
      const uint32_t idesc128 = umma_idesc_f16(CH_BM, 128), idesc96 = umma_idesc_f16(CH_BM, 96);
      const uint32_t act = tmem_base + CH_TM_ACT;
      const uint64_t desc_w0 = umma_desc_sw128(base + CH_OFF_W), desc_a0 = umma_desc_sw128(base);
      constexpr uint64_t kBlkDesc = CH_BLK >> 4;               // one 16 KB block in descriptor address units
      const bool prof = PROF && blockIdx.x == 0;
      long long t_w = 0, t_d = 0, t_a = 0, t_l[5] = {0, 0, 0, 0, 0};
      const long long tstart = PROF ? clock64() : 0;
      for (int pair = cluster; pair < npairs; pair += nclusters, ++it) {
        const uint32_t tp = it & 1u;
#pragma unroll 1
        for (int layer = layer0; layer < 5; ++layer) {
          const int nq = quarters_of(layer), nkb = kblocks_of(layer);
          const bool a_tmem = layer == 2 || layer == 4;
          // "A operand K block kb is in place" barriers of this layer: bar(kb) = abar + 8 kb for kb >= akb0
          //   s1: A0 blocks; linh0: SKIP blocks; linh1 / linh3: ACT blocks (completed twice per tile: linh0 writes phase 0,
          //   linh2 phase 1); linh2: only the H1 blocks (K blocks 5..8) are new, SKIP was already waited for by linh0
          const uint32_t abar = layer == 0 ? bar_afull(0) : layer == 1 ? bar_skip(0) : layer == 3 ? bar_h1(0) - 40u : bar_act(0);
          const int akb0 = layer == 3 ? 5 : 0;
          const uint32_t apar = layer == 2 ? 0u : layer == 4 ? 1u : tp;
#pragma unroll 1
          for (int q = 0; q < nq; ++q) {
            const int b = q & 1;
            long long c0 = prof ? clock64() : 0;
            mbar_wait(bar_dempty(b), ((b ? du1 : du0) & 1u) ^ 1u);     // the epilogue has drained this accumulator
            if (prof) t_d += clock64() - c0;
            tcgen05_fence_after();
            const uint32_t d = tmem_base + static_cast<uint32_t>(b) * 128u;
            const uint32_t idesc = (layer == 2 && q == 1) ? idesc96 : idesc128;
#pragma unroll 1
            for (int kb = 0; kb < nkb; ++kb) {
              if (q == 0 && kb >= akb0) {                                // the A operand of this K block is in place
                if (prof) c0 = clock64();
                mbar_wait(abar + 8u * kb, apar);
                if (prof) t_a += clock64() - c0;
              }
              if (prof) c0 = clock64();
              mbar_wait(bar_wfull(ws), wph);
              if (prof) { const long long c1 = clock64(); t_w += c1 - c0; c0 = c1; }
              tcgen05_fence_after();
              const uint64_t db = desc_w0 + ws * kBlkDesc;
              const int nk = (layer == 3 && kb == 8) ? 2 : (((layer == 1 || layer == 3) && kb == 4) ? 3 : 4);
              // s1 reads the A0 blocks (K blocks 0..3 in the H1 region, 4..7 in SKIP blocks 0..3), linh0 / linh2 SKIP | H1
              const int blk = layer == 0 ? (kb < 4 ? CH_SKIP_BLKS + kb : kb - 4) : kb;
              const uint64_t da = desc_a0 + blk * kBlkDesc;
              const uint32_t at = act + static_cast<uint32_t>(kb * 32);
              if (elect_one()) {
                if (a_tmem) {
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk)
                    if (kk < nk)
                      umma_f16_ts(d, at + static_cast<uint32_t>(kk * 8), db + static_cast<uint64_t>(kk * 2), idesc,
                                  (kb | kk) != 0 ? 1u : 0u);
                } else {
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk)
                    if (kk < nk)
                      umma_f16(d, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc,
                               (kb | kk) != 0 ? 1u : 0u);
                }
                if (CL > 1) umma_commit_mc(bar_wempty(ws), kAllCtas);
                else umma_commit(bar_wempty(ws));
              }
              __syncwarp();
              if (++ws == CH_W_STAGES) { ws = 0; wph ^= 1u; }
              if (prof) t_l[layer] += clock64() - c0;
            }
            // publish the accumulator.  s1: both halves together -- its epilogue overwrites SKIP blocks that still hold
            // A0 K blocks until the second half has been computed
            if (layer == 0) {
              if (q == 1) {
                if (elect_one()) {
                  umma_commit(bar_dfull(0));
                  umma_commit(bar_dfull(1));
                }
                ++du0;
                ++du1;
              }
            } else {
              if (elect_one()) umma_commit(bar_dfull(b));
              if (b) ++du1; else ++du0;
            }
            __syncwarp();
          }
          if (layer == 3 && elect_one()) umma_commit(bar_rfree);   // SKIP / H1 no longer read: the next tile's rows may land
          __syncwarp();
        }
      }
      if (prof && lane == 0) {
        g_chain_prof[0] += t_w;
        g_chain_prof[1] += t_d;
        g_chain_prof[2] += t_a;
        g_chain_prof[3] += clock64() - tstart;
        for (int l = 0; l < 5; ++l) g_chain_prof[8 + l] += t_l[l];
      }
    }
  } else if (!GATHER && warp == 2 + CH_EPI_WARPS) {
    // ------------------------------------------------------------------ row producer (TMA; whole warp, one lane issues)
    {
      uint32_t it = 0;
      for (int pair = cluster; pair < npairs; pair += nclusters, ++it) {
        const int row0 = (pair * CL + static_cast<int>(rank)) * CH_BM;   // beyond the last row: zero-filled tile
        if (it > 0) mbar_wait(bar_rfree, (it - 1) & 1u);
        if (decoder_only) {
          if (elect_one()) {
            for (int kb = 0; kb < 5; ++kb) {
              mbar_expect_tx(bar_skip(kb), CH_BLK);
              tma_load_2d(base + kb * CH_BLK, &map_rows, bar_skip(kb), 64 * kb, row0);
            }
          }
          __syncwarp();
        } else {
          if (elect_one()) {
            for (int kb = 0; kb < 8; ++kb) {
              const int blk = kb < 4 ? CH_SKIP_BLKS + kb : kb - 4;
              mbar_expect_tx(bar_afull(kb), CH_BLK);
              tma_load_2d(base + blk * CH_BLK, &map_rows, bar_afull(kb), 64 * kb, row0);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 2 + CH_EPI_WARPS) {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int ew = warp - 2;
    const int lq = warp & 3;                    // TMEM lane quarter this warp may read
    const int h = ew >> 2;                      // which 64-column half of a 128-column quarter it owns
    const int r = lq * 32 + lane;               // row inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>(lq * 32) << 16;
    uint32_t eu0 = 0u, eu1 = 0u;
    const bool prof = PROF && blockIdx.x == 0 && warp == 2 && lane == 0;
    long long t_e = 0;
    const long long tstart = PROF ? clock64() : 0;

    for (int pair = cluster; pair < npairs; pair += nclusters) {
      const int64_t row = static_cast<int64_t>(pair * CL + static_cast<int>(rank)) * CH_BM + r;
      const bool row_ok = row < p.rows;
      if (!decoder_only) {
        // ---- NeRF embedding + xyz -> SKIP block 4 (columns 256..319).  Every MMA that read the previous tile's SKIP
        // has completed: this warp has already seen that tile's linh3 accumulators.
        uint8_t* blk = gen + 4 * CH_BLK;
        float x[3] = {0.f, 0.f, 0.f};
        if (row_ok) {
          if (p.lattice_index != nullptr) lattice_point(p.lattice_index[row], p.bins, x[0], x[1], x[2]);
          else { x[0] = p.points[row * 3 + 0]; x[1] = p.points[row * 3 + 1]; x[2] = p.points[row * 3 + 2]; }
        }
        if (h == 0) {
          // columns 0..31 of the block: 30 embedding values (octave-major: sin xyz, cos xyz) + x, y
          uint32_t pk[16];
#pragma unroll
          for (int oct = 0; oct < 5; ++oct) {
            float sn[3], cs[3];
#pragma unroll
            for (int w = 0; w < 3; ++w) {
              const float a = x[w] * static_cast<float>(1 << oct);     // exact scaling by 2^k
              sn[w] = sinf(a);
              cs[w] = cosf(a);
            }
            pk[oct * 3 + 0] = cvt_f16x2_sat(sn[0], sn[1]);
            pk[oct * 3 + 1] = cvt_f16x2_sat(sn[2], cs[0]);
            pk[oct * 3 + 2] = cvt_f16x2_sat(cs[1], cs[2]);
          }
          pk[15] = cvt_f16x2_sat(x[0], x[1]);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(blk + sw128_off(r, c)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        } else {
          *reinterpret_cast<uint4*>(blk + sw128_off(r, 4)) = make_uint4(cvt_f16x2_sat(x[2], 0.f), 0u, 0u, 0u);
#pragma unroll
          for (int c = 5; c < 8; ++c) *reinterpret_cast<uint4*>(blk + sw128_off(r, c)) = make_uint4(0u, 0u, 0u, 0u);
        }
        fence_async_smem();                     // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_skip(4));
      }
      float part = 0.f;                         // linh4 partial dot product of this thread's row
#pragma unroll 1
      for (int layer = layer0; layer < 5; ++layer) {
        const int nq = quarters_of(layer);
        const float* bias = layer == 0 ? p.b_s1 : layer == 1 ? p.b[0] : layer == 2 ? p.b[1] : layer == 3 ? p.b[2] : p.b[3];
#pragma unroll 1
        for (int q = 0; q < nq; ++q) {
          const int b = q & 1;
          const long long tc0 = prof ? clock64() : 0;
          mbar_wait(bar_dfull(b), (b ? eu1 : eu0) & 1u);
          if (prof) t_e += clock64() - tc0;
          tcgen05_fence_after();
          const uint32_t t0 = tmem_base + static_cast<uint32_t>(b * 128 + h * 64) + lane_addr;
          // this warp's 64 columns in two passes of 32 (the kernel runs 18 warps in gather mode: 96 registers per thread)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int c0 = 128 * q + 64 * h + 32 * half;     // first output column of this pass
            uint32_t a[32];
            tmem_ld32(t0 + 32 * half, a);
            tmem_ld_wait();
            if (half == 1) {                                   // the accumulator is in registers: hand the buffer back
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_dempty(b));
              if (b) ++eu1; else ++eu0;
            }
            if (layer == 4) {
              // linh3 + linh4: relu(acc + b3) . w4 over this warp's columns
              float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0) + j);
                const float4 ww = __ldg(reinterpret_cast<const float4*>(p.w4 + c0) + j);
                s4[0] = fmaf(fmaxf(__uint_as_float(a[4 * j]) + bb.x, 0.f), ww.x, s4[0]);
                s4[1] = fmaf(fmaxf(__uint_as_float(a[4 * j + 1]) + bb.y, 0.f), ww.y, s4[1]);
                s4[2] = fmaf(fmaxf(__uint_as_float(a[4 * j + 2]) + bb.z, 0.f), ww.z, s4[2]);
                s4[3] = fmaf(fmaxf(__uint_as_float(a[4 * j + 3]) + bb.w, 0.f), ww.w, s4[3]);
              }
              part += (s4[0] + s4[1]) + (s4[2] + s4[3]);
              continue;
            }
            // bias + ReLU + fp16: 32 columns -> 16 packed pairs.  linh1 has 223 outputs: the tail is zero-filled.
            const int nvalid = layer == 2 ? min(32, 223 - c0) : 32;
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
              if (4 * j + 3 < nvalid) {
                bb = __ldg(reinterpret_cast<const float4*>(bias + c0) + j);
              } else {
                if (4 * j < nvalid) bb.x = __ldg(bias + c0 + 4 * j);
                if (4 * j + 1 < nvalid) bb.y = __ldg(bias + c0 + 4 * j + 1);
                if (4 * j + 2 < nvalid) bb.z = __ldg(bias + c0 + 4 * j + 2);
              }
              const float v0 = 4 * j < nvalid ? fmaxf(__uint_as_float(a[4 * j]) + bb.x, 0.f) : 0.f;
              const float v1 = 4 * j + 1 < nvalid ? fmaxf(__uint_as_float(a[4 * j + 1]) + bb.y, 0.f) : 0.f;
              const float v2 = 4 * j + 2 < nvalid ? fmaxf(__uint_as_float(a[4 * j + 2]) + bb.z, 0.f) : 0.f;
              const float v3 = 4 * j + 3 < nvalid ? fmaxf(__uint_as_float(a[4 * j + 3]) + bb.w, 0.f) : 0.f;
              pk[2 * j] = cvt_f16x2_sat(v0, v1);
              pk[2 * j + 1] = cvt_f16x2_sat(v2, v3);
            }
            if (layer == 1 || layer == 3) {
              // -> ACT (TMEM): columns [c0, c0 + 32) = packed columns [c0 / 2, + 16)
              tmem_st16(tmem_base + CH_TM_ACT + static_cast<uint32_t>(c0 >> 1) + lane_addr, pk);
            } else {
              // -> this lane's row of a shared-memory K block (s1 -> SKIP block 2q+h, linh1 -> H1 block 2q+h): 64 bytes
              uint8_t* blk = gen + static_cast<uint32_t>((layer == 0 ? 0 : CH_SKIP_BLKS) + 2 * q + h) * CH_BLK;
#pragma unroll
              for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4*>(blk + sw128_off(r, 4 * half + c)) =
                    make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            }
          }
          if (layer == 4) continue;
          if (layer == 1 || layer == 3) {
            tmem_st_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_act(2 * q + h));
          } else {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(layer == 0 ? bar_skip(2 * q + h) : bar_h1(2 * q + h));
          }
        }
      }
      if (h == 1) red[r] = part;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + lq) : "memory");     // the two warps of this lane quarter
      if (h == 0 && row_ok) {
        float t = tanhf(part + red[r] + __ldg(p.b4));
        if (p.clamp > 0.f) t = fminf(fmaxf(t, -p.clamp), p.clamp);
        p.out[row] = t;
      }
    }
    if (prof) {
      g_chain_prof[4] += t_e;
      g_chain_prof[5] += clock64() - tstart;
    }
  } else if (GATHER) {
    // ------------------------------------------------------------------ gather warps (8): A0 = relu(b0 + sum over the 5
    // levels of the bilinear sample of the projected fp16 map) written straight into the swizzled A-operand blocks.
    // Warp gw owns rows [16 gw, 16 gw + 16) of the tile; per K block (64 channels) it makes 4 passes of 4 rows: lanes
    // 8 i .. 8 i + 7 hold the 8 channel octets of one row, so every tap is one 128-byte line read by 8 lanes.
    // The tap parameters of a row (pixel index, edge flags, fractions for 5 levels) are computed once per tile by lane
    // (row & 15) and broadcast with shuffles.
    const int gw = warp - (2 + CH_EPI_WARPS);
    const int oct = lane & 7, rsub = lane >> 3;
    uint32_t it = 0;
    for (int pair = cluster; pair < npairs; pair += nclusters, ++it) {
      const int64_t row = static_cast<int64_t>(pair * CL + static_cast<int>(rank)) * CH_BM + gw * 16 + (lane & 15);
      float u = 0.f, v = 0.f;
      int64_t sample = 0;
      if (row < p.rows) {
        u = __ldg(p.g.uv + 2 * row);
        v = __ldg(p.g.uv + 2 * row + 1);
        if (p.g.row_offsets == nullptr) {
          sample = row / p.g.rows_per_sample;
        } else {
          int64_t lo = 0, hi = p.g.batch;                      // largest b with offsets[b] <= row
          while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(p.g.row_offsets + mid) <= row) lo = mid; else hi = mid;
          }
          sample = lo;
        }
      }
      // ATen grid_sampler_2d arithmetic (align_corners = True, border padding), as csrc/gather.cu make_taps
      const float gx = __fdiv_rn(__fsub_rn(u, p.g.nx), p.g.nx), gy = __fdiv_rn(__fsub_rn(v, p.g.ny), p.g.ny);
      uint32_t tp_o[5];
      float tp_x[5], tp_y[5];
#pragma unroll
      for (int l = 0; l < 5; ++l) {
        const int W = p.g.w[l], H = p.g.h[l];
        float x = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), static_cast<float>(W - 1));
        float y = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), static_cast<float>(H - 1));
        x = fminf(static_cast<float>(W - 1), fmaxf(x, 0.f));
        y = fminf(static_cast<float>(H - 1), fmaxf(y, 0.f));
        const float x0f = floorf(x), y0f = floorf(y);
        const int x0 = static_cast<int>(x0f), y0 = static_cast<int>(y0f);
        tp_x[l] = x - x0f;
        tp_y[l] = y - y0f;
        const uint32_t pix = static_cast<uint32_t>((static_cast<int>(sample) * H + y0) * W + x0);
        tp_o[l] = pix | (x0 + 1 <= W - 1 ? 1u << 30 : 0u) | (y0 + 1 <= H - 1 ? 1u << 31 : 0u);
      }
      if (it > 0) mbar_wait(bar_rfree, (it - 1) & 1u);          // linh2 of the previous tile has read SKIP / H1
#pragma unroll 1
      for (int kb = 0; kb < 8; ++kb) {
        uint8_t* blk = gen + static_cast<uint32_t>(kb < 4 ? CH_SKIP_BLKS + kb : kb - 4) * CH_BLK;
        const float4 ba = __ldg(reinterpret_cast<const float4*>(p.g.b0 + kb * 64 + oct * 8));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.g.b0 + kb * 64 + oct * 8 + 4));
        const uint32_t coff = static_cast<uint32_t>(kb * 8 + oct);     // uint4 index of this lane's 8 channels inside a pixel
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          const int src = i * 4 + rsub;                                 // row inside this warp's 16 = the lane that holds its taps
          float acc[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int l = 0; l < 5; ++l) {
            if (l < p.g.levels) {
              const uint32_t o = __shfl_sync(0xffffffffu, tp_o[l], src);
              const float tx = __shfl_sync(0xffffffffu, tp_x[l], src), ty = __shfl_sync(0xffffffffu, tp_y[l], src);
              const uint32_t pix = o & 0x3fffffffu, dx = (o >> 30) & 1u, dy = o >> 31;
              const uint4* m = p.g.map[l] + coff;
              const uint32_t p00 = pix * 64u, p01 = (pix + dx) * 64u;
              const uint32_t down = dy * static_cast<uint32_t>(p.g.w[l]) * 64u;
              const uint4 q00 = __ldg(m + p00), q01 = __ldg(m + p01), q10 = __ldg(m + p00 + down), q11 = __ldg(m + p01 + down);
              const float ax = 1.f - tx, ay = 1.f - ty;               // == ATen's (x0 + 1) - x, see DESIGN.md
              const float w00 = ax * ay, w01 = dx ? tx * ay : 0.f, w10 = dy ? ax * ty : 0.f, w11 = (dx & dy) ? tx * ty : 0.f;
              const uint32_t a00[4] = {q00.x, q00.y, q00.z, q00.w}, a01[4] = {q01.x, q01.y, q01.z, q01.w};
              const uint32_t a10[4] = {q10.x, q10.y, q10.z, q10.w}, a11[4] = {q11.x, q11.y, q11.z, q11.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f00 = __half22float2(*reinterpret_cast<const __half2*>(&a00[j]));
                const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&a01[j]));
                const float2 f10 = __half22float2(*reinterpret_cast<const __half2*>(&a10[j]));
                const float2 f11 = __half22float2(*reinterpret_cast<const __half2*>(&a11[j]));
                acc[2 * j] += fmaf(f11.x, w11, fmaf(f10.x, w10, fmaf(f01.x, w01, f00.x * w00)));
                acc[2 * j + 1] += fmaf(f11.y, w11, fmaf(f10.y, w10, fmaf(f01.y, w01, f00.y * w00)));
              }
            }
          }
          uint4 o4;
          o4.x = cvt_f16x2_sat(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f));
          o4.y = cvt_f16x2_sat(fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f));
          o4.z = cvt_f16x2_sat(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f));
          o4.w = cvt_f16x2_sat(fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f));
          *reinterpret_cast<uint4*>(blk + sw128_off(gw * 16 + src, oct)) = o4;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_afull(kb));
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();              // no CTA exits while a peer may still multicast to it / signal its barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(CH_TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static bool chain_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  return make_tiled_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

template <int CL, bool GATHER>
static int chain_max_clusters() {
  static int cached = -1;
  if (cached >= 0) return cached;
  auto kern = sdf_chain_kernel<CL, false, GATHER>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BYTES);
  int n = 0;
  if (CL == 1) {
    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    n = cudaGetDeviceProperties(&prop, dev) == cudaSuccess ? prop.multiProcessorCount : kNumSMs;
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(kNumSMs / CL * CL);
    cfg.blockDim = dim3(GATHER ? CH_THREADS_G : CH_THREADS);
    cfg.dynamicSmemBytes = CH_SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) n = kNumSMs / CL;
    (void)cudaGetLastError();
  }
  cached = n;
  return n;
}

template <int CL, bool PROF, bool GATHER>
static int chain_launch(const CUtensorMap* maps, const ChainParams& p, cudaStream_t s) {
  auto kern = sdf_chain_kernel<CL, PROF, GATHER>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BYTES);
  if (e != cudaSuccess) return static_cast<int>(e);
  const int64_t npairs = ceil_div(p.tiles, CL);
  const int64_t cap = chain_max_clusters<CL, GATHER>();
  const int64_t clusters = npairs < cap ? npairs : cap;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(clusters * CL));
  cfg.blockDim = dim3(GATHER ? CH_THREADS_G : CH_THREADS);
  cfg.dynamicSmemBytes = CH_SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], p);
  if (e != cudaSuccess) return static_cast<int>(e);
  return launch_status();
}

template <int CL>
static int chain_dispatch(const CUtensorMap* maps, const ChainParams& p, bool prof, cudaStream_t s) {
  if (p.mode == CH_MODE_GATHER)
    return prof ? chain_launch<CL, true, true>(maps, p, s) : chain_launch<CL, false, true>(maps, p, s);
  return prof ? chain_launch<CL, true, false>(maps, p, s) : chain_launch<CL, false, false>(maps, p, s);
}

// fp32 -> fp16 (round to nearest, saturating), 8 values per thread: the fp16 copy of the projected maps that gather mode reads
__global__ void f32_to_f16_kernel(const float4* __restrict__ x, uint4* __restrict__ y, int64_t n8) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const float4 a = __ldg(x + 2 * i), b = __ldg(x + 2 * i + 1);
  y[i] = make_uint4(cvt_f16x2_sat(a.x, a.y), cvt_f16x2_sat(a.z, a.w), cvt_f16x2_sat(b.x, b.y), cvt_f16x2_sat(b.z, b.w));
}

}  // namespace hoisdf

using namespace hoisdf;

// developer hook: read and reset the CTA-0 wait-cycle profile (see g_chain_prof)
extern "C" __attribute__((visibility("default"))) int hoisdf_debug_chain_profile(long long* out16) {
  long long zero[16] = {0};
  if (cudaMemcpyFromSymbol(out16, g_chain_prof, sizeof(zero)) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(g_chain_prof, zero, sizeof(zero)) == cudaSuccess ? 0 : -1;
}

HOISDF_API int hoisdf_f32_to_f16(const float* x, uint16_t* y, int64_t n, void* stream) {
  if (x == nullptr || y == nullptr) return HOISDF_E_NULL;
  if (n == 0) return HOISDF_OK;
  if (n < 0 || (n & 7)) return HOISDF_E_SHAPE;
  if (!aligned16(x) || !aligned16(y)) return HOISDF_E_ALIGN;
  f32_to_f16_kernel<<<static_cast<unsigned>(ceil_div(n >> 3, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<uint4*>(y), n >> 3);
  return launch_status();
}

HOISDF_API int hoisdf_sdf_chain_fwd(const hoisdf_sdf_chain_args* a, void* stream) {
  if (a == nullptr || a->out_sdf == nullptr || a->w4 == nullptr || a->b4 == nullptr) return HOISDF_E_NULL;
  const bool decoder_only = a->x != nullptr;
  const bool gather = a->gmaps != nullptr;
  if ((decoder_only ? 1 : 0) + (gather ? 1 : 0) + (a->a0 != nullptr ? 1 : 0) != 1) return HOISDF_E_NULL;   // exactly one source
  if (!decoder_only && (a->w_s1 == nullptr || a->b_s1 == nullptr || (a->lattice_index == nullptr && a->points == nullptr)))
    return HOISDF_E_NULL;
  if (gather && (a->uv == nullptr || a->b_s0 == nullptr)) return HOISDF_E_NULL;
  for (int l = 0; l < 4; ++l)
    if (a->w[l] == nullptr || a->b[l] == nullptr) return HOISDF_E_NULL;
  if (a->rows == 0) return HOISDF_OK;
  if (a->rows < 0 || a->rows >= (int64_t(1) << 30)) return HOISDF_E_SHAPE;
  // weight pitches: linh0 (512, >= 296) zero beyond column 289; linh1 (223, >= 512); linh2 (512, >= 520) in the
  // [input 289 | 0 x7 | h1 223 | 0] column layout; linh3 (512, >= 512)
  if (a->ldw[0] < 296 || a->ldw[1] < 512 || a->ldw[2] < 520 || a->ldw[3] < 512) return HOISDF_E_SHAPE;
  if (!decoder_only && a->ldw_s1 < 512) return HOISDF_E_SHAPE;
  if (a->a0 != nullptr && a->lda0 < 512) return HOISDF_E_SHAPE;
  if (decoder_only && a->ldx < 296) return HOISDF_E_SHAPE;
  for (int l = 0; l < 4; ++l)
    if ((a->ldw[l] & 7) || !aligned16(a->w[l])) return HOISDF_E_ALIGN;
  if (decoder_only && ((a->ldx & 7) || !aligned16(a->x))) return HOISDF_E_ALIGN;
  if (a->a0 != nullptr && ((a->lda0 & 7) || !aligned16(a->a0))) return HOISDF_E_ALIGN;
  if (!decoder_only && ((a->ldw_s1 & 7) || !aligned16(a->w_s1))) return HOISDF_E_ALIGN;
  ChainParams p{};
  if (gather) {
    const hoisdf_pyramid_h* g = a->gmaps;
    if (g->levels < 1 || g->levels > 5 || g->c != 512 || a->batch <= 0) return HOISDF_E_SHAPE;
    if (a->row_offsets == nullptr && a->rows_per_sample <= 0) return HOISDF_E_SHAPE;
    if (!aligned16(a->b_s0)) return HOISDF_E_ALIGN;
    for (int l = 0; l < 5; ++l) {
      const int ll = l < g->levels ? l : 0;                    // unused levels alias level 0 (never read)
      if (g->map[ll] == nullptr) return HOISDF_E_NULL;
      if (!aligned16(g->map[ll])) return HOISDF_E_ALIGN;
      if (g->h[ll] <= 0 || g->w[ll] <= 0 || a->batch * g->h[ll] * g->w[ll] >= (int64_t(1) << 30)) return HOISDF_E_SHAPE;
      p.g.map[l] = reinterpret_cast<const uint4*>(g->map[ll]);
      p.g.h[l] = g->h[ll];
      p.g.w[l] = g->w[ll];
    }
    p.g.levels = g->levels;
    p.g.nx = static_cast<float>(g->img_w - 1) / 2.0f;
    p.g.ny = static_cast<float>(g->img_h - 1) / 2.0f;
    p.g.uv = a->uv; p.g.row_offsets = a->row_offsets; p.g.b0 = a->b_s0;
    p.g.batch = a->batch; p.g.rows_per_sample = a->rows_per_sample;
  }
  const int64_t tiles = ceil_div(a->rows, CH_BM);
  static const int force_cl = [] { const char* e = getenv("HOISDF_CHAIN_CL"); return e ? atoi(e) : 0; }();
  const int cl = force_cl == 1 ? 1 : (tiles >= 2 ? 2 : 1);
  CUtensorMap maps[7];
  const int wbox = 128 / cl;
  bool ok = decoder_only ? chain_map(&maps[0], a->x, a->rows, 296, a->ldx, CH_BM)
            : gather     ? chain_map(&maps[0], a->w[0], 512, 296, a->ldw[0], wbox)      // unused in this mode
                         : chain_map(&maps[0], a->a0, a->rows, 512, a->lda0, CH_BM);
  ok = ok && (decoder_only ? chain_map(&maps[1], a->w[0], 512, 296, a->ldw[0], wbox)      // unused in this mode
                           : chain_map(&maps[1], a->w_s1, 256, 512, a->ldw_s1, wbox));
  ok = ok && chain_map(&maps[2], a->w[0], 512, 296, a->ldw[0], wbox);
  ok = ok && chain_map(&maps[3], a->w[1], 223, 512, a->ldw[1], wbox);
  ok = ok && chain_map(&maps[4], a->w[2], 512, 296, a->ldw[2], wbox);
  ok = ok && chain_map(&maps[5], a->w[2] + 296, 512, 224, a->ldw[2], wbox);
  ok = ok && chain_map(&maps[6], a->w[3], 512, 512, a->ldw[3], wbox);
  if (!ok) return HOISDF_E_UNSUPPORTED;
  p.b_s1 = a->b_s1;
  for (int l = 0; l < 4; ++l) p.b[l] = a->b[l];
  p.w4 = a->w4; p.b4 = a->b4;
  p.lattice_index = a->lattice_index; p.points = a->points; p.bins = a->bins;
  p.out = a->out_sdf; p.rows = a->rows; p.tiles = static_cast<int>(tiles);
  p.mode = decoder_only ? CH_MODE_DECODER : (gather ? CH_MODE_GATHER : CH_MODE_ROWS);
  p.clamp = a->clamp;
  static const int prof = [] { const char* e = getenv("HOISDF_CHAIN_PROF"); return e ? atoi(e) : 0; }();
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return cl == 2 ? chain_dispatch<2>(maps, p, prof != 0, s) : chain_dispatch<1>(maps, p, prof != 0, s);
}
