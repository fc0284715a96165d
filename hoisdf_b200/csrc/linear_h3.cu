// nn.Linear on the 5th-generation tensor cores at FP16 rate with fp32-grade accuracy ("FP16x3").
//
//   x = x_hi + x_lo,  w = w_hi + w_lo   (hi = fp16-rounded value, lo = exact residual; 11 + 11 significand bits,
//                                        the same 22 bits a 3xTF32 split carries -- see tc_common.cuh)
//   X.W^T ~= x_hi.w_hi + x_lo.w_hi + x_hi.w_lo      (dropped x_lo.w_lo < 2^-22 relative)
// kind::f16 MMAs run at twice the kind::tf32 rate and their operands are half as wide, so the same fp32-grade product
// costs half the tensor time and half the L2 -> SM bytes of the 3xTF32 kernel (linear_tc.cu).  To keep every term
// in ONE fp32 accumulator (so that TMEM can hold two of them and the epilogue of tile i overlaps the main loop of
// tile i+1) all three products are formed at a common scale of 2^11:
//   activations (split-half format, produced by the previous kernel's epilogue):  x_hi, x_lo' = x_lo * 2^11
//   weights (3 planes, packed once):  A = w_hi * 2^11,  B = w_hi,  C = w_lo * 2^11          (|w| < 32)
//   acc = x_hi.A + x_lo'.B + x_hi.C = 2^11 * (x_hi.w_hi + x_lo.w_hi + x_hi.w_lo);   y = acc * 2^-11 + bias
//
// Persistent, warp-specialised, one CTA per SM (320 threads), clusters of CL CTAs that own CL consecutive M tiles of
// the same N tile:
//   warp 0      TMA producer: per 32-half K block the CTA's own x_hi / x_lo' tiles (128 rows) and ITS 1/CL slice of
//               the three W planes, multicast to every CTA of the cluster -> 3-stage ring of 64 KB stages
//   warp 1      tcgen05.mma issuer (one thread): 6 MMAs (M128 x N<=256 x K16) per stage; K is accumulated in TMEM one
//               CHUNK (4 stages = 128 of K) at a time, alternating between two 256-column TMEM buffers
//   warps 2..9  drain + epilogue: tcgen05.ld of every finished chunk, added round-to-nearest into 128 fp32 registers
//               per thread (the tensor core truncates on accumulate -- see the note at the kernel); after the last
//               chunk: scale + bias (+ split-half residual) + ReLU, then fp32 or split-half output staged in a
//               swizzled smem box and written by TMA, or direct stores for the fp32-residual / ragged-group modes.
#include <cstdlib>

#include "tc_common.cuh"

namespace hoisdf {
using namespace tc;

constexpr int H3_BM = 128;
constexpr int H3_BN = 256;
constexpr int H3_BK = 32;                               // halfs per stage row = one 64-byte swizzle row
constexpr int H3_STAGES = 3;                            // three-product mode: 3 stages of 64 KB
constexpr int H3_STAGES_1P = 8;                         // single-product mode: 8 stages of 24 KB (x_hi + w_hi only)
constexpr int H3_MAX_STAGES = 8;
constexpr int H3_X_BYTES = H3_BM * H3_BK * 2;           // 8 KB per plane
constexpr int H3_W_BYTES = H3_BN * H3_BK * 2;           // 16 KB per plane
constexpr int H3_STAGE_BYTES = 2 * H3_X_BYTES + 3 * H3_W_BYTES;   // 64 KB
constexpr int H3_STAGE_BYTES_1P = H3_X_BYTES + H3_W_BYTES;         // 24 KB
static_assert(H3_STAGES_1P * H3_STAGE_BYTES_1P <= H3_STAGES * H3_STAGE_BYTES, "single-product ring must fit");
constexpr int H3_EPI_WARPS = 8;                         // drain + epilogue warps (2 per TMEM lane quarter)
constexpr int H3_EPI_SLOT = 4096;                       // one 32 x 32 fp32 box (or hi + lo half boxes) per warp
constexpr int H3_EPI_BYTES = H3_EPI_WARPS * H3_EPI_SLOT;
constexpr int H3_BAR_BYTES = 256;
constexpr int H3_SMEM_BYTES = H3_STAGES * H3_STAGE_BYTES + H3_EPI_BYTES + H3_BAR_BYTES + 1024 /*align*/;
constexpr int H3_THREADS = 64 + 32 * H3_EPI_WARPS;      // 320
constexpr uint32_t H3_TMEM_COLS = 512;                  // two 256-column fp32 chunk accumulators
constexpr int H3_CHUNK_KB = 4;                          // K blocks per accumulation chunk (4 x 32 = 128 of K)

enum { H3_OUT_F32_TMA = 0, H3_OUT_SPLIT_TMA = 1, H3_OUT_F32_DIRECT = 2 };

struct H3Params {
  const float* __restrict__ bias;
  const float* __restrict__ residual;
  const __half* __restrict__ r_hi;   // residual in split-half format (TMA-store output modes), row pitch ldr halfs;
  const __half* __restrict__ r_lo;   // row index = dense output row (Linear) / raster output pixel (convolution)
  int64_t ldr;
  float* __restrict__ y;
  int64_t ldy;
  int64_t rows_per_batch;   // X rows form groups of this many rows (plain GEMM: = M, one group)
  int tiles_per_batch;      // ceil(rows_per_batch / 128)
  int m_tiles;              // groups * tiles_per_batch
  int m_blocks;             // ceil(m_tiles / CL): work items along M
  int n_tiles;
  int n, k, act;
  int out_mode;
  int chunk_kb;             // K blocks accumulated in TMEM before the partial sum is drained into registers
  int single;               // 1: ONE product x_hi . w_hi per K step (11-bit operands, ~5e-4 relative): candidate
                            // pre-screening only -- 1/3 of the tensor work, 3/8 of the operand bytes, deeper ring
  int w_rows;               // three-product mode: rows of W kept per plane in a stage (64 / 128 / 256 >= N of one N tile).
  int bn;                   // columns of one N tile (256; 64 / 128 when the whole layer is narrower, or -- for launches that
                            // would occupy only a few SMs -- to split N into more, narrower tiles with deeper rings)
  int nstages;              // A narrow layer does not pay for 256 rows of (mostly zero-filled) W per K block: its stages
                            // shrink from 64 KB to 40 / 28 KB and the ring deepens from 3 to 4 / 6 stages -- the thin
                            // shapes are bound by load latency, not by tensor work (N <= 64 MMAs sit on the issue floor)
  // implicit-GEMM convolution (taps > 0): X is an NHWC image batch (4-D tensor maps), M = output pixels in (b, y, x)
  // order, K = taps x Cin; a 128-pixel M tile is a (tb x ty x tx) block of the output grid, a warp's 32 rows a
  // (wy x wx) block
  int taps, cin_blocks, out_w, out_h, stride, wx, wy;
  int8_t dy[16], dx[16];
  float w_scale;            // power-of-two scale the packed weights were divided by (|w| >= 32 layers); 1 normally
  const float* __restrict__ y_scale;   // optional device scalar multiplied into the product (training backward); may be null
  // split-K (plain GEMM, fp32 TMA output, no activation): a work item is (M block, N tile, K range); every item ADDS its
  // partial product into the zero-initialised output with a TMA reduction.  For the weight gradients of the training step
  // (a 256 x 256 output over a 51 200-long contraction is ONE tile pair otherwise: 2 of 148 SMs busy)
  int splits;               // <= 1: off
  int kb_per_split;         // K blocks per item
};

// Accuracy note (why the accumulator is drained in chunks).  The tensor core TRUNCATES when it adds a K=16 partial
// product into the fp32 TMEM accumulator, so a long contraction accumulated entirely in TMEM picks up a bias that
// grows linearly with K (measured 1.7e-5 relative at K = 4608 -- 10x the fp32 FMA pipe, and a chain of convolutions
// amplifies it).  Here TMEM only ever holds the sum over one chunk of `chunk_kb` K blocks; the eight epilogue warps
// pull every finished chunk out with tcgen05.ld and add it, round-to-nearest, into fp32 registers (128 per thread)
// while the tensor core fills the other TMEM buffer with the next chunk.  24 truncating adds per chunk instead of
// 3K/16 per tile: the result is as accurate as an fp32 FMA GEMM, at the same tensor-core rate.
// Template parameters: CL = cluster size (CTAs sharing an N tile); OUT = H3_OUT_*; SINGLE = one tensor-core product per
// K step (pre-screening); RES = split-half residual added in the epilogue.  Compile-time modes keep each
// instantiation's epilogue small (instruction cache) and branch-free.
template <int CL, int OUT, bool SINGLE, bool RES, bool PAIR>
__global__ void __launch_bounds__(H3_THREADS, 1)
linear_h3_kernel(const __grid_constant__ CUtensorMap map_xhi, const __grid_constant__ CUtensorMap map_xlo,
                 const __grid_constant__ CUtensorMap map_wa, const __grid_constant__ CUtensorMap map_wb,
                 const __grid_constant__ CUtensorMap map_wc, const __grid_constant__ CUtensorMap map_y0,
                 const __grid_constant__ CUtensorMap map_y1, const H3Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t epi = base + H3_STAGES * H3_STAGE_BYTES;
  const uint32_t bars = epi + H3_EPI_BYTES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + H3_STAGES * H3_STAGE_BYTES + H3_EPI_BYTES + 192);
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 8u * (H3_MAX_STAGES + s); };
  auto bar_cfull = [&](uint32_t b) { return bars + 8u * (2 * H3_MAX_STAGES + b); };
  auto bar_cempty = [&](uint32_t b) { return bars + 8u * (2 * H3_MAX_STAGES + 2 + b); };
  constexpr bool single = SINGLE;
  constexpr bool DIRECT = OUT == H3_OUT_F32_DIRECT;
  constexpr bool RES_STAGED = RES && OUT == H3_OUT_SPLIT_TMA;     // residual tile staged through the epilogue's smem slot
  static_assert(!PAIR || (CL == 2 && !SINGLE), "cta_group::2 mode: a pair of CTAs, three-product arithmetic");
  // PAIR (tcgen05.mma.cta_group::2): the two CTAs of a cluster compute their two M tiles with ONE M = 256 instruction stream
  // issued by the leader (rank 0).  Each CTA stages its own X tiles and only HALF of the W rows (the tensor cores of both
  // SMs read both halves), so a K block costs 40 KB of TMA writes + 48 KB of operand reads per SM instead of 64 + 72:
  // the shared-memory bandwidth that capped the fat shapes at 72 % tensor activity is no longer the limit.
  const int nstages = single ? H3_STAGES_1P : p.nstages;
  const uint32_t w_bytes = single ? H3_W_BYTES                                                   // one W plane of a stage,
                                  : static_cast<uint32_t>(PAIR ? p.w_rows / 2 : p.w_rows) * H3_BK * 2;   // as THIS CTA stores it
  const uint32_t stage_bytes = single ? H3_STAGE_BYTES_1P : 2 * H3_X_BYTES + 3 * w_bytes;

  // warp index made provably warp-uniform: the producer and MMA roles run their loops on all 32 lanes and predicate
  // only the issuing instructions (tc::elect_one), so that the compiler keeps TMA / UMMA operands in uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
  const int cluster = static_cast<int>(blockIdx.x) / CL;
  const int nclusters = static_cast<int>(gridDim.x) / CL;
  const int splits = p.splits > 1 ? p.splits : 1;
  const int items = p.m_blocks * p.n_tiles * splits;
  const int num_kb = p.taps > 0 ? p.taps * p.cin_blocks : (p.k + H3_BK - 1) / H3_BK;
  const int chb = p.chunk_kb;
  constexpr uint16_t kAllCtas = static_cast<uint16_t>((1u << CL) - 1u);
  const int kSliceRows = (single ? H3_BN : p.w_rows) / CL;   // W rows each CTA fetches (and multicasts)
  const uint32_t kSliceBytes = static_cast<uint32_t>(kSliceRows) * H3_BK * 2;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xlo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wa) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wb) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wc) : "memory");
    for (int s = 0; s < H3_MAX_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), PAIR ? 1 : CL);   // every CTA's tensor core must have consumed the stage (PAIR: one commit)
    }
    for (uint32_t b = 0; b < 2; ++b) {
      mbar_init(bar_cfull(b), 1);
      mbar_init(bar_cempty(b), PAIR ? 2 * H3_EPI_WARPS : H3_EPI_WARPS);   // PAIR: both CTAs' drain warps free the leader's buffer
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(H3_TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(H3_TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();            // peers' barriers are initialised before any multicast lands there
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // work item -> tile coordinates of THIS CTA
  auto k_range = [&](int item, int& kb0, int& kb1) {           // K blocks [kb0, kb1) of a work item (split-K)
    const int sp = item % splits;
    kb0 = splits > 1 ? sp * p.kb_per_split : 0;
    kb1 = splits > 1 ? min(num_kb, kb0 + p.kb_per_split) : num_kb;
    return sp;
  };
  auto tile_of = [&](int item, int& m_tile, int& grp, int& m0, int& n0) {
    item /= splits;
    const int mb = item / p.n_tiles;
    m_tile = mb * CL + static_cast<int>(rank);
    // cluster padding (M tiles beyond the last): group index = #groups, every X row out of bounds -> zero-filled
    grp = m_tile < p.m_tiles ? m_tile / p.tiles_per_batch : p.m_tiles / p.tiles_per_batch;
    m0 = m_tile < p.m_tiles ? (m_tile - grp * p.tiles_per_batch) * H3_BM : 0;
    n0 = (item - mb * p.n_tiles) * p.bn;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp; one lane issues)
    {
      int s = 0;                                                   // ring stage and the parity of its "free" phase
      uint32_t ph = 1u;
      for (int item = cluster; item < items; item += nclusters) {
        int m_tile, grp, m0, n0;
        tile_of(item, m_tile, grp, m0, n0);
        const int p_n_inst = (min(p.bn, p.n - n0) + 15) & ~15;
        // PAIR: the M = 256 instruction reads W rows [0, N/2) from the leader and [N/2, N) from its peer
        const int wrow = n0 + static_cast<int>(rank) * (PAIR ? p_n_inst / 2 : kSliceRows);
        const uint32_t wo = PAIR ? 0u : rank * kSliceBytes;
        // convolution: first output pixel of the tile -> (image, row, column); tiles beyond the last land at b >= B
        int cb0 = 0, cy0 = 0, cx0 = 0;
        if (p.taps > 0) {
          const int hw = p.out_h * p.out_w;
          const int pix = m_tile * H3_BM;
          cb0 = pix / hw;
          const int rem = pix - cb0 * hw;
          cy0 = (rem / p.out_w) * p.stride;
          cx0 = (rem - (rem / p.out_w) * p.out_w) * p.stride;
        }
        int tap = 0, cblk = 0;
        int kb0, kb1;
        k_range(item, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_empty(s), ph);
          const uint32_t st = base + s * stage_bytes;
          const bool issue = elect_one();
          // PAIR: both CTAs' loads complete on the LEADER's barrier, which expects the bytes of both stages
          const uint32_t fbar = PAIR ? map_to_rank(bar_full(s), 0u) : bar_full(s);
          if (issue && (!PAIR || rank == 0)) mbar_expect_tx(bar_full(s), PAIR ? 2 * stage_bytes : stage_bytes);
          if (p.taps > 0) {
            // shifted window of the input image for this tap; out-of-image pixels are zero-filled = zero padding
            const int ix = cx0 + p.dx[tap], iy = cy0 + p.dy[tap];
            if (issue) {
              if (PAIR) {
                tma_load_4d_pair(st, &map_xhi, fbar, cblk * H3_BK, ix, iy, cb0);
                tma_load_4d_pair(st + H3_X_BYTES, &map_xlo, fbar, cblk * H3_BK, ix, iy, cb0);
              } else {
                tma_load_4d(st, &map_xhi, bar_full(s), cblk * H3_BK, ix, iy, cb0);
                if (!single) tma_load_4d(st + H3_X_BYTES, &map_xlo, bar_full(s), cblk * H3_BK, ix, iy, cb0);
              }
            }
            if (++cblk == p.cin_blocks) { cblk = 0; ++tap; }
          } else if (issue) {
            if (PAIR) {
              tma_load_3d_pair(st, &map_xhi, fbar, kb * H3_BK, m0, grp);
              tma_load_3d_pair(st + H3_X_BYTES, &map_xlo, fbar, kb * H3_BK, m0, grp);
            } else {
              tma_load_3d(st, &map_xhi, bar_full(s), kb * H3_BK, m0, grp);
              if (!single) tma_load_3d(st + H3_X_BYTES, &map_xlo, bar_full(s), kb * H3_BK, m0, grp);
            }
          }
          if (issue) {
            if (PAIR) {                         // stage = [x_hi | x_lo | this CTA's half of A | of B | of C]
              const uint32_t w0 = st + 2 * H3_X_BYTES;
              tma_load_2d_pair(w0, &map_wa, fbar, kb * H3_BK, wrow);
              tma_load_2d_pair(w0 + w_bytes, &map_wb, fbar, kb * H3_BK, wrow);
              tma_load_2d_pair(w0 + 2 * w_bytes, &map_wc, fbar, kb * H3_BK, wrow);
            } else if (single) {                // stage = [x_hi 8 KB | w_hi 16 KB]
              const uint32_t w0 = st + H3_X_BYTES + wo;
              if (CL > 1) tma_load_2d_mc(w0, &map_wb, bar_full(s), kb * H3_BK, wrow, kAllCtas);
              else tma_load_2d(w0, &map_wb, bar_full(s), kb * H3_BK, wrow);
            } else {
              const uint32_t w0 = st + 2 * H3_X_BYTES + wo;
              if (CL > 1) {
                tma_load_2d_mc(w0, &map_wa, bar_full(s), kb * H3_BK, wrow, kAllCtas);
                tma_load_2d_mc(w0 + w_bytes, &map_wb, bar_full(s), kb * H3_BK, wrow, kAllCtas);
                tma_load_2d_mc(w0 + 2 * w_bytes, &map_wc, bar_full(s), kb * H3_BK, wrow, kAllCtas);
              } else {
                tma_load_2d(w0, &map_wa, bar_full(s), kb * H3_BK, wrow);
                tma_load_2d(w0 + w_bytes, &map_wb, bar_full(s), kb * H3_BK, wrow);
                tma_load_2d(w0 + 2 * w_bytes, &map_wc, bar_full(s), kb * H3_BK, wrow);
              }
            }
          }
          __syncwarp();
          if (++s == nstages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp; one lane issues: tcgen05.mma
    // issue blocks while the tensor core is busy, so everything else in this loop is kept to a minimum)
    {
      uint32_t cc = 0, ph = 0u;                                    // chunk counter; parity of the stage's "full" phase
      int s = 0;                                                   // ring stage
      for (int item = (PAIR && rank != 0) ? items : cluster; item < items; item += nclusters) {   // PAIR: the leader issues
        int m_tile, grp, m0, n0;
        tile_of(item, m_tile, grp, m0, n0);
        const int n_here = min(p.bn, p.n - n0);
        const int n_inst = (n_here + 15) & ~15;                 // UMMA N (multiple of 16 for M = 128)
        const uint32_t idesc = umma_idesc_f16(PAIR ? 2 * H3_BM : H3_BM, n_inst);
        int in_chunk = 0;
        uint32_t acc = tmem_base;
        int kb0, kb1;
        k_range(item, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          if (in_chunk == 0) {                                     // new chunk: the drain of chunk cc - 2 has finished
            const uint32_t buf = cc & 1u;
            mbar_wait(bar_cempty(buf), ((cc >> 1) & 1u) ^ 1u);
            tcgen05_fence_after();
            acc = tmem_base + buf * H3_BN;
          }
          mbar_wait(bar_full(s), ph);
          tcgen05_fence_after();
          const uint32_t st = base + s * stage_bytes;
          if (elect_one()) {
          if (single) {
            const uint64_t d_xhi = umma_desc_sw64(st), d_wb = umma_desc_sw64(st + H3_X_BYTES);
#pragma unroll
            for (int kk = 0; kk < H3_BK / 16; ++kk) {
              const uint64_t adv = static_cast<uint64_t>(kk * 2);
              umma_f16(acc, d_xhi + adv, d_wb + adv, idesc, (in_chunk | kk) != 0 ? 1u : 0u);
            }
          } else {
            const uint64_t d_xhi = umma_desc_sw64(st), d_xlo = umma_desc_sw64(st + H3_X_BYTES);
            const uint64_t d_wa = umma_desc_sw64(st + 2 * H3_X_BYTES);
            const uint64_t d_wb = umma_desc_sw64(st + 2 * H3_X_BYTES + w_bytes);
            const uint64_t d_wc = umma_desc_sw64(st + 2 * H3_X_BYTES + 2 * w_bytes);
#pragma unroll
            for (int kk = 0; kk < H3_BK / 16; ++kk) {
              const uint64_t adv = static_cast<uint64_t>(kk * 2);    // 16 halfs = 32 bytes = 2 x 16-byte units
              if (PAIR) {
                umma2_f16(acc, d_xhi + adv, d_wa + adv, idesc, (in_chunk | kk) != 0 ? 1u : 0u);
                umma2_f16(acc, d_xlo + adv, d_wb + adv, idesc, 1u);
                umma2_f16(acc, d_xhi + adv, d_wc + adv, idesc, 1u);
              } else {
                umma_f16(acc, d_xhi + adv, d_wa + adv, idesc, (in_chunk | kk) != 0 ? 1u : 0u);
                umma_f16(acc, d_xlo + adv, d_wb + adv, idesc, 1u);
                umma_f16(acc, d_xhi + adv, d_wc + adv, idesc, 1u);
              }
            }
          }
          if (PAIR) {
            umma2_commit_mc(bar_empty(s), kAllCtas);               // both CTAs' producers may refill the stage
            if (in_chunk + 1 == chb || kb == kb1 - 1) umma2_commit_mc(bar_cfull(cc & 1u), kAllCtas);   // both drain
          } else {
          if (CL > 1) umma_commit_mc(bar_empty(s), kAllCtas);      // stage refillable once ALL CTAs' MMAs have read it
          else umma_commit(bar_empty(s));
          if (in_chunk + 1 == chb || kb == kb1 - 1) umma_commit(bar_cfull(cc & 1u));   // chunk complete -> drain
          }
          }
          __syncwarp();
          if (++s == nstages) { s = 0; ph ^= 1u; }
          if (++in_chunk == chb || kb == kb1 - 1) {
            ++cc;
            in_chunk = 0;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ drain + epilogue (8 warps)
    const int ew = warp - 2;
    const int q = warp & 3;                                         // TMEM lane quarter this warp may read
    const int hcol = ew >> 2;                                       // which 128-column half of the tile it owns
    const float oscale = (single ? 1.f : kLoInv) * p.w_scale *      // single product: x_hi . w_hi is unscaled
                         (p.y_scale != nullptr ? __ldg(p.y_scale) : 1.f);
    // PAIR: the accumulator buffers of both CTAs are refilled by the leader's MMA warp -> free them on ITS barrier
    auto cempty_arrive = [&](uint32_t b) {
      if (PAIR) mbar_arrive_cluster(map_to_rank(bar_cempty(b), 0u));
      else mbar_arrive(bar_cempty(b));
    };
    const uint32_t box_sh = epi + static_cast<uint32_t>(ew) * H3_EPI_SLOT;
    uint8_t* box = gen + H3_STAGES * H3_STAGE_BYTES + ew * H3_EPI_SLOT;
    uint32_t cc = 0;
    for (int item = cluster; item < items; item += nclusters) {
      int m_tile, grp, m0, n0;
      tile_of(item, m_tile, grp, m0, n0);
      const int n_here = min(p.bn, p.n - n0);
      const int n_inst = (n_here + 15) & ~15;
      const bool tile_ok = m_tile < p.m_tiles;
      int ob = 0, oy = 0, ox = 0;              // convolution: (image, row, column) of this warp's first output pixel
      if (p.taps > 0) {
        const int hw = p.out_h * p.out_w;
        const int pix = m_tile * H3_BM + q * 32;
        ob = pix / hw;
        const int rem = pix - ob * hw;
        oy = rem / p.out_w;
        ox = rem - oy * p.out_w;
      }
      const int64_t lrow = static_cast<int64_t>(m0) + q * 32 + lane;            // row inside the group
      const bool row_ok = tile_ok && lrow < p.rows_per_batch;
      const int64_t row = static_cast<int64_t>(grp) * p.rows_per_batch + lrow;   // output rows are dense
      const int row0 = static_cast<int>(static_cast<int64_t>(grp) * p.rows_per_batch + m0 + q * 32);

      if (RES && tile_ok) {
        // pull this lane's residual row (128 columns x 2 planes = 4 x 128 B) into L2 now: the epilogue reads it 16 B
        // at a time, and a DRAM miss per read would serialise ~1.5 us four times per tile
        const int64_t pr = p.taps > 0 ? static_cast<int64_t>(m_tile) * H3_BM + q * 32 + lane : row;
        if (p.taps > 0 ? pr < p.rows_per_batch : row_ok) {
          const int cb = n0 + hcol * 128;
          if (cb < p.n) {
            const __half* ph = p.r_hi + pr * p.ldr + cb;
            const __half* pl = p.r_lo + pr * p.ldr + cb;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ph));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pl));
            if (cb + 64 < p.n) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(ph + 64));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(pl + 64));
            }
          }
        }
      }
      // ---- drain all chunks but the last into registers (the first one is loaded straight into them)
      float acc[128];
      int kb0, kb1;
      const int sp = k_range(item, kb0, kb1);
      const int nchunks = (kb1 - kb0 + chb - 1) / chb;
      for (int c = 0; c + 1 < nchunks; ++c, ++cc) {
        const uint32_t buf = cc & 1u;
        mbar_wait(bar_cfull(buf), (cc >> 1) & 1u);
        tcgen05_fence_after();
        const uint32_t t0 = tmem_base + buf * H3_BN + (static_cast<uint32_t>(q * 32) << 16) +
                            static_cast<uint32_t>(hcol * 128);
        if (c == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (hcol * 128 + j * 32 < n_inst) tmem_ld32(t0 + j * 32, reinterpret_cast<uint32_t*>(&acc[j * 32]));
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {                               // 16 columns at a time: registers are scarce here
            if (hcol * 128 + j * 16 < n_inst) {                        // warp-uniform
              uint32_t r[16];
              tmem_ld16(t0 + j * 16, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) acc[j * 16 + i] += __uint_as_float(r[i]);
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) cempty_arrive(buf);                 // the tensor core may refill this buffer
      }
      // ---- the last chunk is read 32 columns at a time, combined with the register sums and emitted right away
      const uint32_t lbuf = cc & 1u;
      mbar_wait(bar_cfull(lbuf), (cc >> 1) & 1u);
      tcgen05_fence_after();
      ++cc;
      const uint32_t tl = tmem_base + lbuf * H3_BN + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(hcol * 128);
      const int jlast = hcol * 128 < n_inst ? min(3, (n_inst - hcol * 128 - 1) >> 5) : -1;   // last 32-column pass of this warp
      if (jlast < 0) {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) cempty_arrive(lbuf);
      }

      // ---- epilogue: scale, bias, residual, activation, store -- 32 columns per pass, 8 at a time in registers
      const int64_t rrow = p.taps > 0 ? static_cast<int64_t>(m_tile) * H3_BM + q * 32 + lane : row;   // residual row
      const bool res_ok = RES && tile_ok && (p.taps > 0 ? rrow < p.rows_per_batch : row_ok);
      const bool relu = p.act == HOISDF_ACT_RELU;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c0 = hcol * 128 + j * 32;
        if (c0 >= n_inst) continue;                                  // warp-uniform
        const float bl = (p.bias != nullptr && sp == 0 && c0 + lane < n_here) ? __ldg(p.bias + n0 + c0 + lane) : 0.f;
        float* a = &acc[j * 32];                                    // (static indices: stays in registers)
        if (nchunks == 1) {
          tmem_ld32(tl + j * 32, reinterpret_cast<uint32_t*>(a));
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t r[16];
            tmem_ld16(tl + j * 32 + h * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) a[h * 16 + i] += __uint_as_float(r[i]);
          }
        }
        if (j == jlast) {                                            // TMEM buffer fully read: hand it back
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) cempty_arrive(lbuf);
        }
        if (!DIRECT) {
          // TMA-store paths stage the 32 x 32 block in this warp's smem slot: the previous store issued from it must
          // have finished READING it
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        }
        if (RES_STAGED) {
          // split-half residual of this warp's 32 rows x 32 columns, both planes, pulled into the staging slot with
          // COALESCED loads (4 lanes cover one row's 64 B per plane: 8 rows per instruction) in the layout the output will
          // take (64-byte rows, 16-byte chunks XOR-swizzled).  The per-lane form -- every lane reading 16 B of its own row
          // -- costs 32 L1 wavefronts per instruction and made conv3 + shortcut 3x slower than the same GEMM without it.
          const int64_t wrow0 = p.taps > 0 ? static_cast<int64_t>(m_tile) * H3_BM + q * 32 : row - lane;
          const int64_t wlim = p.taps > 0 ? static_cast<int64_t>(p.rows_per_batch)
                                          : static_cast<int64_t>(grp + 1) * p.rows_per_batch;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = (lane >> 2) + 8 * i, ch = lane & 3;
            uint4 vh = make_uint4(0u, 0u, 0u, 0u), vl = vh;
            if (tile_ok && wrow0 + r < wlim && c0 + ch * 8 < n_here) {
              vh = __ldg(reinterpret_cast<const uint4*>(p.r_hi + (wrow0 + r) * p.ldr + n0 + c0 + ch * 8));
              vl = __ldg(reinterpret_cast<const uint4*>(p.r_lo + (wrow0 + r) * p.ldr + n0 + c0 + ch * 8));
            }
            const int roff = r * 64 + ((ch ^ ((r >> 1) & 3)) << 4);
            *reinterpret_cast<uint4*>(box + roff) = vh;
            *reinterpret_cast<uint4*>(box + 2048 + roff) = vl;
          }
          __syncwarp();
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {                                  // 8 columns: c0 + 8g .. c0 + 8g + 7
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            v[i] = fmaf(a[g * 8 + i], oscale, __shfl_sync(0xffffffffu, bl, g * 8 + i));
          if (RES && res_ok && c0 + g * 8 < n_here) {
            // split-half residual (ResNet bottleneck shortcut): this lane's row, 8 columns = 16 B per plane
            uint4 a, b;
            if (RES_STAGED) {       // this lane's row from the staging slot (the output chunk g then overwrites exactly these bytes)
              const int roff = lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4);
              a = *reinterpret_cast<const uint4*>(box + roff);
              b = *reinterpret_cast<const uint4*>(box + 2048 + roff);
            } else {
              a = __ldg(reinterpret_cast<const uint4*>(p.r_hi + rrow * p.ldr + n0 + c0 + g * 8));
              b = __ldg(reinterpret_cast<const uint4*>(p.r_lo + rrow * p.ldr + n0 + c0 + g * 8));
            }
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              v[2 * i] += join_half(__ushort_as_half(static_cast<unsigned short>(aw[i] & 0xffffu)),
                                    __ushort_as_half(static_cast<unsigned short>(bw[i] & 0xffffu)));
              v[2 * i + 1] += join_half(__ushort_as_half(static_cast<unsigned short>(aw[i] >> 16)),
                                        __ushort_as_half(static_cast<unsigned short>(bw[i] >> 16)));
            }
          }
          if (DIRECT) {
            if (row_ok) {
              float* yrow = p.y + row * p.ldy + n0;
              const float* rr = p.residual ? p.residual + row * p.ldy + n0 : nullptr;
              const bool vec = ((p.ldy & 3) == 0) && aligned16(p.y) && (p.residual == nullptr || aligned16(p.residual));
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int c = c0 + g * 8 + h * 4;
                if (c >= n_here) break;
                if (vec && c + 3 < n_here) {
                  float4 o = make_float4(v[h * 4], v[h * 4 + 1], v[h * 4 + 2], v[h * 4 + 3]);
                  if (rr != nullptr) {
                    const float4 t = *reinterpret_cast<const float4*>(rr + c);
                    o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
                  }
                  if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                  *reinterpret_cast<float4*>(yrow + c) = o;
                } else {
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    if (c + i < n_here) {
                      float o = v[h * 4 + i];
                      if (rr != nullptr) o += rr[c + i];
                      if (relu) o = fmaxf(o, 0.f);
                      yrow[c + i] = o;
                    }
                  }
                }
              }
            }
            continue;
          }
          if (relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (OUT == H3_OUT_F32_TMA) {                     // 128-byte rows, 128B swizzle: two 16-byte units
            *reinterpret_cast<float4*>(box + lane * 128 + (((2 * g) ^ (lane & 7)) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(box + lane * 128 + (((2 * g + 1) ^ (lane & 7)) << 4)) =
                make_float4(v[4], v[5], v[6], v[7]);
          } else {
            // split-half output: hi box at +0, lo' box at +2048, 64-byte rows, 64B swizzle.  The single-product mode
            // writes the hi plane only (its consumers read nothing else).
            const int off = lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4);
            uint32_t hw[4], lw[4];
            if (single) {
#pragma unroll
              for (int i = 0; i < 4; ++i) hw[i] = cvt_f16x2_sat(v[2 * i], v[2 * i + 1]);
              *reinterpret_cast<uint4*>(box + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) split_half2(v[2 * i], v[2 * i + 1], hw[i], lw[i]);
              *reinterpret_cast<uint4*>(box + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(box + 2048 + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          }
        }
        if (DIRECT) continue;
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (tile_ok && c0 < n_here) {
            constexpr bool two = OUT == H3_OUT_SPLIT_TMA && !single;
            if (p.taps > 0) {
              tma_store_4d(&map_y0, box_sh, n0 + c0, ox, oy, ob);
              if (two) tma_store_4d(&map_y1, box_sh + 2048, n0 + c0, ox, oy, ob);
            } else {
              if (OUT == H3_OUT_F32_TMA && splits > 1) tma_reduce_add_2d(&map_y0, box_sh, n0 + c0, row0);
              else tma_store_2d(&map_y0, box_sh, n0 + c0, row0);
              if (two) tma_store_2d(&map_y1, box_sh + 2048, n0 + c0, row0);
            }
          }
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_read<0>();     // smem sources consumed; kernel end flushes the global writes
    __syncwarp();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();              // no CTA exits while a peer may still multicast to it / signal its barriers
  if (warp == 1) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(H3_TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(H3_TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// format conversion kernels
// ---------------------------------------------------------------------------------------------------
// W (n, ldw fp32, k valid columns) -> planes A = fp16(w * 2^11), B = fp16(A * 2^-11), C = fp16((w - A * 2^-11) * 2^11),
// each (n, ldh halfs) with columns [k, ldh) zeroed
__global__ void pack_h3_kernel(const float* __restrict__ w, int64_t n, int64_t k, int64_t ldw, __half* __restrict__ a,
                               __half* __restrict__ b, __half* __restrict__ c, int64_t ldh) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n * ldh) return;
  const int64_t r = i / ldh, col = i - r * ldh;
  float v = col < k ? w[r * ldw + col] : 0.f;
  const float vs = fminf(fmaxf(v * kLoScale, -65504.f), 65504.f);
  const __half ha = __float2half_rn(vs);
  const float whi = __half2float(ha) * kLoInv;           // exact (power-of-two scaling)
  a[i] = ha;
  b[i] = __float2half_rn(whi);
  c[i] = __float2half_rn(fminf(fmaxf((v - whi) * kLoScale, -65504.f), 65504.f));
}

// X (m, ldx fp32, k valid columns) -> split-half planes (m, ldh halfs), columns [k, kpad) zeroed; 4 columns per thread
__global__ void split_rows_kernel(const float* __restrict__ x, int64_t m, int64_t k, int64_t ldx, int64_t kpad,
                                  __half* __restrict__ hi, __half* __restrict__ lo, int64_t ldh, int vec) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t per_row = kpad >> 2;
  if (i >= m * per_row) return;
  const int64_t r = i / per_row, c = (i - r * per_row) << 2;
  float v[4];
  if (vec && c + 3 < k) {
    const float4 f = *reinterpret_cast<const float4*>(x + r * ldx + c);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = c + j < k ? x[r * ldx + c + j] : 0.f;
  }
  __half h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_half(v[j], h[j], l[j]);
  uint2 ph, pl;
  ph.x = static_cast<uint32_t>(__half_as_ushort(h[0])) | (static_cast<uint32_t>(__half_as_ushort(h[1])) << 16);
  ph.y = static_cast<uint32_t>(__half_as_ushort(h[2])) | (static_cast<uint32_t>(__half_as_ushort(h[3])) << 16);
  pl.x = static_cast<uint32_t>(__half_as_ushort(l[0])) | (static_cast<uint32_t>(__half_as_ushort(l[1])) << 16);
  pl.y = static_cast<uint32_t>(__half_as_ushort(l[2])) | (static_cast<uint32_t>(__half_as_ushort(l[3])) << 16);
  *reinterpret_cast<uint2*>(hi + r * ldh + c) = ph;
  *reinterpret_cast<uint2*>(lo + r * ldh + c) = pl;
}

// split-half planes -> fp32 (m, k)
__global__ void join_rows_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, int64_t ldh, int64_t m,
                                 int64_t k, float* __restrict__ x, int64_t ldx) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= m * k) return;
  const int64_t r = i / k, c = i - r * k;
  x[r * ldx + c] = join_half(hi[r * ldh + c], lo[r * ldh + c]);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static bool map_half_2d(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                        int box_rows, CUtensorMapSwizzle sw, CUtensorMapL2promotion l2) {
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  return make_tiled_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptr, dims, strides, box, estr, sw, l2);
}

// X plane as (groups, rows_per_group, K): box = 1 x 128 rows x 32 halfs
static bool map_x_3d(CUtensorMap* map, const void* ptr, int64_t groups, int64_t rows, int64_t cols, int64_t ld,
                     int64_t group_stride) {
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(groups)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(group_stride) * 2};
  cuuint32_t box[3] = {H3_BK, H3_BM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return make_tiled_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, ptr, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

template <int CL>
static int max_clusters() {
  static int cached = -1;
  if (cached >= 0) return cached;
  auto kern = linear_h3_kernel<CL, H3_OUT_SPLIT_TMA, false, false, false>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, H3_SMEM_BYTES);
  int n = 0;
  if (CL == 1) {
    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    n = cudaGetDeviceProperties(&prop, dev) == cudaSuccess ? prop.multiProcessorCount : kNumSMs;
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(kNumSMs / CL * CL);
    cfg.blockDim = dim3(H3_THREADS);
    cfg.dynamicSmemBytes = H3_SMEM_BYTES;
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

static int g_h3_force_cluster = 0;   // developer hook: 0 = automatic

// W rows per stage plane for a layer with `n` outputs (see H3Params::w_rows); HOISDF_H3_WIDE_STAGES=1 keeps 256 always
static int h3_w_rows(int64_t n, bool single) {
  static const bool wide_only = [] { const char* e = getenv("HOISDF_H3_WIDE_STAGES"); return e != nullptr && atoi(e) != 0; }();
  if (single || wide_only || n > 128) return H3_BN;
  return n > 64 ? 128 : 64;
}
static int h3_stage_count(int w_rows) {
  const int stage = 2 * H3_X_BYTES + 3 * w_rows * H3_BK * 2;
  const int n = H3_STAGES * H3_STAGE_BYTES / stage;
  return n < H3_MAX_STAGES ? n : H3_MAX_STAGES;
}
static int h3_stage_count_pair(int w_rows) {        // cta_group::2: each CTA stages half of the W rows
  const int stage = 2 * H3_X_BYTES + 3 * (w_rows / 2) * H3_BK * 2;
  const int n = H3_STAGES * H3_STAGE_BYTES / stage;
  return n < H3_MAX_STAGES ? n : H3_MAX_STAGES;
}
static int g_h3_chunk_kb = H3_CHUNK_KB;   // developer hook: K blocks per accumulation chunk
// tcgen05.mma.cta_group::2 for launches that run as clusters of two CTAs: 0 never, 1 where it pays (launch_h3), 2 always
// split-K for few-tile / long-contraction GEMMs (hoisdf_linear_h3_fwd): HOISDF_H3_SPLITK=0 turns it off
static int g_h3_splitk = [] { const char* e = getenv("HOISDF_H3_SPLITK"); return e == nullptr ? 1 : atoi(e); }();
static int g_h3_pair = [] { const char* e = getenv("HOISDF_H3_PAIR"); return e == nullptr ? 1 : atoi(e); }();

template <int CL, int OUT, bool SINGLE, bool RES, bool PAIR = false>
static int launch_h3_inst(const CUtensorMap* maps, const H3Params& p, int64_t clusters, cudaStream_t s) {
  auto kern = linear_h3_kernel<CL, OUT, SINGLE, RES, PAIR>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, H3_SMEM_BYTES);
  if (e != cudaSuccess) return static_cast<int>(e);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(clusters * CL));
  cfg.blockDim = dim3(H3_THREADS);
  cfg.dynamicSmemBytes = H3_SMEM_BYTES;
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
static int launch_h3(const CUtensorMap* maps, const H3Params& p0, int64_t m_tiles, cudaStream_t s) {
  H3Params p = p0;
  if (p.chunk_kb <= 0) p.chunk_kb = g_h3_chunk_kb > 0 ? g_h3_chunk_kb : H3_CHUNK_KB;
  p.m_blocks = static_cast<int>(ceil_div(m_tiles, CL));
  const int64_t items = static_cast<int64_t>(p.m_blocks) * p.n_tiles * (p.splits > 1 ? p.splits : 1);
  const int64_t clusters = items < max_clusters<CL>() ? items : max_clusters<CL>();
  const bool res = p.r_hi != nullptr;
  if (p.single && (res || p.out_mode == H3_OUT_F32_DIRECT)) return HOISDF_E_UNSUPPORTED;
  if constexpr (CL == 2) {
    // tcgen05.mma.cta_group::2 (half of W per CTA, deeper ring) pays where the K loop is long and wide enough to be bound by
    // shared-memory bandwidth: measured +8 .. +20 % for N >= 256, K >= 1024 (U-Net 3x3 / transposed convolutions, the point
    // MLP, FFN2), but the cross-CTA hand-overs cost 10 .. 80 % on thin / short-K / drain-every-K-block launches
    // (profiles/r02s_h3_pair_vs_single.txt).  HOISDF_H3_PAIR: 0 never, 1 automatic (default), 2 always.
    const int64_t kk = p.taps > 0 ? static_cast<int64_t>(p.taps) * p.cin_blocks * H3_BK : p.k;
    const bool pays = p.n >= 256 && p.bn == H3_BN && kk >= 1024 && p.chunk_kb != 1;
    if ((g_h3_pair == 2 || (g_h3_pair == 1 && pays)) && !p.single) {
      p.nstages = h3_stage_count_pair(p.w_rows);
      switch (p.out_mode) {
        case H3_OUT_F32_TMA:
          return res ? launch_h3_inst<CL, H3_OUT_F32_TMA, false, true, true>(maps, p, clusters, s)
                     : launch_h3_inst<CL, H3_OUT_F32_TMA, false, false, true>(maps, p, clusters, s);
        case H3_OUT_SPLIT_TMA:
          return res ? launch_h3_inst<CL, H3_OUT_SPLIT_TMA, false, true, true>(maps, p, clusters, s)
                     : launch_h3_inst<CL, H3_OUT_SPLIT_TMA, false, false, true>(maps, p, clusters, s);
        default:
          if (res) return HOISDF_E_UNSUPPORTED;
          return launch_h3_inst<CL, H3_OUT_F32_DIRECT, false, false, true>(maps, p, clusters, s);
      }
    }
  }
  switch (p.out_mode) {
    case H3_OUT_F32_TMA:
      if (p.single) return launch_h3_inst<CL, H3_OUT_F32_TMA, true, false>(maps, p, clusters, s);
      return res ? launch_h3_inst<CL, H3_OUT_F32_TMA, false, true>(maps, p, clusters, s)
                 : launch_h3_inst<CL, H3_OUT_F32_TMA, false, false>(maps, p, clusters, s);
    case H3_OUT_SPLIT_TMA:
      if (p.single) return launch_h3_inst<CL, H3_OUT_SPLIT_TMA, true, false>(maps, p, clusters, s);
      return res ? launch_h3_inst<CL, H3_OUT_SPLIT_TMA, false, true>(maps, p, clusters, s)
                 : launch_h3_inst<CL, H3_OUT_SPLIT_TMA, false, false>(maps, p, clusters, s);
    default:
      if (res) return HOISDF_E_UNSUPPORTED;
      return launch_h3_inst<CL, H3_OUT_F32_DIRECT, false, false>(maps, p, clusters, s);
  }
}

}  // namespace hoisdf

using namespace hoisdf;

extern "C" __attribute__((visibility("default"))) void hoisdf_debug_h3_cluster(int cl) { g_h3_force_cluster = cl; }
extern "C" __attribute__((visibility("default"))) void hoisdf_debug_h3_chunk(int kb) { g_h3_chunk_kb = kb; }
extern "C" __attribute__((visibility("default"))) void hoisdf_debug_h3_pair(int mode) { g_h3_pair = mode; }

HOISDF_API int hoisdf_linear_h3_fwd(const hoisdf_linear_h3_args* a, void* stream) {
  if (a == nullptr || a->x_hi == nullptr || a->x_lo == nullptr || a->w_a == nullptr || a->w_b == nullptr ||
      a->w_c == nullptr)
    return HOISDF_E_NULL;
  const bool split_out = a->y_hi != nullptr || a->y_lo != nullptr;
  if (split_out ? (a->y_hi == nullptr || a->y_lo == nullptr || a->y != nullptr) : a->y == nullptr) return HOISDF_E_NULL;
  if (a->m == 0) return HOISDF_OK;
  if (a->m < 0 || a->n <= 0 || a->k <= 0 || a->m >= 0x7fffffffLL || a->n >= 0x7fffffffLL || a->k >= 0x7fffffffLL)
    return HOISDF_E_SHAPE;
  if ((a->ldx & 7) || (a->ldw & 7) || !aligned16(a->x_hi) || !aligned16(a->x_lo) || !aligned16(a->w_a) ||
      !aligned16(a->w_b) || !aligned16(a->w_c) || a->ldx < a->k || a->ldw < a->k)
    return HOISDF_E_ALIGN;
  if (split_out && (a->residual != nullptr || (a->ldyh & 7) || !aligned16(a->y_hi) || !aligned16(a->y_lo)))
    return a->residual != nullptr ? HOISDF_E_UNSUPPORTED : HOISDF_E_ALIGN;
  int64_t groups = 1, rpb = a->m, gstride = a->m * a->ldx;
  if (a->x_rows_per_batch > 0) {
    if (a->m % a->x_rows_per_batch != 0) return HOISDF_E_SHAPE;
    rpb = a->x_rows_per_batch;
    groups = a->m / rpb;
    gstride = a->x_batch_stride;
    if (groups > 1 && (gstride & 7)) return HOISDF_E_ALIGN;
  }
  if (groups == 1) gstride = rpb * a->ldx;
  const bool tile_safe = groups == 1 || rpb % H3_BM == 0;     // tiles never straddle two row groups
  int out_mode;
  if (split_out) {
    if (!tile_safe) return HOISDF_E_UNSUPPORTED;
    out_mode = H3_OUT_SPLIT_TMA;
  } else {
    out_mode = (a->residual == nullptr && (a->ldy & 3) == 0 && aligned16(a->y) && tile_safe) ? H3_OUT_F32_TMA
                                                                                             : H3_OUT_F32_DIRECT;
  }
  const int64_t tpb = ceil_div(rpb, H3_BM);
  const int64_t m_tiles = groups * tpb;
  int cl = m_tiles >= 2 ? 2 : 1;   // measured: pairs beat quads (quads leave SMs idle and add lockstep stalls)
  if (g_h3_force_cluster == 1 || g_h3_force_cluster == 2) cl = g_h3_force_cluster;
  CUtensorMap maps[7];
  if (!map_x_3d(&maps[0], a->x_hi, groups, rpb, a->k, a->ldx, gstride)) return HOISDF_E_UNSUPPORTED;
  if (!map_x_3d(&maps[1], a->x_lo, groups, rpb, a->k, a->ldx, gstride)) return HOISDF_E_UNSUPPORTED;
  // launches with very few tiles (the 17-query decoder layers, M = 544): N tiles of 64 columns instead of 256 -> 4x the
  // CTAs and a 6-stage ring, the K loop is pure load latency there
  // split-K: a handful of output tiles over a long contraction (the weight gradients of the training step)
  int splits = 1, kb_per_split = 0;
  {
    const int64_t num_kb = ceil_div(a->k, H3_BK);
    const int64_t items0 = ceil_div(m_tiles, cl) * ceil_div(a->n, h3_w_rows(a->n, false));
    const int64_t room = (kNumSMs / cl) / items0;
    if (g_h3_splitk && a->split_k != 0 && out_mode == H3_OUT_F32_TMA && a->single_pass == 0 && a->act == HOISDF_ACT_NONE &&
        a->res_hi == nullptr && a->res_lo == nullptr && groups == 1 && room >= 2 && num_kb >= 64) {
      const int64_t want = room < num_kb / 8 ? room : num_kb / 8;          // >= 8 K blocks (256 of K) per item
      const int64_t per = ceil_div(ceil_div(num_kb, want), 4) * 4;         // whole TMEM chunks
      splits = static_cast<int>(ceil_div(num_kb, per));
      kb_per_split = static_cast<int>(per);
      if (splits < 2) splits = 1;
    }
  }
  const bool few = splits == 1 && a->single_pass == 0 && a->n > 64 && m_tiles * ceil_div(a->n, H3_BN) <= kNumSMs / 4;
  const int w_rows = few ? 64 : h3_w_rows(a->n, a->single_pass != 0);
  const int bn = w_rows;                                 // three-product mode: one N tile = the W rows of a stage
  const int slice = w_rows / cl;
  const void* wp[3] = {a->w_a, a->w_b, a->w_c};
  for (int i = 0; i < 3; ++i)
    if (!map_half_2d(&maps[2 + i], wp[i], a->n, a->k, a->ldw, H3_BK, slice, CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
      return HOISDF_E_UNSUPPORTED;
  maps[5] = maps[0];
  maps[6] = maps[0];
  if (out_mode == H3_OUT_F32_TMA) {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(a->n), static_cast<cuuint64_t>(a->m)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(a->ldy) * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    if (!make_tiled_map(&maps[5], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a->y, dims, strides, box, estr,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE))
      return HOISDF_E_UNSUPPORTED;
  } else if (out_mode == H3_OUT_SPLIT_TMA) {
    if (!map_half_2d(&maps[5], a->y_hi, a->m, a->n, a->ldyh, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE) ||
        !map_half_2d(&maps[6], a->y_lo, a->m, a->n, a->ldyh, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE))
      return HOISDF_E_UNSUPPORTED;
  }
  H3Params p{};
  p.bias = a->bias; p.residual = a->residual; p.y = a->y; p.ldy = a->ldy;
  p.rows_per_batch = rpb; p.tiles_per_batch = static_cast<int>(tpb); p.m_tiles = static_cast<int>(m_tiles);
  p.bn = a->single_pass ? H3_BN : bn;
  p.n_tiles = static_cast<int>(ceil_div(a->n, p.bn));
  p.n = static_cast<int>(a->n); p.k = static_cast<int>(a->k); p.act = a->act; p.out_mode = out_mode;
  p.chunk_kb = a->chunk_kb;
  p.single = a->single_pass ? 1 : 0;
  p.w_rows = w_rows; p.nstages = h3_stage_count(w_rows);
  p.w_scale = a->w_scale > 0.f ? a->w_scale : 1.f;
  p.y_scale = a->y_scale;
  p.splits = splits; p.kb_per_split = kb_per_split;
  if (a->res_hi != nullptr || a->res_lo != nullptr) {
    if (a->res_hi == nullptr || a->res_lo == nullptr) return HOISDF_E_NULL;
    if (out_mode == H3_OUT_F32_DIRECT || a->residual != nullptr || (a->n & 31)) return HOISDF_E_UNSUPPORTED;
    if ((a->ldr & 7) || a->ldr < a->n || !aligned16(a->res_hi) || !aligned16(a->res_lo)) return HOISDF_E_ALIGN;
    p.r_hi = reinterpret_cast<const __half*>(a->res_hi);
    p.r_lo = reinterpret_cast<const __half*>(a->res_lo);
    p.ldr = a->ldr;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (splits > 1) {                            // the items add into the output
    const cudaError_t e = cudaMemset2DAsync(a->y, static_cast<size_t>(a->ldy) * 4, 0, static_cast<size_t>(a->n) * 4,
                                            static_cast<size_t>(a->m), s);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  if (cl == 2) return launch_h3<2>(maps, p, m_tiles, s);
  return launch_h3<1>(maps, p, m_tiles, s);
}

HOISDF_API int hoisdf_conv_h3_fwd(const hoisdf_conv_h3_args* a, void* stream) {
  if (a == nullptr || a->x_hi == nullptr || a->x_lo == nullptr || a->w_a == nullptr || a->w_b == nullptr ||
      a->w_c == nullptr)
    return HOISDF_E_NULL;
  const bool split_out = a->y_hi != nullptr || a->y_lo != nullptr;
  if (split_out ? (a->y_hi == nullptr || a->y_lo == nullptr || a->y != nullptr) : a->y == nullptr) return HOISDF_E_NULL;
  if (a->batch <= 0 || a->in_h <= 0 || a->in_w <= 0 || a->cin <= 0 || a->cout <= 0 || a->out_h <= 0 || a->out_w <= 0 ||
      a->taps <= 0 || a->taps > 16 || a->stride < 1 || a->stride > 2)
    return HOISDF_E_SHAPE;
  if (a->cin % H3_BK != 0) return HOISDF_E_UNSUPPORTED;         // a K block never straddles two taps
  const int64_t m = a->batch * a->out_h * a->out_w;
  if (m >= 0x7fffffffLL || a->taps * a->cin >= 0x7fffffffLL) return HOISDF_E_SHAPE;
  if ((a->ldx & 7) || (a->ldw & 7) || !aligned16(a->x_hi) || !aligned16(a->x_lo) || !aligned16(a->w_a) ||
      !aligned16(a->w_b) || !aligned16(a->w_c) || a->ldx < a->cin || a->ldw < a->taps * a->cin)
    return HOISDF_E_ALIGN;
  const int esz = split_out ? 2 : 4;
  if ((a->y_sx * esz) % 16 || (a->y_sy * esz) % 16 || (a->y_sb * esz) % 16) return HOISDF_E_ALIGN;
  if (split_out ? (!aligned16(a->y_hi) || !aligned16(a->y_lo)) : !aligned16(a->y)) return HOISDF_E_ALIGN;
  // tile geometry: 128 consecutive output pixels = tb images x ty rows x tx columns
  const int tx = static_cast<int>(a->out_w < H3_BM ? a->out_w : H3_BM);
  if (a->out_w % tx != 0 || H3_BM % tx != 0) return HOISDF_E_UNSUPPORTED;
  const int ty_max = H3_BM / tx;
  const int ty = static_cast<int>(a->out_h < ty_max ? a->out_h : ty_max);
  if (a->out_h % ty != 0 || ty_max % ty != 0) return HOISDF_E_UNSUPPORTED;
  const int tb = H3_BM / (tx * ty);
  if (tx * ty < 32 || tx * a->stride > 256 || ty * a->stride > 256) return HOISDF_E_UNSUPPORTED;
  const int wx = tx < 32 ? tx : 32, wy = 32 / wx;
  CUtensorMap maps[7];
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(a->cin), static_cast<cuuint64_t>(a->in_w),
                          static_cast<cuuint64_t>(a->in_h), static_cast<cuuint64_t>(a->batch)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(a->ldx) * 2, static_cast<cuuint64_t>(a->in_w * a->ldx) * 2,
                             static_cast<cuuint64_t>(a->in_h * a->in_w * a->ldx) * 2};
    cuuint32_t box[4] = {H3_BK, static_cast<cuuint32_t>(tx * a->stride), static_cast<cuuint32_t>(ty * a->stride),
                         static_cast<cuuint32_t>(tb)};
    cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(a->stride), static_cast<cuuint32_t>(a->stride), 1};
    if (!make_tiled_map(&maps[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, a->x_hi, dims, strides, box, estr,
                        CU_TENSOR_MAP_SWIZZLE_64B) ||
        !make_tiled_map(&maps[1], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, a->x_lo, dims, strides, box, estr,
                        CU_TENSOR_MAP_SWIZZLE_64B))
      return HOISDF_E_UNSUPPORTED;
  }
  const int64_t m_tiles = ceil_div(m, H3_BM);
  int cl = m_tiles >= 2 ? 2 : 1;
  if (g_h3_force_cluster == 1 || g_h3_force_cluster == 2) cl = g_h3_force_cluster;
  const void* wp[3] = {a->w_a, a->w_b, a->w_c};
  const int w_rows = h3_w_rows(a->cout, a->single_pass != 0);
  for (int i = 0; i < 3; ++i)
    if (!map_half_2d(&maps[2 + i], wp[i], a->cout, a->taps * a->cin, a->ldw, H3_BK, w_rows / cl,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
      return HOISDF_E_UNSUPPORTED;
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(a->cout), static_cast<cuuint64_t>(a->out_w),
                          static_cast<cuuint64_t>(a->out_h), static_cast<cuuint64_t>(a->batch)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(a->y_sx) * esz, static_cast<cuuint64_t>(a->y_sy) * esz,
                             static_cast<cuuint64_t>(a->y_sb) * esz};
    cuuint32_t box[4] = {32, static_cast<cuuint32_t>(wx), static_cast<cuuint32_t>(wy), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (split_out) {
      if (!make_tiled_map(&maps[5], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, a->y_hi, dims, strides, box, estr,
                          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE) ||
          !make_tiled_map(&maps[6], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, a->y_lo, dims, strides, box, estr,
                          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE))
        return HOISDF_E_UNSUPPORTED;
    } else {
      if (!make_tiled_map(&maps[5], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->y, dims, strides, box, estr,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE))
        return HOISDF_E_UNSUPPORTED;
      maps[6] = maps[5];
    }
  }
  H3Params p{};
  p.bias = a->bias; p.rows_per_batch = m; p.tiles_per_batch = static_cast<int>(m_tiles);
  p.bn = a->single_pass ? H3_BN : w_rows;
  p.m_tiles = static_cast<int>(m_tiles); p.n_tiles = static_cast<int>(ceil_div(a->cout, p.bn));
  p.n = static_cast<int>(a->cout); p.k = static_cast<int>(a->taps * a->cin); p.act = a->act;
  p.out_mode = split_out ? H3_OUT_SPLIT_TMA : H3_OUT_F32_TMA;
  p.chunk_kb = a->chunk_kb;
  p.single = a->single_pass ? 1 : 0;
  p.w_rows = w_rows; p.nstages = h3_stage_count(w_rows);
  p.w_scale = a->w_scale > 0.f ? a->w_scale : 1.f;
  p.taps = a->taps; p.cin_blocks = static_cast<int>(a->cin / H3_BK);
  p.out_w = static_cast<int>(a->out_w); p.out_h = static_cast<int>(a->out_h); p.stride = a->stride;
  p.wx = wx; p.wy = wy;
  for (int t = 0; t < a->taps; ++t) {
    if (a->tap_dy[t] < -64 || a->tap_dy[t] > 64 || a->tap_dx[t] < -64 || a->tap_dx[t] > 64) return HOISDF_E_SHAPE;
    p.dy[t] = static_cast<int8_t>(a->tap_dy[t]);
    p.dx[t] = static_cast<int8_t>(a->tap_dx[t]);
  }
  if (a->res_hi != nullptr || a->res_lo != nullptr) {
    if (a->res_hi == nullptr || a->res_lo == nullptr) return HOISDF_E_NULL;
    if (a->cout & 31) return HOISDF_E_UNSUPPORTED;
    if ((a->ldr & 7) || a->ldr < a->cout || !aligned16(a->res_hi) || !aligned16(a->res_lo)) return HOISDF_E_ALIGN;
    p.r_hi = reinterpret_cast<const __half*>(a->res_hi);
    p.r_lo = reinterpret_cast<const __half*>(a->res_lo);
    p.ldr = a->ldr;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (cl == 2) return launch_h3<2>(maps, p, m_tiles, s);
  return launch_h3<1>(maps, p, m_tiles, s);
}

HOISDF_API int hoisdf_pack_h3(const float* w, int64_t n, int64_t k, int64_t ldw, uint16_t* w_a, uint16_t* w_b,
                              uint16_t* w_c, int64_t ldh, void* stream) {
  if (w == nullptr || w_a == nullptr || w_b == nullptr || w_c == nullptr) return HOISDF_E_NULL;
  if (n <= 0 || k <= 0 || ldh < k || ldw < k) return HOISDF_E_SHAPE;
  const int64_t total = n * ldh;
  pack_h3_kernel<<<static_cast<unsigned>(ceil_div(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, n, k, ldw, reinterpret_cast<__half*>(w_a), reinterpret_cast<__half*>(w_b), reinterpret_cast<__half*>(w_c), ldh);
  return launch_status();
}

HOISDF_API int hoisdf_split_rows(const float* x, int64_t m, int64_t k, int64_t ldx, int64_t kpad, uint16_t* hi,
                                 uint16_t* lo, int64_t ldh, void* stream) {
  if (x == nullptr || hi == nullptr || lo == nullptr) return HOISDF_E_NULL;
  if (m == 0) return HOISDF_OK;
  if (m < 0 || k <= 0 || kpad < k || (kpad & 3) || ldh < kpad || (ldh & 3) || ldx < k) return HOISDF_E_SHAPE;
  if ((reinterpret_cast<uintptr_t>(hi) & 7) || (reinterpret_cast<uintptr_t>(lo) & 7)) return HOISDF_E_ALIGN;
  const int vec = ((ldx & 3) == 0 && aligned16(x)) ? 1 : 0;
  const int64_t total = m * (kpad >> 2);
  split_rows_kernel<<<static_cast<unsigned>(ceil_div(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, m, k, ldx, kpad, reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo), ldh, vec);
  return launch_status();
}

HOISDF_API int hoisdf_join_rows(const uint16_t* hi, const uint16_t* lo, int64_t ldh, int64_t m, int64_t k, float* x,
                                int64_t ldx, void* stream) {
  if (x == nullptr || hi == nullptr || lo == nullptr) return HOISDF_E_NULL;
  if (m == 0) return HOISDF_OK;
  if (m < 0 || k <= 0 || ldh < k || ldx < k) return HOISDF_E_SHAPE;
  join_rows_kernel<<<static_cast<unsigned>(ceil_div(m * k, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(hi), reinterpret_cast<const __half*>(lo), ldh, m, k, x, ldx);
  return launch_status();
}
