// 2-SM (cta_group::2) variant of the 3xTF32 tensor-core Linear (numerics and the 1-SM kernel: linear_tc.cu).
//
// Why: the 1-SM kernel is bound by operand delivery -- every SM ingests its 128 X rows AND the whole 256-row W tile
// (hi + lo) per K slice: 26 FLOP per byte crossing the L2 -> SM fabric, which saturates at ~11.4 TB/s aggregate
// (measured with the CTA timeline trace).  Here a PAIR of SMs computes a 256 x 256 output tile
// with ONE tcgen05.mma.cta_group::2 (M = 256): each CTA stages only its own 128 X rows and HALF of the W rows, and the
// tensor cores of both SMs read both halves.  Per SM that is 24 KB instead of 40 KB per K slice (43.7 FLOP/B), half
// the shared-memory operand reads, and a 32 KB stage, i.e. a 6-deep ring instead of 4.
//
// Roles per CTA (192 threads), barriers marked (L) live in the leader CTA (cluster rank 0):
//   warp 0      TMA producer: own X rows -> local `x_full`; own half of W_hi / W_lo -> `w_full` (L) (cta_group::2 TMA)
//   warps 2..5  wait `x_full`, split X in place into x_hi / x_lo, arrive on `conv` (L) (remote arrive for rank 1),
//               later the epilogue of the CTA's own 128 rows
//   warp 1      leader only: waits `w_full` + `conv`, issues the M256 x N x K8 MMAs (main / corr accumulators in both
//               CTAs' TMEM), tcgen05.commit multicast frees the stage in both CTAs / signals both epilogues
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"

namespace hoisdf {

namespace tc2 {

constexpr int BM = 128;                 // rows per CTA (256 per pair)
constexpr int BN = 256;
constexpr int BK = 16;
constexpr int STAGES = 6;
constexpr int A_BYTES = BM * BK * 4;            // 8 KB
constexpr int BH_ROWS = BN / 2;                 // W rows staged per CTA
constexpr int BH_BYTES = BH_ROWS * BK * 4;      // 8 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * BH_BYTES;   // x_hi x_lo w_hi/2 w_lo/2 = 32 KB
constexpr int BAR_BYTES = 512;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + BAR_BYTES + BN * 4;
constexpr int THREADS = 192;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t kSpin = 1u << 27;

struct Params {
  const float* __restrict__ bias;
  const float* __restrict__ residual;
  float* __restrict__ y;
  int64_t ldy;
  int64_t rows_per_batch;
  int tiles_per_batch;
  int m_tiles;
  int n, k, act;
  int passes;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > kSpin) __trap();
  }
}
__device__ __forceinline__ void tma_x_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// destination: this CTA's shared memory; completion: the LEADER's barrier (cluster address)
__device__ __forceinline__ void tma_w_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ uint64_t desc_sw64(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (static_cast<uint64_t>(512 >> 4) << 32) |
         (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(z)
      : "memory");
}
__device__ __forceinline__ void commit2_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask)
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

__global__ void __launch_bounds__(THREADS, 1)
linear_tf32x3_2sm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_whi,
                         const __grid_constant__ CUtensorMap map_wlo, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  // x_full[S] | w_full[S] | conv[S] | empty[S] | done | tmem slot
  auto bar_xf = [&](int s) { return bars + 8u * s; };
  auto bar_wf = [&](int s) { return bars + 8u * (STAGES + s); };
  auto bar_cv = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto bar_em = [&](int s) { return bars + 8u * (3 * STAGES + s); };
  const uint32_t bar_done = bars + 8u * (4 * STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + STAGES * STAGE_BYTES + 8 * (4 * STAGES + 1));
  float* bias_s = reinterpret_cast<float*>(gen + STAGES * STAGE_BYTES + BAR_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const bool leader = rank == 0;
  const int n_tiles = (p.n + BN - 1) / BN;
  const int cid = static_cast<int>(blockIdx.x) >> 1;
  const int m_tile = (cid / n_tiles) * 2 + static_cast<int>(rank);
  const int grp = m_tile < p.m_tiles ? m_tile / p.tiles_per_batch : p.m_tiles / p.tiles_per_batch;
  const int m0 = m_tile < p.m_tiles ? (m_tile - grp * p.tiles_per_batch) * BM : 0;
  const int n0 = (cid % n_tiles) * BN;
  const int n_here = min(BN, p.n - n0);
  const int n_inst = (n_here + 15) & ~15;
  const int n_half = n_inst >> 1;                    // W rows owned by each CTA of the pair
  const int num_kb = (p.k + BK - 1) / BK;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_whi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wlo) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_xf(s), 1);
      mbar_init(bar_wf(s), 1);
      mbar_init(bar_cv(s), 8);      // 4 converter warps in each of the two CTAs
      mbar_init(bar_em(s), 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t w_bytes_total = 2u * (p.passes == 3 ? 2u : 1u) * BH_BYTES;   // both CTAs' halves
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_em(s), ph ^ 1u);
        const uint32_t st = base + s * STAGE_BYTES;
        mbar_expect_tx(bar_xf(s), A_BYTES);
        tma_x_3d(st, &map_x, bar_xf(s), kb * BK, m0, grp);
        if (leader) mbar_expect_tx(bar_wf(s), w_bytes_total);
        const uint32_t leader_wf = map_to_rank(bar_wf(s), 0);
        const int wrow = n0 + static_cast<int>(rank) * n_half;
        tma_w_2sm(st + 2 * A_BYTES, &map_whi, leader_wf, kb * BK, wrow);
        if (p.passes == 3) tma_w_2sm(st + 2 * A_BYTES + BH_BYTES, &map_wlo, leader_wf, kb * BK, wrow);
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      // D = F32, A = B = TF32, K-major, N >> 3 at bit 17, M = 256 -> 16 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n_inst >> 3) << 17) |
                             (static_cast<uint32_t>(256 >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_wf(s), ph);
        mbar_wait(bar_cv(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = base + s * STAGE_BYTES;
        const uint64_t d_xhi = desc_sw64(st), d_xlo = desc_sw64(st + A_BYTES);
        const uint64_t d_whi = desc_sw64(st + 2 * A_BYTES), d_wlo = desc_sw64(st + 2 * A_BYTES + BH_BYTES);
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {
          const uint64_t adv = static_cast<uint64_t>(kk * 2);
          const uint32_t acc = (kb | kk) != 0 ? 1u : 0u;
          umma2_tf32(tmem_acc, d_xhi + adv, d_whi + adv, idesc, acc);
          if (p.passes == 3) {
            umma2_tf32(tmem_acc + BN, d_xlo + adv, d_whi + adv, idesc, acc);
            umma2_tf32(tmem_acc + BN, d_xhi + adv, d_wlo + adv, idesc, 1u);
          }
        }
        commit2_mc(bar_em(s), 3);
      }
      commit2_mc(bar_done, 3);
    }
  } else {
    const int t = threadIdx.x - 64;
    for (int c = t; c < BN; c += 128) bias_s[c] = (p.bias != nullptr && c < n_here) ? __ldg(p.bias + n0 + c) : 0.f;
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(bar_xf(s), ph);
      if (p.passes == 3) {
        float4* hi = reinterpret_cast<float4*>(gen + s * STAGE_BYTES);
        float4* lo = reinterpret_cast<float4*>(gen + s * STAGE_BYTES + A_BYTES);
#pragma unroll
        for (int j = 0; j < A_BYTES / 16 / 128; ++j) {
          const int i = t + 128 * j;
          const float4 v = hi[i];
          float4 h, l;
          h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
          l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
          hi[i] = h;
          lo[i] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(map_to_rank(bar_cv(s), 0));
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    const int64_t lrow = static_cast<int64_t>(m0) + q * 32 + lane;
    const bool row_ok = lrow < p.rows_per_batch && m_tile < p.m_tiles;
    const int64_t row = static_cast<int64_t>(grp) * p.rows_per_batch + lrow;
    float* yrow = p.y + (row_ok ? row * p.ldy : 0) + n0;
    const float* rrow = p.residual ? p.residual + (row_ok ? row * p.ldy : 0) + n0 : nullptr;
    const bool vec = ((p.ldy & 3) == 0) && aligned16(p.y) && (p.residual == nullptr || aligned16(p.residual));
    for (int c0 = 0; c0 < n_inst; c0 += 32) {
      uint32_t r[32], rc[32];
      const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c0);
      tmem_ld32(taddr, r);
      if (p.passes == 3) {
        tmem_ld32(taddr + BN, rc);
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
        for (int j = 0; j < 4; ++j)
          v[j] = (__uint_as_float(r[g * 4 + j]) + __uint_as_float(rc[g * 4 + j])) + bias_s[c + j];
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
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(TMEM_COLS) : "memory");
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
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

static bool make_map_w(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t ld) {
  auto enc = encode_fn();
  if (enc == nullptr) return false;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {BK, BH_ROWS};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool make_map_x(CUtensorMap* map, const float* ptr, int64_t groups, int64_t rows, int64_t cols, int64_t ld,
                       int64_t group_stride) {
  auto enc = encode_fn();
  if (enc == nullptr) return false;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(groups)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 4, static_cast<cuuint64_t>(group_stride) * 4};
  cuuint32_t box[3] = {BK, BM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc2

int launch_linear_tf32x3_2sm(const hoisdf_linear_args* a, cudaStream_t s) {
  using namespace tc2;
  CUtensorMap mx, mhi, mlo;
  int64_t groups = 1, rpb = a->m, gstride = a->m * a->ldx;
  if (a->x_rows_per_batch > 0) {
    if (a->m % a->x_rows_per_batch != 0) return HOISDF_E_SHAPE;
    rpb = a->x_rows_per_batch;
    groups = a->m / rpb;
    gstride = a->x_batch_stride;
  }
  if (groups > 1 && ((gstride * 4) % 16 != 0)) return HOISDF_E_ALIGN;
  if (groups == 1) gstride = rpb * a->ldx;
  if (!make_map_x(&mx, a->x, groups, rpb, a->k, a->ldx, gstride)) return HOISDF_E_UNSUPPORTED;
  if (!make_map_w(&mhi, a->w, a->n, a->k, a->ldw)) return HOISDF_E_UNSUPPORTED;
  if (!make_map_w(&mlo, a->w_lo, a->n, a->k, a->ldw)) return HOISDF_E_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(linear_tf32x3_2sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return static_cast<int>(e);
  const int64_t tpb = ceil_div(rpb, BM);
  const int64_t m_tiles = groups * tpb;
  Params p{a->bias, a->residual, a->y, a->ldy, rpb, static_cast<int>(tpb), static_cast<int>(m_tiles),
           static_cast<int>(a->n), static_cast<int>(a->k), a->act, a->tf32_passes == 1 ? 1 : 3};
  const int64_t ctas = ceil_div(m_tiles, 2) * 2 * ceil_div(a->n, BN);
  if (ctas > 0x7fffffffLL || m_tiles > 0x3fffffffLL) return HOISDF_E_SHAPE;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(ctas));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, linear_tf32x3_2sm_kernel, mx, mhi, mlo, p);
  if (e != cudaSuccess) return static_cast<int>(e);
  return launch_status();
}

}  // namespace hoisdf
