// Microbenchmark (developer tool): issue rate of tcgen05.mma kind::f16, M = 128, K = 16, for N = 32 ... 256 with the
// A operand in shared memory (SS) or in tensor memory (TS).  One CTA per SM, one issuing thread, R back-to-back MMAs
// into one accumulator (or alternating between two), timed with clock64 around issue -> commit -> mbarrier wait.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/_bin/umma_rate scripts/umma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "../hoisdf_b200/csrc/tc_common.cuh"

using namespace hoisdf::tc;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

template <bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int reps, int two_acc, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  for (uint32_t i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x)          // finite operands (fp16 1.0)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + 4 * i), "r"(0x3c003c00u));
  if (threadIdx.x == 0) mbar_init(smem_u32(&bar), 1);
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, n);
    const uint64_t da = umma_desc_sw128(base), db = umma_desc_sw128(base + 16384);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tmem + ((two_acc && (r & 1)) ? 256u : 0u);
      if (TS) umma_f16_ts(d, tmem + 480u, db, idesc, r > 1 ? 1u : 0u);
      else umma_f16(d, da, db, idesc, r > 1 ? 1u : 0u);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 148 * sizeof(long long));
  const int reps = 4096;
  cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int ts = 0; ts < 2; ++ts)
    for (int two = 0; two < 2; ++two)
      for (int n : {32, 64, 128, 256}) {
        if (two && n == 256 && ts) continue;      // second accumulator would overlap the TMEM A operand
        for (int grid : {1, 148}) {
          for (int w = 0; w < 2; ++w) {
            if (ts) rate_kernel<true><<<grid, 128, 64 * 1024>>>(n, reps, two, out);
            else rate_kernel<false><<<grid, 128, 64 * 1024>>>(n, reps, two, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
              printf("error %s\n", cudaGetErrorString(e));
              return 1;
            }
          }
          long long mx = 0;
          for (int i = 0; i < grid; ++i) mx = out[i] > mx ? out[i] : mx;
          printf("%s N=%3d accumulators=%d grid=%3d : %6.1f cycles per MMA (M128 x N x K16) -> %5.0f flop/clk/SM\n",
                 ts ? "TS" : "SS", n, two + 1, grid, double(mx) / reps, 2.0 * 128 * n * 16 * reps / double(mx));
        }
      }
  return 0;
}
