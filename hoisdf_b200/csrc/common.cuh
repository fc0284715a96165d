// Shared helpers for the hoisdf_b200 kernels (sm_100a only).
#pragma once
#ifdef HOISDF_EMULATE
// tests/emu: the CPU thread emulator that runs the simple (non-tensor-core) kernels in the "not gpu" test suite
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hoisdf_b200.h"

#define HOISDF_API extern "C" __attribute__((visibility("default")))
// kernel launch on `stream` with no dynamic shared memory; tests/emu/cuda_emu.h redefines it to run the kernel on CPU threads
#define HOISDF_LAUNCH(kernel, grid, block, stream, ...) kernel<<<grid, block, 0, stream>>>(__VA_ARGS__)
#define HOISDF_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define HOISDF_DYNAMIC_SMEM(type, name) extern __shared__ __align__(16) type name[]
#endif

namespace hoisdf {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

static inline int launch_status() {
#ifdef HOISDF_EMULATE
  return HOISDF_OK;
#else
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HOISDF_OK : static_cast<int>(e);
#endif
}

__host__ __device__ static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// The sheared lattice of upstream main/model.py:260-273, op for op (each step separately rounded).
//   col2 = idx % n ; col1 = (idx / n) mod n (TRUE division) ; col0 = ((idx / n) / n) mod n
//   s = col * (2/(n-1)) + (-1)
__device__ __forceinline__ void lattice_point(int idx, int bins, float& s0, float& s1, float& s2) {
  const float fn = static_cast<float>(bins);
  const float vs = static_cast<float>(2.0 / static_cast<double>(bins - 1));
  const float fi = static_cast<float>(idx);  // idx < 2^24: exact
  const float c2 = static_cast<float>(idx % bins);
  const float d1 = __fdiv_rn(fi, fn);
  const float c1 = fmodf(d1, fn);
  const float c0 = fmodf(__fdiv_rn(d1, fn), fn);
  s0 = __fadd_rn(__fmul_rn(c0, vs), -1.0f);
  s1 = __fadd_rn(__fmul_rn(c1, vs), -1.0f);
  s2 = __fadd_rn(__fmul_rn(c2, vs), -1.0f);
}

// ---- counter-based dropout decisions (nn.MultiheadAttention's dropout on the attention probabilities, upstream cfg.dropout):
// keep(row, col) is a pure function of (seed, row, col), so the backward regenerates the forward's decisions from the seed
// alone.  Two levels: a 32-bit key per probability ROW (splitmix64 finaliser of (seed, row), computed once per row) and a
// 32-bit avalanche hash of (key, col) per element -- ~10 integer instructions, cheap enough for the softmax warps of the
// tensor-core attention kernel.  An element is dropped when its top 24 hash bits fall below p_drop * 2^24.
__host__ __device__ __forceinline__ uint32_t dropout_row_key(uint64_t seed, uint64_t row) {
  uint64_t z = seed + (row + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<uint32_t>(z ^ (z >> 32));
}
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p_drop) {       // p_drop in [0, 1)
  return static_cast<uint32_t>(p_drop * 16777216.0f + 0.5f);
}
__host__ __device__ __forceinline__ bool dropout_keep(uint32_t row_key, uint32_t col, uint32_t threshold) {
  uint32_t x = row_key + col * 0x9E3779B1u;
  x ^= x >> 16; x *= 0x7FEB352Du;
  x ^= x >> 15; x *= 0x846CA68Bu;
  x ^= x >> 16;
  return (x >> 8) >= threshold;
}

}  // namespace hoisdf
