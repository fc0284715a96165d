// TEST INFRASTRUCTURE: a minimal CPU emulation of the CUDA execution model, enough to run the simple (non-tensor-core)
// kernels of hoisdf_b200/csrc/*.cu unchanged on a host without a GPU.  Every CUDA thread of a block is a FIBER (ucontext)
// on one OS thread; blocks run one after another; `__shared__` = function-local static storage; `__syncthreads()` and
// the warp collectives (shuffles, ballots, __syncwarp) are cooperative barriers: a fiber that arrives yields to the next
// one until the barrier's generation advances.  As on the GPU, threads that have left the kernel are not waited for.
// Deterministic (round-robin schedule) and free of kernel-level thread switches.  Built by
// tests/test_kernel_emulation.py with `g++ -std=c++20 -DHOISDF_EMULATE`; never part of libhoisdf_b200.so.
#pragma once
#include <stdint.h>

#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <vector>

#include "../../include/hoisdf_b200.h"

#define HOISDF_API extern "C" __attribute__((visibility("default")))
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
using cudaStream_t = void*;
using cudaError_t = int;
constexpr cudaError_t cudaSuccess = 0;
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }

namespace emu {
inline dim3 thread_idx, block_idx;          // of the running fiber (one OS thread: plain globals)
inline dim3 block_dim, grid_dim;
inline unsigned char exchange[1024][16];
alignas(16) inline unsigned char dynamic_smem[228 * 1024];      // `extern __shared__` storage of the running block

struct Barrier {
  unsigned alive = 0, count = 0, generation = 0;
  void release_if_complete() {
    if (alive > 0 && count >= alive) {
      count = 0;
      ++generation;
    }
  }
};

struct Fiber {
  ucontext_t ctx;
  std::unique_ptr<unsigned char[]> stack;
  dim3 tid;
  bool done = false;
};

constexpr size_t kStackBytes = 256 * 1024;
inline ucontext_t scheduler_ctx;
inline std::vector<Fiber> fibers;
inline unsigned current = 0;
inline Barrier block_barrier;
inline std::vector<Barrier> warp_barrier;
inline std::function<void()> block_body;

inline void yield() { swapcontext(&fibers[current].ctx, &scheduler_ctx); }

inline void wait(Barrier& b) {
  const unsigned gen = b.generation;
  ++b.count;
  b.release_if_complete();
  while (b.generation == gen) yield();
}

inline void fiber_entry() {
  block_body();
  Fiber& f = fibers[current];
  f.done = true;                              // a thread that has left the kernel is no longer waited for
  --block_barrier.alive;
  block_barrier.release_if_complete();
  Barrier& w = warp_barrier[current >> 5];
  --w.alive;
  w.release_if_complete();
  swapcontext(&f.ctx, &scheduler_ctx);        // never resumed
}

template <typename K, typename... Args>
void launch(K kernel, dim3 grid, dim3 block, Args... args) {
  grid_dim = grid;
  block_dim = block;
  const unsigned n = block.x * block.y * block.z;
  if (fibers.size() < n) {
    const size_t old = fibers.size();
    fibers.resize(n);
    for (size_t i = old; i < n; ++i) fibers[i].stack.reset(new unsigned char[kStackBytes]);
  }
  block_body = [&]() { kernel(args...); };
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        block_idx = dim3(bx, by, bz);
        block_barrier = Barrier{n, 0, 0};
        warp_barrier.assign((n + 31) / 32, Barrier{});
        for (unsigned w = 0; w * 32 < n; ++w) warp_barrier[w].alive = std::min(32u, n - w * 32);
        for (unsigned t = 0; t < n; ++t) {
          Fiber& f = fibers[t];
          f.done = false;
          f.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack.get();
          f.ctx.uc_stack.ss_size = kStackBytes;
          f.ctx.uc_link = nullptr;
          makecontext(&f.ctx, fiber_entry, 0);
        }
        unsigned remaining = n;
        while (remaining > 0) {               // round-robin until every thread of the block has left the kernel
          remaining = 0;
          for (unsigned t = 0; t < n; ++t) {
            if (fibers[t].done) continue;
            current = t;
            thread_idx = fibers[t].tid;
            swapcontext(&scheduler_ctx, &fibers[t].ctx);
            if (!fibers[t].done) ++remaining;
          }
        }
      }
}
}  // namespace emu

#define threadIdx emu::thread_idx
#define blockIdx emu::block_idx
#define blockDim emu::block_dim
#define gridDim emu::grid_dim
#define HOISDF_LAUNCH(kernel, grid, block, stream, ...) emu::launch(kernel, dim3(grid), dim3(block), __VA_ARGS__)
#define HOISDF_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) emu::launch(kernel, dim3(grid), dim3(block), __VA_ARGS__)
#define HOISDF_DYNAMIC_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::dynamic_smem)

inline void __syncthreads() { emu::wait(emu::block_barrier); }
inline unsigned emu_tid() { return emu::current; }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::wait(emu::warp_barrier[emu::current >> 5]); }

template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  static_assert(sizeof(T) <= 16, "exchange slot too small");
  const unsigned tid = emu_tid();
  std::memcpy(emu::exchange[tid], &v, sizeof(T));
  __syncwarp();
  T r;
  std::memcpy(&r, emu::exchange[(tid & ~31u) | ((tid ^ static_cast<unsigned>(lane_mask)) & 31u)], sizeof(T));
  __syncwarp();
  return r;
}

template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
  const unsigned tid = emu_tid();
  std::memcpy(emu::exchange[tid], &v, sizeof(T));
  __syncwarp();
  T r = v;                                                   // lanes below `delta` keep their own value
  if ((tid & 31u) >= delta) std::memcpy(&r, emu::exchange[tid - delta], sizeof(T));
  __syncwarp();
  return r;
}

inline unsigned __ballot_sync(unsigned, bool pred) {
  const unsigned tid = emu_tid();
  const unsigned n = blockDim.x * blockDim.y * blockDim.z;
  emu::exchange[tid][0] = pred ? 1 : 0;
  __syncwarp();
  unsigned m = 0;
  for (unsigned l = 0; l < 32; ++l) {
    const unsigned t = (tid & ~31u) | l;
    if (t < n && emu::exchange[t][0]) m |= 1u << l;
  }
  __syncwarp();
  return m;
}

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline float atomicAdd(float* p, float v) { const float old = *p; *p = old + v; return old; }   // fibers on ONE OS thread
template <typename T> inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }

// separately rounded IEEE single-precision operations (x86-64 SSE arithmetic is IEEE; `volatile` keeps the compiler from
// contracting or reassociating them)
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
inline float __double2float_rn(double a) { volatile float r = static_cast<float>(a); return r; }

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }

// software fp16 (IEEE binary16 through the compiler's _Float16: conversions round to nearest even like cvt.rn)
struct __half { _Float16 v; };
struct __half2 { __half x, y; };
inline __half __float2half_rn(float f) { return __half{static_cast<_Float16>(f)}; }
inline float __half2float(__half h) { return static_cast<float>(h.v); }
inline unsigned short __half_as_ushort(__half h) { unsigned short u; std::memcpy(&u, &h.v, 2); return u; }
inline __half __ushort_as_half(unsigned short u) { __half h; std::memcpy(&h.v, &u, 2); return h; }
inline float2 __half22float2(__half2 h) { return float2{__half2float(h.x), __half2float(h.y)}; }

using std::isnan;
using std::max;
using std::min;
