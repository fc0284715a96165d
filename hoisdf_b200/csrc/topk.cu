// Near-surface point selection (upstream main/model.py:345-354): per sample the `num_points` candidates with
// the smallest |sdf|, in ascending order, then the gathers of lattice coordinates / posenc / clamped SDF.
//
// Integer work, deterministic: the sort key is the 64-bit composite (bits(|sdf|) << 32 | local row), which is
// unique, so ties resolve to the lower lattice index and the result is independent of thread scheduling
// (torch.sort upstream is not stable; tie order there is unspecified).  One CTA per sample:
//   8-pass byte-wise radix SELECT of the P-th smallest composite (shared 256-bin histogram), collection of
//   the exactly P composites <= threshold, bitonic sort of those P in shared memory.
#include "common.cuh"

namespace hoisdf {

constexpr int kMaxSel = 8192;
constexpr int kSelThreads = 1024;

__device__ __forceinline__ uint64_t composite_key(float sdf, uint32_t i) {
  return (static_cast<uint64_t>(__float_as_uint(fabsf(sdf))) << 32) | i;
}

__global__ void __launch_bounds__(kSelThreads) select_points_kernel(
    const float* __restrict__ sdf, const int64_t* __restrict__ offsets, const int32_t* __restrict__ cand_index,
    int num_points, int bins, float clamp, int order_by_row, int32_t* __restrict__ sel_index,
    int32_t* __restrict__ sel_row, float* __restrict__ points, float* __restrict__ out_sdf,
    float* __restrict__ posenc, int32_t* __restrict__ status_flag) {
  HOISDF_DYNAMIC_SMEM(uint64_t, keys);               // npad entries (next power of two >= num_points)
  __shared__ unsigned hist[256];
  __shared__ uint64_t s_prefix;
  __shared__ unsigned s_need;
  __shared__ unsigned s_count;

  const int64_t b = blockIdx.x;
  const int64_t base = offsets[b];
  const int64_t n = offsets[b + 1] - base;
  const int tid = threadIdx.x;
  const float* sd = sdf + base;

  if (n < num_points) {  // upstream raises here (model.py:348 shape mismatch); the host shim turns the flag into that error
    if (tid == 0 && status_flag != nullptr) atomicExch(status_flag, 1);
    for (int j = tid; j < num_points; j += kSelThreads) {
      sel_index[b * num_points + j] = -1;
      if (sel_row != nullptr) sel_row[b * num_points + j] = -1;
      out_sdf[b * num_points + j] = 0.f;
      for (int c = 0; c < 3; ++c) points[(b * num_points + j) * 3 + c] = 0.f;
      for (int c = 0; c < 30; ++c) posenc[(b * num_points + j) * 30 + c] = 0.f;
    }
    return;
  }

  if (tid == 0) { s_prefix = 0; s_need = static_cast<unsigned>(num_points); s_count = 0; }
  uint64_t mask = 0;
  for (int pass = 7; pass >= 0; --pass) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const uint64_t prefix = s_prefix;
    const int shift = pass * 8;
    for (int64_t i = tid; i < n; i += kSelThreads) {
      const uint64_t k = composite_key(sd[i], static_cast<uint32_t>(i));
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid < 32) {
      // 256 bins over 32 lanes: each lane owns 8 consecutive bins
      unsigned local[8], tot = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { local[j] = hist[tid * 8 + j]; tot += local[j]; }
      unsigned incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += t;
      }
      const unsigned excl = incl - tot;
      const unsigned need = s_need;
      __syncwarp();
      if (need > excl && need <= incl) {  // the target bin lives in this lane's 8 bins (exactly one lane)
        unsigned run = excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (need > run && need <= run + local[j]) {
            s_prefix = prefix | (static_cast<uint64_t>(tid * 8 + j) << shift);
            s_need = need - run;
          }
          run += local[j];
        }
      }
    }
    mask |= 0xffull << shift;
    __syncthreads();
  }
  const uint64_t thresh = s_prefix;  // the P-th smallest composite (unique)

  // padded length for the bitonic network
  int npad = 1;
  while (npad < num_points) npad <<= 1;
  for (int j = tid; j < npad; j += kSelThreads) keys[j] = ~0ull;
  __syncthreads();
  for (int64_t i = tid; i < n; i += kSelThreads) {
    const uint64_t k = composite_key(sd[i], static_cast<uint32_t>(i));
    if (k <= thresh) keys[atomicAdd(&s_count, 1u)] = k;
  }
  __syncthreads();
  if (order_by_row) {
    // screening mode: emit the selected SET in ascending row (= lattice) order instead of |sdf| order, so that a
    // later exact re-ranking of this subset breaks ties exactly like a full-precision pass over all candidates
    for (int j = tid; j < num_points; j += kSelThreads) keys[j] = ((keys[j] & 0xffffffffull) << 32) | (keys[j] >> 32);
    __syncthreads();
  }

  for (int size = 2; size <= npad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (npad >> 1); t += kSelThreads) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const uint64_t a = keys[lo], c = keys[hi];
        if ((a > c) == up) { keys[lo] = c; keys[hi] = a; }
      }
      __syncthreads();
    }
  }

  const int rshift = order_by_row ? 32 : 0;
  for (int j = tid; j < num_points; j += kSelThreads) {
    const int64_t row = base + static_cast<int64_t>((keys[j] >> rshift) & 0xffffffffull);
    const int idx = cand_index[row];
    sel_index[b * num_points + j] = idx;
    if (sel_row != nullptr) sel_row[b * num_points + j] = static_cast<int32_t>(row);
    float s0, s1, s2;
    lattice_point(idx, bins, s0, s1, s2);
    float* pt = points + (b * num_points + j) * 3;
    pt[0] = s0; pt[1] = s1; pt[2] = s2;
    const float raw_sdf = sd[row - base];
    out_sdf[b * num_points + j] = clamp > 0.f ? fminf(fmaxf(raw_sdf, -clamp), clamp) : raw_sdf;
  }
  // NeRF embedding of the selected points: same sinf/cosf as posenc_kernel, so the values are bit-identical
  // to the ones the SDF decoder consumed (upstream copies them, model.py:350)
  for (int e = tid; e < num_points * 30; e += kSelThreads) {
    const int j = e / 30, c = e - j * 30;
    const int64_t row = base + static_cast<int64_t>((keys[j] >> rshift) & 0xffffffffull);
    float xyz[3];
    lattice_point(cand_index[row], bins, xyz[0], xyz[1], xyz[2]);
    const int oct = c / 6, w = c % 6;
    const float a = xyz[w % 3] * static_cast<float>(1 << oct);
    posenc[(b * num_points + j) * 30 + c] = (w < 3) ? sinf(a) : cosf(a);
  }
}

}  // namespace hoisdf

using namespace hoisdf;

HOISDF_API int hoisdf_select_points(const float* sdf, const int64_t* offsets, const int32_t* cand_index,
                                    int64_t batch, int64_t num_points, int32_t bins,
                                    float clamp, int32_t order_by_row, int32_t* sel_index, int32_t* sel_row,
                                    float* points, float* out_sdf, float* posenc, int32_t* status_flag,
                                    void* stream) {
  if (sdf == nullptr || offsets == nullptr || cand_index == nullptr || sel_index == nullptr ||
      points == nullptr || out_sdf == nullptr || posenc == nullptr)
    return HOISDF_E_NULL;
  if (batch <= 0 || batch > 65535 || num_points <= 0 || num_points > kMaxSel) return HOISDF_E_SHAPE;
  int npad = 1;
  while (npad < num_points) npad <<= 1;
  const size_t smem = static_cast<size_t>(npad) * sizeof(uint64_t);
#ifndef HOISDF_EMULATE
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(select_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
#endif
  HOISDF_LAUNCH_SMEM(select_points_kernel, static_cast<unsigned>(batch), kSelThreads, smem,
                     static_cast<cudaStream_t>(stream), sdf, offsets, cand_index, static_cast<int>(num_points), bins,
                     clamp, order_by_row, sel_index, sel_row, points, out_sdf, posenc, status_flag);
  return launch_status();
}
