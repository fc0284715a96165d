// sdf_infer candidate generation: sheared 64^3 lattice -> camera -> pixels -> strict bbox test -> stable
// compaction (upstream main/model.py:257-302), plus the pinhole projection of explicit points
// (main/model.py:148-150, 190-192).
//
// Integer/bit-exact work: every float op is a separately rounded IEEE op in the order the upstream CPU code
// executes it (SURVEY.md section 7 "Bit-exact index masks"):
//   s    = col * fp32(2/63) + (-1)          (mul, then add; no FMA contraction)
//   cam  = s / fp32(3.1) + center           (IEEE division, then add)
//   uvw  = fma(z, K[:,2], fma(y, K[:,1], x * K[:,0]))   (what the MKL sgemm of model.py:290 evaluates)
//   uv   = uvw[:2] / uvw[2]                 (IEEE division)
// One warp owns 1024 consecutive lattice indices ("chunk"); ballot/popc keeps the compaction stable.
#include "common.cuh"

namespace hoisdf {

constexpr int kChunk = 1024;

struct Camera {
  float cx, cy, cz;
  float k[9];
  float b0, b1, b2, b3;
};

__device__ __forceinline__ Camera load_camera(const float* center, const float* K, const float* bbox, int64_t b) {
  Camera c;
  c.cx = center[b * 3 + 0]; c.cy = center[b * 3 + 1]; c.cz = center[b * 3 + 2];
#pragma unroll
  for (int i = 0; i < 9; ++i) c.k[i] = K[b * 9 + i];
  if (bbox != nullptr) {
    c.b0 = bbox[b * 4 + 0]; c.b1 = bbox[b * 4 + 1]; c.b2 = bbox[b * 4 + 2]; c.b3 = bbox[b * 4 + 3];
  } else {
    c.b0 = c.b1 = c.b2 = c.b3 = 0.f;
  }
  return c;
}

__device__ __forceinline__ void project(const Camera& c, float x, float y, float z, float& u, float& v) {
  const float w0 = __fmaf_rn(z, c.k[2], __fmaf_rn(y, c.k[1], __fmul_rn(x, c.k[0])));
  const float w1 = __fmaf_rn(z, c.k[5], __fmaf_rn(y, c.k[4], __fmul_rn(x, c.k[3])));
  const float w2 = __fmaf_rn(z, c.k[8], __fmaf_rn(y, c.k[7], __fmul_rn(x, c.k[6])));
  u = __fdiv_rn(w0, w2);
  v = __fdiv_rn(w1, w2);
}

__device__ __forceinline__ bool lattice_candidate(const Camera& c, int idx, int bins, float scale, float& u,
                                                  float& v) {
  float s0, s1, s2;
  lattice_point(idx, bins, s0, s1, s2);
  const float x = __fadd_rn(__fdiv_rn(s0, scale), c.cx);
  const float y = __fadd_rn(__fdiv_rn(s1, scale), c.cy);
  const float z = __fadd_rn(__fdiv_rn(s2, scale), c.cz);
  project(c, x, y, z, u, v);
  return (u > c.b0) && (u < c.b2) && (v > c.b1) && (v < c.b3);
}

// grid (chunks, B), 32 threads
__global__ void __launch_bounds__(32) lattice_count_kernel(const float* __restrict__ center,
                                                           const float* __restrict__ K,
                                                           const float* __restrict__ bbox, float scale, int bins,
                                                           int total, int chunks, int32_t* __restrict__ counts) {
  const int64_t b = blockIdx.y;
  const Camera c = load_camera(center, K, bbox, b);
  const int base = blockIdx.x * kChunk;
  int n = 0;
  for (int it = 0; it < kChunk / 32; ++it) {
    const int idx = base + it * 32 + threadIdx.x;
    float u, v;
    const bool keep = idx < total && lattice_candidate(c, idx, bins, scale, u, v);
    n += __popc(__ballot_sync(0xffffffffu, keep));
  }
  if (threadIdx.x == 0) counts[b * chunks + blockIdx.x] = n;
}

// one CTA: per-sample exclusive scan of the chunk counts (in place -> global row offsets), sample offsets
__global__ void __launch_bounds__(1024) lattice_scan_kernel(int32_t* __restrict__ counts, int64_t batch, int chunks,
                                                            int64_t* __restrict__ offsets) {
  __shared__ int warp_tot[32];
  __shared__ int64_t running;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t b = 0; b < batch; ++b) {
    int carry = 0;  // chunks may exceed blockDim: iterate in slabs of 1024
    for (int c0 = 0; c0 < chunks; c0 += 1024) {
      const int i = c0 + threadIdx.x;
      const int v = i < chunks ? counts[b * chunks + i] : 0;
      int s = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
      }
      if (lane == 31) warp_tot[wid] = s;
      __syncthreads();
      if (wid == 0) {
        int t = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, t, o);
          if (lane >= o) t += u;
        }
        warp_tot[lane] = t;  // inclusive
      }
      __syncthreads();
      const int before = (wid > 0 ? warp_tot[wid - 1] : 0) + carry;
      const int64_t base = running;
      if (i < chunks) counts[b * chunks + i] = static_cast<int32_t>(base + before + s - v);
      carry += warp_tot[31];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      offsets[b] = running;
      running += carry;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[batch] = running;
}

// grid (chunks, B), 32 threads; chunk_offsets = exclusive global row offsets from the scan
__global__ void __launch_bounds__(32) lattice_compact_kernel(const float* __restrict__ center,
                                                             const float* __restrict__ K,
                                                             const float* __restrict__ bbox, float scale, int bins,
                                                             int total, int chunks,
                                                             const int32_t* __restrict__ chunk_offsets,
                                                             int32_t* __restrict__ cand_index,
                                                             float* __restrict__ cand_uv) {
  const int64_t b = blockIdx.y;
  const Camera c = load_camera(center, K, bbox, b);
  const int base = blockIdx.x * kChunk;
  int64_t pos = chunk_offsets[b * chunks + blockIdx.x];
  for (int it = 0; it < kChunk / 32; ++it) {
    const int idx = base + it * 32 + threadIdx.x;
    float u = 0.f, v = 0.f;
    const bool keep = idx < total && lattice_candidate(c, idx, bins, scale, u, v);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int64_t o = pos + __popc(m & ((1u << threadIdx.x) - 1u));
      cand_index[o] = idx;
      reinterpret_cast<float2*>(cand_uv)[o] = make_float2(u, v);
    }
    pos += __popc(m);
  }
}

__global__ void project_points_kernel(const float* __restrict__ points, const float* __restrict__ center,
                                      const float* __restrict__ K, float scale, int64_t batch, int64_t p,
                                      float* __restrict__ cam, float* __restrict__ uv) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= batch * p) return;
  const int64_t b = i / p;
  const Camera c = load_camera(center, K, nullptr, b);
  const float x = __fadd_rn(__fdiv_rn(points[i * 3 + 0], scale), c.cx);
  const float y = __fadd_rn(__fdiv_rn(points[i * 3 + 1], scale), c.cy);
  const float z = __fadd_rn(__fdiv_rn(points[i * 3 + 2], scale), c.cz);
  float u, v;
  project(c, x, y, z, u, v);
  if (cam != nullptr) {
    cam[i * 3 + 0] = x; cam[i * 3 + 1] = y; cam[i * 3 + 2] = z;
  }
  uv[i * 2 + 0] = u;
  uv[i * 2 + 1] = v;
}

}  // namespace hoisdf

using namespace hoisdf;

static int lattice_check(const float* center, const float* K, const float* bbox, int64_t batch, int32_t bins) {
  if (center == nullptr || K == nullptr || bbox == nullptr) return HOISDF_E_NULL;
  if (batch <= 0 || batch > 4096 || bins < 2 || bins > 256) return HOISDF_E_SHAPE;
  return HOISDF_OK;
}

HOISDF_API int hoisdf_lattice_chunks(int32_t bins) {
  const int64_t total = static_cast<int64_t>(bins) * bins * bins;
  return static_cast<int>(ceil_div(total, kChunk));
}

HOISDF_API int hoisdf_lattice_count(const float* center, const float* cam_intr, const float* bbox, float sdf_scale,
                                    int64_t batch, int32_t bins, int32_t* chunk_counts, int64_t* offsets,
                                    void* stream) {
  int st = lattice_check(center, cam_intr, bbox, batch, bins);
  if (st != HOISDF_OK) return st;
  if (chunk_counts == nullptr || offsets == nullptr) return HOISDF_E_NULL;
  const int total = bins * bins * bins;
  const int chunks = hoisdf_lattice_chunks(bins);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  HOISDF_LAUNCH(lattice_count_kernel, dim3(chunks, static_cast<unsigned>(batch)), 32, s, center, cam_intr, bbox,
                sdf_scale, bins, total, chunks, chunk_counts);
  HOISDF_LAUNCH(lattice_scan_kernel, 1, 1024, s, chunk_counts, batch, chunks, offsets);
  return launch_status();
}

HOISDF_API int hoisdf_lattice_compact(const float* center, const float* cam_intr, const float* bbox,
                                      float sdf_scale, int64_t batch, int32_t bins, const int32_t* chunk_counts,
                                      const int64_t* offsets, int32_t* cand_index, float* cand_uv, void* stream) {
  int st = lattice_check(center, cam_intr, bbox, batch, bins);
  if (st != HOISDF_OK) return st;
  if (chunk_counts == nullptr || offsets == nullptr || cand_index == nullptr || cand_uv == nullptr)
    return HOISDF_E_NULL;
  if (reinterpret_cast<uintptr_t>(cand_uv) & 7u) return HOISDF_E_ALIGN;
  const int total = bins * bins * bins;
  const int chunks = hoisdf_lattice_chunks(bins);
  HOISDF_LAUNCH(lattice_compact_kernel, dim3(chunks, static_cast<unsigned>(batch)), 32, static_cast<cudaStream_t>(stream),
                center, cam_intr, bbox, sdf_scale, bins, total, chunks, chunk_counts, cand_index, cand_uv);
  return launch_status();
}

HOISDF_API int hoisdf_project_points(const float* points, const float* center, const float* cam_intr,
                                     float sdf_scale, int64_t batch, int64_t p, float* cam, float* uv,
                                     void* stream) {
  if (points == nullptr || center == nullptr || cam_intr == nullptr || uv == nullptr) return HOISDF_E_NULL;
  if (batch <= 0 || p <= 0) return HOISDF_E_SHAPE;
  const int64_t n = batch * p;
  HOISDF_LAUNCH(project_points_kernel, static_cast<unsigned>(ceil_div(n, 256)), 256, static_cast<cudaStream_t>(stream),
                points, center, cam_intr, sdf_scale, batch, p, cam, uv);
  return launch_status();
}
