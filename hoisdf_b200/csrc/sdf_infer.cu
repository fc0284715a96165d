// hoisdf_sdf_infer_fwd -- upstream Model.sdf_infer (main/model.py:246-355) for the WHOLE batch behind ONE C entry point:
// candidate generation (sheared 64^3 lattice -> camera -> pixels -> strict bbox test -> stable compaction), the verified
// coarse-to-fine selection cascade of DESIGN.md 4.2 and the final top-P by |sdf| -- the orchestration that
// hoisdf_b200/model.py:Model.sdf_infer used to do in Python with ATen glue (kthvalue / index_select / abs().max()).
//
//   stage A   every candidate row: fp16 gather of the projected maps (hoisdf_gather_sum_h16_fwd) -> ONE persistent tcgen05
//             kernel linear_sdfin.1 -> NeRF embedding -> SDFDecoder (hoisdf_sdf_chain_fwd), single-product arithmetic
//   screening keep the  P + margin  smallest |sdf| per sample, in lattice order (hoisdf_select_points, order_by_row)
//   final     those rows again on the FP16x3 kernels draining TMEM every K block (fp32-FMA-grade values):
//             hoisdf_gather_split_fwd -> hoisdf_linear_h3_fwd -> hoisdf_posenc_split_fwd -> hoisdf_sdf_decoder_h3_fwd
//   verdict   err = max |coarse - fine|,  gap_b = max_b |coarse| - (P-th smallest |fine|)_b,  verified = all(gap > 3 err)
//             -- one kernel, results stay on the device (the caller reads `verified` when it pleases)
//   top-P     hoisdf_select_points on the final values: points, sdf (clamped), NeRF embedding, lattice indices
//
// The caller owns every buffer (workspace sized by hoisdf_sdf_infer_workspace_bytes).  The only host interaction is the read
// of the B + 1 row offsets that size the launches: `host_offsets` is caller-provided pinned memory; either the caller ran
// hoisdf_sdf_infer_plan earlier (the copy has long landed: no stall) or this function plans and waits itself.
#include <cstring>

#include "common.cuh"

namespace hoisdf {
namespace {

// cand_index / cand_uv rows picked by `rows` (the screening survivors, lattice order)
__global__ void take_rows_kernel(const int32_t* __restrict__ rows, int64_t n, const int32_t* __restrict__ src_index,
                                 const float2* __restrict__ src_uv, int32_t* __restrict__ dst_index,
                                 float2* __restrict__ dst_uv) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t r = rows[i];
  dst_index[i] = src_index[r];
  dst_uv[i] = src_uv[r];
}

__global__ void iota_offsets_kernel(int64_t* __restrict__ offsets, int64_t batch, int64_t keep) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i <= batch) offsets[i] = i * keep;
}

// One CTA per sample: coarse (screening) vs fine (final stage) values of its `keep` survivors.
//   err_part[b] = max |coarse - fine|;  gap[b] = max |coarse| - (target-th smallest |fine|)
// The target-th smallest |fine| by a bitonic sort of the |fine| bit patterns in shared memory (keep <= 8192).
__global__ void __launch_bounds__(1024)
screen_verify_kernel(const float* __restrict__ coarse, const float* __restrict__ fine, int keep, int target,
                     float* __restrict__ err_part, float* __restrict__ gap) {
  extern __shared__ float sh[];
  __shared__ float red[2][32];
  const int b = blockIdx.x, tid = threadIdx.x;
  int n2 = 1;
  while (n2 < keep) n2 <<= 1;
  float e = 0.f, cm = 0.f;
  for (int i = tid; i < n2; i += blockDim.x) {
    float f = 3.402823466e+38f;
    if (i < keep) {
      const float c = coarse[static_cast<int64_t>(b) * keep + i];
      f = fabsf(fine[static_cast<int64_t>(b) * keep + i]);
      e = fmaxf(e, fabsf(c - fine[static_cast<int64_t>(b) * keep + i]));
      cm = fmaxf(cm, fabsf(c));
    }
    sh[i] = f;
  }
  e = warp_max(e);
  cm = warp_max(cm);
  if ((tid & 31) == 0) { red[0][tid >> 5] = e; red[1][tid >> 5] = cm; }
  __syncthreads();
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float a = sh[i], c = sh[ixj];
          const bool up = (i & k) == 0;
          if ((a > c) == up) { sh[i] = c; sh[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  if (tid == 0) {
    float em = 0.f, cmm = 0.f;
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; ++w) { em = fmaxf(em, red[0][w]); cmm = fmaxf(cmm, red[1][w]); }
    err_part[b] = em;
    gap[b] = cmm - sh[target - 1];
  }
}

__global__ void screen_verdict_kernel(const float* __restrict__ err_part, const float* __restrict__ gap, int batch,
                                      float* __restrict__ err, int32_t* __restrict__ verified) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float e = 0.f;
  for (int b = 0; b < batch; ++b) e = fmaxf(e, err_part[b]);
  int ok = 1;
  for (int b = 0; b < batch; ++b) ok &= (gap[b] > 3.0f * e) ? 1 : 0;
  *err = e;
  *verified = ok;
}

constexpr int64_t kAlign = 256;
inline int64_t up(int64_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

struct Layout {
  int64_t counts, offsets, cand_index, cand_uv, sdf_all, a0, s_sel, s_row, s_pts, s_sdf, s_pe, s_flag, n_index, n_uv, n_sdf,
      n_offsets, hs, rs, hs2, err_part, total;
};

Layout make_layout(int64_t batch, int64_t max_rows, int64_t pass_rows, int64_t keep, int32_t bins) {
  Layout L{};
  int64_t o = 0;
  auto take = [&](int64_t bytes) { const int64_t at = o; o += up(bytes); return at; };
  const int64_t chunks = hoisdf_lattice_chunks(bins);
  const int64_t bk = batch * keep;
  L.counts = take(batch * chunks * 4);
  L.offsets = take((batch + 1) * 8);
  L.cand_index = take(max_rows * 4);
  L.cand_uv = take(max_rows * 8);
  L.sdf_all = take(max_rows * 4);
  L.a0 = take(pass_rows * 512 * 2);              // stage A: fp16 hi plane of relu(linear_sdfin.0), one pass of rows
  L.s_sel = take(bk * 4);
  L.s_row = take(bk * 4);
  L.s_pts = take(bk * 3 * 4);
  L.s_sdf = take(bk * 4);
  L.s_pe = take(bk * 30 * 4);
  L.s_flag = take(4);
  L.n_index = take(bk * 4);
  L.n_uv = take(bk * 8);
  L.n_sdf = take(bk * 4);
  L.n_offsets = take((batch + 1) * 8);
  L.hs = take(bk * 2 * 512 * 2);                 // final stage: split-half rows (hi | lo per row)
  L.rs = take(bk * 2 * 520 * 2);
  L.hs2 = take(bk * 2 * 512 * 2);
  L.err_part = take(batch * 4);
  L.total = o;
  return L;
}

constexpr int64_t kPassRows = 1 << 23;           // most candidate rows one call takes (a0 buffer: 1 KB per row)

}  // namespace
}  // namespace hoisdf

using namespace hoisdf;

HOISDF_API int64_t hoisdf_sdf_infer_keep(int64_t num_points, int64_t margin) {
  const int64_t k = num_points + margin;
  return k < 8192 ? k : 8192;
}

HOISDF_API int64_t hoisdf_sdf_infer_workspace_bytes(int64_t batch, int64_t max_rows, int64_t num_points, int64_t margin,
                                                    int32_t bins) {
  if (batch <= 0 || max_rows <= 0 || num_points <= 0 || margin < 0) return 0;
  const int64_t pass = max_rows < kPassRows ? max_rows : kPassRows;
  return make_layout(batch, max_rows, pass, hoisdf_sdf_infer_keep(num_points, margin), bins).total;
}

// Pass 1 of the candidate generation + the asynchronous copy of the row offsets to the caller's pinned buffer.
HOISDF_API int hoisdf_sdf_infer_plan(const float* center, const float* cam_intr, const float* bbox, float sdf_scale,
                                     int64_t batch, int32_t bins, int32_t* chunk_counts, int64_t* offsets,
                                     int64_t* host_offsets, void* stream) {
  if (host_offsets == nullptr) return HOISDF_E_NULL;
  const int st = hoisdf_lattice_count(center, cam_intr, bbox, sdf_scale, batch, bins, chunk_counts, offsets, stream);
  if (st != HOISDF_OK) return st;
  const cudaError_t e = cudaMemcpyAsync(host_offsets, offsets, sizeof(int64_t) * (batch + 1), cudaMemcpyDeviceToHost,
                                        static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? HOISDF_OK : static_cast<int>(e);
}

HOISDF_API int hoisdf_sdf_infer_fwd(const hoisdf_sdf_infer_args* a, void* stream) {
  if (a == nullptr || a->center == nullptr || a->cam_intr == nullptr || a->bbox == nullptr || a->gmaps == nullptr ||
      a->gmaps16 == nullptr || a->bias0 == nullptr || a->s1_a == nullptr || a->s1_b == nullptr || a->s1_c == nullptr ||
      a->dec == nullptr || a->workspace == nullptr || a->host_offsets == nullptr || a->points == nullptr ||
      a->sdf == nullptr || a->posenc == nullptr || a->sel_index == nullptr || a->screen_err == nullptr ||
      a->screen_gap == nullptr || a->verified == nullptr || a->status_flag == nullptr)
    return HOISDF_E_NULL;
  if (a->batch <= 0 || a->batch > 65535 || a->num_points <= 0 || a->margin <= 0 || a->bins <= 0) return HOISDF_E_SHAPE;
  if (!aligned16(a->workspace)) return HOISDF_E_ALIGN;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t B = a->batch, P = a->num_points;
  const int64_t keep = hoisdf_sdf_infer_keep(P, a->margin);
  if (keep <= P) return HOISDF_E_UNSUPPORTED;
  // the workspace was sized for max_rows candidate rows: recover max_rows from its size (monotone in max_rows)
  int64_t max_rows = a->max_rows;
  if (max_rows <= 0) return HOISDF_E_SHAPE;
  const int64_t pass_rows = max_rows < kPassRows ? max_rows : kPassRows;
  const Layout L = make_layout(B, max_rows, pass_rows, keep, a->bins);
  if (L.total > a->workspace_bytes) return HOISDF_E_WORKSPACE;
  char* ws = static_cast<char*>(a->workspace);
  auto at = [&](int64_t off) { return static_cast<void*>(ws + off); };
  int32_t* counts = a->chunk_counts != nullptr ? a->chunk_counts : static_cast<int32_t*>(at(L.counts));
  int64_t* offsets = a->offsets != nullptr ? a->offsets : static_cast<int64_t*>(at(L.offsets));
  int st;
  if (!a->planned) {
    st = hoisdf_sdf_infer_plan(a->center, a->cam_intr, a->bbox, a->sdf_scale, B, a->bins, counts, offsets, a->host_offsets,
                               stream);
    if (st != HOISDF_OK) return st;
    const cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return static_cast<int>(e);
  } else if (a->chunk_counts == nullptr || a->offsets == nullptr) {
    return HOISDF_E_NULL;
  }
  // (planned: the caller guarantees the copy into host_offsets has completed -- it synchronised on its own event)
  const int64_t total = a->host_offsets[B];
  int64_t nmin = total;
  for (int64_t b = 0; b < B; ++b) {
    const int64_t nf = a->host_offsets[b + 1] - a->host_offsets[b];
    if (a->n_f != nullptr) a->n_f[b] = nf;
    nmin = nf < nmin ? nf : nmin;
  }
  if (nmin < P) return HOISDF_E_TOO_FEW_POINTS;       // upstream fails here too (model.py:348: shape mismatch)
  if (nmin <= keep) return HOISDF_E_UNSUPPORTED;      // no room for the screening margin: the caller ranks all rows exactly
  if (total > max_rows) return HOISDF_E_WORKSPACE;

  int32_t* cand_index = static_cast<int32_t*>(at(L.cand_index));
  float* cand_uv = static_cast<float*>(at(L.cand_uv));
  float* sdf_all = static_cast<float*>(at(L.sdf_all));
  st = hoisdf_lattice_compact(a->center, a->cam_intr, a->bbox, a->sdf_scale, B, a->bins, counts, offsets, cand_index,
                              cand_uv, stream);
  if (st != HOISDF_OK) return st;

  // ---- stage A: every candidate row, single-product fp16
  const hoisdf_sdf_weights_h3* d = a->dec;
  uint16_t* a0 = static_cast<uint16_t*>(at(L.a0));
  if (total > pass_rows) return HOISDF_E_UNSUPPORTED;     // beyond 8 M rows the caller splits the batch
  {
    st = hoisdf_gather_sum_h16_fwd(a->gmaps16, cand_uv, total, offsets, B, 0, a->bias0, HOISDF_ACT_RELU, a0, 512, stream);
    if (st != HOISDF_OK) return st;
    hoisdf_sdf_chain_args c;
    std::memset(&c, 0, sizeof(c));
    c.a0 = a0; c.lda0 = 512;
    c.lattice_index = cand_index; c.bins = a->bins;
    c.w_s1 = static_cast<const uint16_t*>(a->s1_b); c.ldw_s1 = a->ld_s1; c.b_s1 = a->b_s1;
    for (int l = 0; l < 4; ++l) { c.w[l] = d->w[l][1]; c.ldw[l] = d->ldw[l]; c.b[l] = d->b[l]; }
    c.w4 = d->w4; c.b4 = d->b4;
    c.rows = total; c.clamp = 0.f; c.out_sdf = sdf_all;
    st = hoisdf_sdf_chain_fwd(&c, stream);
    if (st != HOISDF_OK) return st;
  }

  // ---- screening: the `keep` best rows per sample, lattice order
  int32_t* s_row = static_cast<int32_t*>(at(L.s_row));
  float* s_sdf = static_cast<float*>(at(L.s_sdf));
  int32_t* s_flag = static_cast<int32_t*>(at(L.s_flag));
  if (cudaMemsetAsync(s_flag, 0, 4, s) != cudaSuccess) return HOISDF_E_SHAPE;
  st = hoisdf_select_points(sdf_all, offsets, cand_index, B, keep, a->bins, 0.f, 1, static_cast<int32_t*>(at(L.s_sel)), s_row,
                            static_cast<float*>(at(L.s_pts)), s_sdf, static_cast<float*>(at(L.s_pe)), s_flag, stream);
  if (st != HOISDF_OK) return st;
  const int64_t bk = B * keep;
  int32_t* n_index = static_cast<int32_t*>(at(L.n_index));
  float* n_uv = static_cast<float*>(at(L.n_uv));
  float* n_sdf = static_cast<float*>(at(L.n_sdf));
  int64_t* n_offsets = static_cast<int64_t*>(at(L.n_offsets));
  take_rows_kernel<<<static_cast<unsigned>(ceil_div(bk, 256)), 256, 0, s>>>(
      s_row, bk, cand_index, reinterpret_cast<const float2*>(cand_uv), n_index, reinterpret_cast<float2*>(n_uv));
  iota_offsets_kernel<<<static_cast<unsigned>(ceil_div(B + 1, 256)), 256, 0, s>>>(n_offsets, B, keep);

  // ---- final stage: FP16x3, TMEM drained every K block
  uint16_t* hs = static_cast<uint16_t*>(at(L.hs));
  uint16_t* rs = static_cast<uint16_t*>(at(L.rs));
  uint16_t* hs2 = static_cast<uint16_t*>(at(L.hs2));
  const int64_t ldh = 2 * 512, ldr = 2 * 520;                         // row pitch of the (rows, 2, ld) split-half buffers
  st = hoisdf_gather_split_fwd(a->gmaps, n_uv, bk, nullptr, B, keep, HOISDF_GATHER_SUM, a->bias0, HOISDF_ACT_RELU, hs, hs + 512,
                               ldh, stream);
  if (st != HOISDF_OK) return st;
  hoisdf_linear_h3_args l;
  std::memset(&l, 0, sizeof(l));
  l.x_hi = hs; l.x_lo = hs + 512; l.ldx = ldh;
  l.w_a = static_cast<const uint16_t*>(a->s1_a); l.w_b = static_cast<const uint16_t*>(a->s1_b);
  l.w_c = static_cast<const uint16_t*>(a->s1_c); l.ldw = a->ld_s1; l.bias = a->b_s1;
  l.y_hi = rs; l.y_lo = rs + 520; l.ldyh = ldr;
  l.m = bk; l.n = 256; l.k = 512; l.act = HOISDF_ACT_RELU; l.chunk_kb = 1; l.single_pass = 0; l.w_scale = a->s1_scale;
  st = hoisdf_linear_h3_fwd(&l, stream);
  if (st != HOISDF_OK) return st;
  st = hoisdf_posenc_split_fwd(n_index, nullptr, bk, a->bins, rs, rs + 520, ldr, stream);
  if (st != HOISDF_OK) return st;
  hoisdf_sdf_weights_h3 dw = *d;
  dw.chunk_kb = 1;
  dw.single_pass = 0;
  st = hoisdf_sdf_decoder_h3_fwd(&dw, rs, rs + 520, ldr, bk, hs, hs + 512, hs2, hs2 + 512, ldh, n_sdf, 0.f, stream);
  if (st != HOISDF_OK) return st;

  // ---- verdict of the screening step, on the device
  int n2 = 1;
  while (n2 < keep) n2 <<= 1;
  float* err_part = static_cast<float*>(at(L.err_part));
  screen_verify_kernel<<<static_cast<unsigned>(B), 1024, n2 * sizeof(float), s>>>(s_sdf, n_sdf, static_cast<int>(keep),
                                                                                  static_cast<int>(P), err_part, a->screen_gap);
  screen_verdict_kernel<<<1, 32, 0, s>>>(err_part, a->screen_gap, static_cast<int>(B), a->screen_err, a->verified);

  // ---- the P smallest |sdf| of the final values (upstream model.py:345-354: sort, take, clamp)
  st = hoisdf_select_points(n_sdf, n_offsets, n_index, B, P, a->bins, a->clamp, 0, a->sel_index,
                            static_cast<int32_t*>(at(L.s_sel)), a->points, a->sdf, a->posenc, a->status_flag, stream);
  if (st != HOISDF_OK) return st;
  if (a->cand_sdf != nullptr) {      // diagnostics for the parity tests: stage-A value of every candidate + its lattice index
    if (cudaMemcpyAsync(a->cand_sdf, sdf_all, total * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(a->cand_index, cand_index, total * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
      return HOISDF_E_SHAPE;
  }
  if (a->screen_rows != nullptr &&
      cudaMemcpyAsync(a->screen_rows, s_row, bk * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
    return HOISDF_E_SHAPE;
  if (a->exact_sdf != nullptr) {
    if (cudaMemcpyAsync(a->exact_sdf, n_sdf, bk * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(a->exact_index, n_index, bk * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
      return HOISDF_E_SHAPE;
  }
  return launch_status();
}
