// Output tail of the hot path:
//   * joint voting (upstream common/nets/loss.py:31-36,54-57 -- the part of JointvoteLoss that produces
//     `hand_joints`, main/model.py:632-638)
//   * ManoHead + ManoLayer (common/nets/mano_head.py:185-256, manopth/manopth/manolayer.py:111-276):
//     rot6d -> R -> quaternion -> axis-angle -> Rodrigues -> blend shapes -> kinematic chain -> LBS.
// Both are tiny; one CTA per (layer, sample) keeps everything in shared memory and replaces the dozens of
// small ATen kernels the upstream code launches.
#include "common.cuh"

namespace hoisdf {

// grid L*B, 640 threads: warp j <-> joint j, lanes stride the points
__global__ void __launch_bounds__(640) vote_joints_kernel(const float* __restrict__ points,
                                                          const float* __restrict__ off,
                                                          const float* __restrict__ cls, int64_t batch, int64_t p,
                                                          float* __restrict__ joints) {
  const int64_t lb = blockIdx.x;
  const int64_t b = lb % batch;
  const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* c = cls + lb * p * 20 + j;
  const float* o = off + lb * p * 60 + j * 3;
  const float* pt = points + b * p * 3;
  float mx = -INFINITY;
  for (int64_t i = lane; i < p; i += 32) mx = fmaxf(mx, c[i * 20]);
  mx = warp_max(mx);
  float se = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
  for (int64_t i = lane; i < p; i += 32) {
    const float e = expf(c[i * 20] - mx);
    se += e;
    ax = fmaf(e, pt[i * 3 + 0] + o[i * 60 + 0], ax);
    ay = fmaf(e, pt[i * 3 + 1] + o[i * 60 + 1], ay);
    az = fmaf(e, pt[i * 3 + 2] + o[i * 60 + 2], az);
  }
  se = warp_sum(se); ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
  if (lane == 0) {
    float* d = joints + (lb * 20 + j) * 3;
    d[0] = __fdiv_rn(ax, se); d[1] = __fdiv_rn(ay, se); d[2] = __fdiv_rn(az, se);
  }
}

struct ManoParams {
  const float* __restrict__ shapedirs;
  const float* __restrict__ posedirs;
  const float* __restrict__ v_template;
  const float* __restrict__ j_regressor;
  const float* __restrict__ weights;
  const float* __restrict__ hands_mean;
  const float* __restrict__ pose6d;     // (N,16,6) rot6d, or NULL when pose_aa is given
  const float* __restrict__ pose_aa;    // (N,48) axis-angle (ground-truth MANO parameters), or NULL
  const float* __restrict__ betas;
  float* __restrict__ verts;
  float* __restrict__ joints;
};

__device__ __forceinline__ void normalize3(float* a) {  // F.normalize: a / max(||a||, 1e-12)
  const float n = fmaxf(sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), 1e-12f);
  a[0] = __fdiv_rn(a[0], n); a[1] = __fdiv_rn(a[1], n); a[2] = __fdiv_rn(a[2], n);
}

__device__ void rodrigues_rotation(const float* aa_in, const float* mean3, float* R);

// upstream mano_head.py:185-217 then rodrigues_layer.py:43-56 for one joint
__device__ void joint_rotation(const float* x6, const float* mean3, float* R) {
  float b1[3] = {x6[0], x6[1], x6[2]}, a2[3] = {x6[3], x6[4], x6[5]};
  normalize3(b1);
  const float d = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
  float b2[3] = {a2[0] - d * b1[0], a2[1] - d * b1[1], a2[2] - d * b1[2]};
  normalize3(b2);
  const float b3[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
  // M = [b1 b2 b3] as columns; upstream's mat2quat works on r = M^T
  float r[3][3];
  for (int i = 0; i < 3; ++i) { r[0][i] = b1[i]; r[1][i] = b2[i]; r[2][i] = b3[i]; }
  const bool d2 = r[2][2] < 1e-6f, d01 = r[0][0] > r[1][1], d0n1 = r[0][0] < -r[1][1];
  float q[4], t;
  if (d2 && d01) {
    t = 1.f + r[0][0] - r[1][1] - r[2][2];
    q[0] = r[1][2] - r[2][1]; q[1] = t; q[2] = r[0][1] + r[1][0]; q[3] = r[2][0] + r[0][2];
  } else if (d2) {
    t = 1.f - r[0][0] + r[1][1] - r[2][2];
    q[0] = r[2][0] - r[0][2]; q[1] = r[0][1] + r[1][0]; q[2] = t; q[3] = r[1][2] + r[2][1];
  } else if (d0n1) {
    t = 1.f - r[0][0] - r[1][1] + r[2][2];
    q[0] = r[0][1] - r[1][0]; q[1] = r[2][0] + r[0][2]; q[2] = r[1][2] + r[2][1]; q[3] = t;
  } else {
    t = 1.f + r[0][0] + r[1][1] + r[2][2];
    q[0] = t; q[1] = r[1][2] - r[2][1]; q[2] = r[2][0] - r[0][2]; q[3] = r[0][1] - r[1][0];
  }
  const float st = sqrtf(t);
  for (int i = 0; i < 4; ++i) q[i] = __fdiv_rn(q[i], st) * 0.5f;
  // quaternion -> axis-angle (mano_head.py:54-87), NaN -> 0 (:216)
  const float s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  const float s = sqrtf(s2);
  const float two_theta = 2.f * (q[0] < 0.f ? atan2f(-s, -q[0]) : atan2f(s, q[0]));
  const float kk = s2 > 0.f ? __fdiv_rn(two_theta, s) : 2.f;
  float aa[3];
  for (int i = 0; i < 3; ++i) {
    aa[i] = q[i + 1] * kk;
    if (isnan(aa[i])) aa[i] = 0.f;
  }
  rodrigues_rotation(aa, mean3, R);
}

// axis-angle (+ th_hands_mean for the 15 finger joints, manolayer.py:139-142) -> rotation matrix
__device__ void rodrigues_rotation(const float* aa_in, const float* mean3, float* R) {
  float aa[3];
  for (int i = 0; i < 3; ++i) aa[i] = aa_in[i] + (mean3 ? mean3[i] : 0.f);
  // Rodrigues through a quaternion (rodrigues_layer.py:43-56, :16-40)
  const float e0 = aa[0] + 1e-8f, e1 = aa[1] + 1e-8f, e2 = aa[2] + 1e-8f;
  const float ang = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
  const float hs = sinf(ang * 0.5f), hc = cosf(ang * 0.5f);
  float w = hc, x = hs * __fdiv_rn(aa[0], ang), y = hs * __fdiv_rn(aa[1], ang), z = hs * __fdiv_rn(aa[2], ang);
  const float qn = sqrtf(w * w + x * x + y * y + z * z);
  w = __fdiv_rn(w, qn); x = __fdiv_rn(x, qn); y = __fdiv_rn(y, qn); z = __fdiv_rn(z, qn);
  const float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
  const float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
  R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz;     R[2] = 2 * wy + 2 * xz;
  R[3] = 2 * wz + 2 * xy;   R[4] = w2 - x2 + y2 - z2;   R[5] = 2 * yz - 2 * wx;
  R[6] = 2 * xz - 2 * wy;   R[7] = 2 * wx + 2 * yz;     R[8] = w2 - x2 - y2 + z2;
}

// C(3x4) = A(3x4 as affine) * B(3x4 as affine)
__device__ __forceinline__ void affine_mul(const float* A, const float* B, float* C) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 4; ++j) {
      float s = A[i * 4 + 0] * B[0 * 4 + j] + A[i * 4 + 1] * B[1 * 4 + j] + A[i * 4 + 2] * B[2 * 4 + j];
      if (j == 3) s += A[i * 4 + 3];
      C[i * 4 + j] = s;
    }
  }
}

constexpr int NV = 778;

__global__ void __launch_bounds__(256) mano_kernel(const ManoParams p) {
  __shared__ float rot[16][9];
  __shared__ float pose_map[135];
  __shared__ float vs[NV * 3];     // v_shaped
  __shared__ float vp[NV * 3];     // v_posed, then posed vertices
  __shared__ float J[16][3];
  __shared__ float G[16][12];      // global joint transforms (3x4)
  __shared__ float A[16][12];      // with the rest-pose joint removed (results2)
  __shared__ float beta[10];
  const int64_t n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  if (tid < 16) {
    const float* mean3 = tid > 0 ? p.hands_mean + (tid - 1) * 3 : nullptr;
    if (p.pose_aa != nullptr) rodrigues_rotation(p.pose_aa + n * 48 + tid * 3, mean3, rot[tid]);
    else joint_rotation(p.pose6d + (n * 16 + tid) * 6, mean3, rot[tid]);
  }
  if (tid >= 32 && tid < 42) beta[tid - 32] = p.betas[n * 10 + tid - 32];
  __syncthreads();
  if (tid < 135) {
    const int k = tid / 9 + 1, e = tid % 9;
    pose_map[tid] = rot[k][e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
  }
  for (int i = tid; i < NV * 3; i += 256) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 10; ++k) s = fmaf(__ldg(p.shapedirs + i * 10 + k), beta[k], s);
    vs[i] = s + __ldg(p.v_template + i);
  }
  __syncthreads();
  // joints of the shaped template: 48 dot products of length 778
  for (int o = wid; o < 48; o += 8) {
    const int j = o / 3, c = o % 3;
    float s = 0.f;
    for (int v = lane; v < NV; v += 32) s = fmaf(__ldg(p.j_regressor + j * NV + v), vs[v * 3 + c], s);
    s = warp_sum(s);
    if (lane == 0) J[j][c] = s;
  }
  // pose blend shapes
  for (int i = tid; i < NV * 3; i += 256) {
    const float* pd = p.posedirs + static_cast<int64_t>(i) * 135;
    float s = 0.f;
    for (int k = 0; k < 135; ++k) s = fmaf(__ldg(pd + k), pose_map[k], s);
    vp[i] = vs[i] + s;
  }
  __syncthreads();
  // kinematic chain: thread f walks finger f (joints 1+3f, 2+3f, 3+3f); thread 0 also writes the root
  if (tid < 5) {
    float g0[12];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) g0[i * 4 + j] = rot[0][i * 3 + j];
      g0[i * 4 + 3] = J[0][i];
    }
    if (tid == 0) for (int e = 0; e < 12; ++e) G[0][e] = g0[e];
    float cur[12];
    for (int e = 0; e < 12; ++e) cur[e] = g0[e];
    int parent = 0;
    for (int lev = 0; lev < 3; ++lev) {
      const int jn = 1 + 3 * tid + lev;
      float rel[12], nxt[12];
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) rel[i * 4 + j] = rot[jn][i * 3 + j];
        rel[i * 4 + 3] = J[jn][i] - J[parent][i];
      }
      affine_mul(cur, rel, nxt);
      for (int e = 0; e < 12; ++e) { cur[e] = nxt[e]; G[jn][e] = nxt[e]; }
      parent = jn;
    }
  }
  __syncthreads();
  if (tid < 16) {
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) A[tid][i * 4 + j] = G[tid][i * 4 + j];
      A[tid][i * 4 + 3] = G[tid][i * 4 + 3] -
                          (G[tid][i * 4 + 0] * J[tid][0] + G[tid][i * 4 + 1] * J[tid][1] + G[tid][i * 4 + 2] * J[tid][2]);
    }
  }
  __syncthreads();
  // linear blend skinning; the root joint (centre) is subtracted afterwards
  const float cx = G[0][3], cy = G[0][7], cz = G[0][11];
  for (int v = tid; v < NV; v += 256) {
    float T[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) T[e] = 0.f;
    for (int j = 0; j < 16; ++j) {
      const float w = __ldg(p.weights + v * 16 + j);
#pragma unroll
      for (int e = 0; e < 12; ++e) T[e] = fmaf(w, A[j][e], T[e]);
    }
    const float x = vp[v * 3 + 0], y = vp[v * 3 + 1], z = vp[v * 3 + 2];
    const float ox = T[0] * x + T[1] * y + T[2] * z + T[3];
    const float oy = T[4] * x + T[5] * y + T[6] * z + T[7];
    const float oz = T[8] * x + T[9] * y + T[10] * z + T[11];
    vs[v * 3 + 0] = ox; vs[v * 3 + 1] = oy; vs[v * 3 + 2] = oz;  // vs is dead: reuse for posed vertices
    float* d = p.verts + (n * NV + v) * 3;
    // upstream scales to millimetres (manolayer.py:273-275) and ManoHead divides by 1000 again (mano_head.py:243-248)
    d[0] = __fdiv_rn((ox - cx) * 1000.f, 1000.f);
    d[1] = __fdiv_rn((oy - cy) * 1000.f, 1000.f);
    d[2] = __fdiv_rn((oz - cz) * 1000.f, 1000.f);
  }
  __syncthreads();
  if (tid < 21) {
    // manolayer.py:249-262: 16 chain joints + 5 fingertip vertices, then the visualisation reorder
    const int order[21] = {0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20};
    const int tips[5] = {745, 317, 444, 556, 673};
    const int s = order[tid];
    float x, y, z;
    if (s < 16) { x = G[s][3]; y = G[s][7]; z = G[s][11]; }
    else { const int v = tips[s - 16]; x = vs[v * 3]; y = vs[v * 3 + 1]; z = vs[v * 3 + 2]; }
    float* d = p.joints + (n * 21 + tid) * 3;
    d[0] = __fdiv_rn((x - cx) * 1000.f, 1000.f);
    d[1] = __fdiv_rn((y - cy) * 1000.f, 1000.f);
    d[2] = __fdiv_rn((z - cz) * 1000.f, 1000.f);
  }
}

}  // namespace hoisdf

using namespace hoisdf;

HOISDF_API int hoisdf_vote_joints_fwd(const float* points, const float* off, const float* cls, int64_t layers,
                                      int64_t batch, int64_t p, float* joints, void* stream) {
  if (points == nullptr || off == nullptr || cls == nullptr || joints == nullptr) return HOISDF_E_NULL;
  if (layers <= 0 || batch <= 0 || p <= 0 || layers * batch > 0x7fffffffLL) return HOISDF_E_SHAPE;
  HOISDF_LAUNCH(vote_joints_kernel, static_cast<unsigned>(layers * batch), 640, static_cast<cudaStream_t>(stream), points,
                off, cls, batch, p, joints);
  return launch_status();
}

HOISDF_API int hoisdf_mano_fwd(const hoisdf_mano_model* m, const float* pose6d, const float* betas, int64_t n,
                               float* verts, float* joints, void* stream) {
  if (m == nullptr || pose6d == nullptr || betas == nullptr || verts == nullptr || joints == nullptr)
    return HOISDF_E_NULL;
  if (m->shapedirs == nullptr || m->posedirs == nullptr || m->v_template == nullptr || m->j_regressor == nullptr ||
      m->weights == nullptr || m->hands_mean == nullptr)
    return HOISDF_E_NULL;
  if (n <= 0 || n > 0x7fffffffLL) return HOISDF_E_SHAPE;
  ManoParams p{m->shapedirs, m->posedirs, m->v_template, m->j_regressor, m->weights, m->hands_mean,
               pose6d, nullptr, betas, verts, joints};
  HOISDF_LAUNCH(mano_kernel, static_cast<unsigned>(n), 256, static_cast<cudaStream_t>(stream), p);
  return launch_status();
}

HOISDF_API int hoisdf_mano_aa_fwd(const hoisdf_mano_model* m, const float* pose_aa, const float* betas, int64_t n,
                                  float* verts, float* joints, void* stream) {
  if (m == nullptr || pose_aa == nullptr || betas == nullptr || verts == nullptr || joints == nullptr)
    return HOISDF_E_NULL;
  if (m->shapedirs == nullptr || m->posedirs == nullptr || m->v_template == nullptr || m->j_regressor == nullptr ||
      m->weights == nullptr || m->hands_mean == nullptr)
    return HOISDF_E_NULL;
  if (n <= 0 || n > 0x7fffffffLL) return HOISDF_E_SHAPE;
  ManoParams p{m->shapedirs, m->posedirs, m->v_template, m->j_regressor, m->weights, m->hands_mean,
               nullptr, pose_aa, betas, verts, joints};
  HOISDF_LAUNCH(mano_kernel, static_cast<unsigned>(n), 256, static_cast<cudaStream_t>(stream), p);
  return launch_status();
}
