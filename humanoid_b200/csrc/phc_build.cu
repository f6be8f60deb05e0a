// Motion-library build on the device: phc_motion_build (include/phc_b200.h).
//
// Replaces the per-clip work of MotionLibSMPL.load_motions (PHC/motion_lib.py:257-428, worker
// load_motion_with_skeleton :748-824): optional random heading, local rotations, forward
// kinematics, velocities by np.gradient + gaussian_filter1d(sigma 2), angular velocities from
// consecutive global rotations, dof velocities from consecutive local rotations.  Three launches
// over all frames of all clips at once instead of ~1000 ATen / numpy / scipy calls per clip.
//
// The arithmetic follows the reference's dtypes and op order, which are mixed (SkeletonState /
// SkeletonMotion of PHC/poselib_skeleton.py, helpers in PHC/torch_utils.py):
//   * local rotations are computed in fp64 from the fp64 global rotations and ROUNDED to fp32
//     (quat_identity_like allocates fp32, poselib_skeleton.py:575-594);
//   * the root translation is rounded into the tree's fp32 local_translation (:606-620), so the
//     whole forward-kinematics chain (:519-539, transform_mul torch_utils.py:322-330) is fp32;
//   * np.gradient differentiates those fp32 positions in fp32; the gaussian filter accumulates in
//     fp64 (scipy's correlate1d: centre tap first, then symmetric pairs from the outside in) and
//     rounds to fp32 (:1231-1238);
//   * angular velocities stay fp64 until the final .float() (:1241-1251, motion_lib.py:401);
//   * dof velocities use the fp32 local rotations (compute_motion_dof_vels_jit, motion_lib.py:120).
// ATen's norm over 4 is ((x*x + y*y) + z*z) + w*w without FMA, over 3 it is
// fma(z,z,fma(y,y,x*x)) (measured, DESIGN.md "oracle notes"); both are reproduced here, which makes
// gts / grs / lrs / gvs bit-identical to the reference and gavs / dvs identical up to the last ulp of
// acos / atan2 / sin / cos.
//
// Compiled with -fmad=false: a*b+c stays two roundings, as in ATen's op-by-op evaluation.
#include <cstdint>

#include <cuda_runtime.h>

#include "../../include/phc_b200.h"

namespace phc {
int record_cuda_error(int cuda_error);  // phc_kernels.cu; returns PHC_ERR_CUDA
}

namespace {

constexpr int J = PHC_NUM_BODIES;
constexpr int R = PHC_BUILD_FILTER_RADIUS;

template <typename T>
struct Q4 {
  T x, y, z, w;
};
template <typename T>
struct V3 {
  T x, y, z;
};

// quat_mul, torch_utils.py:55-75 — the reference's 8-multiply form, term for term.
template <typename T>
__device__ __forceinline__ Q4<T> qmul(Q4<T> a, Q4<T> b) {
  const T ww = (a.z + a.x) * (b.x + b.y);
  const T yy = (a.w - a.y) * (b.w + b.z);
  const T zz = (a.w + a.y) * (b.w - b.z);
  const T xx = ww + yy + zz;
  const T qq = T(0.5) * (xx + (a.z - a.x) * (b.x - b.y));
  Q4<T> r;
  r.w = qq - ww + (a.z - a.y) * (b.y - b.z);
  r.x = qq - xx + (a.x + a.w) * (b.x + b.w);
  r.y = qq - yy + (a.w - a.x) * (b.y + b.z);
  r.z = qq - zz + (a.z + a.y) * (b.w - b.x);
  return r;
}
template <typename T>
__device__ __forceinline__ Q4<T> qconj(Q4<T> q) {
  return {-q.x, -q.y, -q.z, q.w};
}
__device__ __forceinline__ float sqrt_rn(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_rn(double x) { return sqrt(x); }

// quat_normalize = quat_unit(quat_pos(q)), torch_utils.py:154-196
template <typename T>
__device__ __forceinline__ Q4<T> qnormalize(Q4<T> q) {
  if (q.w < T(0)) q = {-q.x, -q.y, -q.z, -q.w};
  T n = sqrt_rn(((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w);
  n = n < T(1e-9) ? T(1e-9) : n;
  return {q.x / n, q.y / n, q.z / n, q.w / n};
}
template <typename T>
__device__ __forceinline__ Q4<T> qmul_norm(Q4<T> a, Q4<T> b) {  // torch_utils.py:254-259
  return qnormalize(qmul(a, b));
}
// quat_rotate, torch_utils.py:263-268
__device__ __forceinline__ V3<float> qrotate(Q4<float> rot, V3<float> v) {
  const Q4<float> r = qmul(qmul(rot, Q4<float>{v.x, v.y, v.z, 0.0f}), qconj(rot));
  return {r.x, r.y, r.z};
}

struct BuildParams {
  const double* quat;    // [F,24,4]
  const double* trans;   // [F,3]
  const double* aa;      // [F,72] or null
  const float* lt;       // [M,24,3]
  const int64_t* nf;     // [M]
  const int64_t* starts; // [M]
  const double* fps;     // [M]
  const double* heading; // [M,2] (z, w) or null
  int64_t F, M;
  float *gts, *grs, *lrs, *gvs, *gavs, *dvs, *motion_aa;
  double* av_raw;  // [F,24,3] scratch: unfiltered angular velocity
  float* vel_raw;  // [F,24,3] scratch: np.gradient(gts) / dt
  double w[2 * R + 1];
  int8_t parent[J];
  int8_t depth[J];
  int max_depth;
};

// clip of a global frame index: the last m with starts[m] <= f
__device__ __forceinline__ int clip_of(const int64_t* __restrict__ starts, int64_t M, int64_t f) {
  int64_t lo = 0, hi = M;  // invariant: starts[lo] <= f, answer in [lo, hi)
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(starts + mid) <= f) lo = mid;
    else hi = mid;
  }
  return (int)lo;
}

// The global rotation the reference works on: the file's quaternion, or — with the random heading
// (motion_lib.py:789-797) — heading * from_quat(q): scipy normalises q, then composes.  The heading
// quaternion is (0, 0, z, w), so only the surviving terms of the product are kept.
__device__ __forceinline__ Q4<double> load_global_rot(const BuildParams& p, int64_t f, int j, int m) {
  const double4 v = *reinterpret_cast<const double4*>(p.quat + (f * J + j) * 4);
  Q4<double> q{v.x, v.y, v.z, v.w};
  if (p.heading) {
    const double hz = __ldg(p.heading + 2 * m), hw = __ldg(p.heading + 2 * m + 1);
    const double n = sqrt(((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w);
    q = {q.x / n, q.y / n, q.z / n, q.w / n};
    q = {hw * q.x - hz * q.y, hw * q.y + hz * q.x, hw * q.z + hz * q.w, hw * q.w - hz * q.z};
  }
  return q;
}

template <typename T>
__device__ __forceinline__ T shfl(T v, int src) {
  return __shfl_sync(0xffffffffu, v, src);
}

// scipy Rotation.from_rotvec -> compose with the heading -> as_rotvec, for pose_aa[:, :3] (:794)
__device__ __forceinline__ V3<double> heading_on_rotvec(V3<double> v, double hz, double hw) {
  const double ang = sqrt((v.x * v.x + v.y * v.y) + v.z * v.z);
  double sc;
  if (ang <= 1e-3) {
    const double a2 = ang * ang;
    sc = 0.5 - a2 / 48.0 + a2 * a2 / 3840.0;
  } else {
    sc = sin(ang / 2) / ang;
  }
  Q4<double> q{sc * v.x, sc * v.y, sc * v.z, cos(ang / 2)};
  q = {hw * q.x - hz * q.y, hw * q.y + hz * q.x, hw * q.z + hz * q.w, hw * q.w - hz * q.z};
  if (q.w < 0) q = {-q.x, -q.y, -q.z, -q.w};
  const double n3 = sqrt((q.x * q.x + q.y * q.y) + q.z * q.z);
  const double a = 2 * atan2(n3, q.w);
  double s2;
  if (a <= 1e-3) {
    const double a2 = a * a;
    s2 = 2 + a2 / 12 + 7 * a2 * a2 / 2880;
  } else {
    s2 = a / sin(a / 2);
  }
  return {s2 * q.x, s2 * q.y, s2 * q.z};
}

// ---------------------------------------------------------------------------------------------
// Kernel A — one warp per frame, lane = joint: grs, lrs, forward kinematics -> gts, motion_aa.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) build_pose_kernel(const BuildParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t f = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (f >= p.F) return;
  const int m = clip_of(p.starts, p.M, f);
  const int j = lane < J ? lane : 0;
  const int par = p.parent[j];

  const Q4<double> g = load_global_rot(p, f, j, m);
  Q4<float> lr;
  if (par < 0) {
    lr = {(float)g.x, (float)g.y, (float)g.z, (float)g.w};
  } else {  // fp64 product, rounded on the store into the fp32 tensor (poselib_skeleton.py:586-590)
    const Q4<double> l = qmul_norm(qconj(load_global_rot(p, f, par, m)), g);
    lr = {(float)l.x, (float)l.y, (float)l.z, (float)l.w};
  }
  V3<float> lt;
  if (par < 0) {
    double tx = p.trans[f * 3 + 0], ty = p.trans[f * 3 + 1], tz = p.trans[f * 3 + 2];
    if (p.heading) {  // trans @ as_matrix().T of the z rotation (:798)
      const double hz = __ldg(p.heading + 2 * m), hw = __ldg(p.heading + 2 * m + 1);
      const double zz = hz * hz, ww = hw * hw, zw = hz * hw;
      const double c = ww - zz, s = 2 * zw;
      const double nx = tx * c + ty * (-s), ny = tx * s + ty * c;
      tx = nx;
      ty = ny;
      tz = tz * (zz + ww);
    }
    lt = {(float)tx, (float)ty, (float)tz};
  } else {
    const float* t = p.lt + ((int64_t)m * J + j) * 3;
    lt = {__ldg(t), __ldg(t + 1), __ldg(t + 2)};
  }
  if (lane < J) {
    *reinterpret_cast<float4*>(p.grs + (f * J + j) * 4) = make_float4((float)g.x, (float)g.y, (float)g.z, (float)g.w);
    *reinterpret_cast<float4*>(p.lrs + (f * J + j) * 4) = make_float4(lr.x, lr.y, lr.z, lr.w);
  }

  // forward kinematics, one tree level per round: the parent's global transform comes by shuffle
  Q4<float> gr = lr;
  V3<float> gt = lt;
  const int my_depth = p.depth[j];
  const int src = par < 0 ? 0 : par;
  for (int level = 1; level <= p.max_depth; ++level) {
    const Q4<float> pr{shfl(gr.x, src), shfl(gr.y, src), shfl(gr.z, src), shfl(gr.w, src)};
    const V3<float> pt{shfl(gt.x, src), shfl(gt.y, src), shfl(gt.z, src)};
    if (my_depth == level) {  // transform_mul(parent, local)
      gr = qmul_norm(pr, lr);
      const V3<float> r = qrotate(pr, lt);
      gt = {r.x + pt.x, r.y + pt.y, r.z + pt.z};
    }
  }
  if (lane < J) {
    float* o = p.gts + (f * J + j) * 3;
    o[0] = gt.x;
    o[1] = gt.y;
    o[2] = gt.z;
  }

  if (p.aa) {  // _motion_aa: pose_aa as fp32 (:377,:391), root rotated by the heading (:794)
    const double* a = p.aa + f * 72;
    float* o = p.motion_aa + f * 72;
    for (int e = lane; e < 72; e += 32) {
      double v = a[e];
      if (p.heading && e < 3) {
        const V3<double> rv = heading_on_rotvec({a[0], a[1], a[2]}, __ldg(p.heading + 2 * m), __ldg(p.heading + 2 * m + 1));
        v = e == 0 ? rv.x : (e == 1 ? rv.y : rv.z);
      }
      o[e] = (float)v;
    }
  }
}

// np.gradient over the clip's frame axis, uniform spacing 1, edge_order 1; then / time_delta, all in fp32
__device__ __forceinline__ float gradient_at(const float* __restrict__ col, int64_t i, int64_t nf, float dtf) {
  float g;
  if (i == 0) g = col[72] - col[0];
  else if (i == nf - 1) g = col[i * 72] - col[(i - 1) * 72];
  else g = (col[(i + 1) * 72] - col[(i - 1) * 72]) / 2.0f;
  return g / dtf;
}

// ---------------------------------------------------------------------------------------------
// Kernel B — one thread per (frame, joint): unfiltered linear (fp32) and angular (fp64) velocity into the
//            scratch, and the dof velocity.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) build_diff_kernel(const BuildParams p) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.F * J) return;
  const int64_t f = idx / J;
  const int j = (int)(idx - f * J);
  const int m = clip_of(p.starts, p.M, f);
  const int64_t k = f - __ldg(p.starts + m);
  const int64_t nf = __ldg(p.nf + m);
  const double fps = __ldg(p.fps + m);
  const double td = 1.0 / fps;

  // _compute_angular_velocity, poselib_skeleton.py:1241-1246: the last frame's difference is identity
  V3<double> av{0.0, 0.0, 0.0};
  if (k < nf - 1) {
    const Q4<double> d = qmul_norm(load_global_rot(p, f + 1, j, m), qconj(load_global_rot(p, f, j, m)));
    double s = 2 * (d.w * d.w) - 1;  // quat_angle_axis, torch_utils.py:219-228
    s = s < -1.0 ? -1.0 : (s > 1.0 ? 1.0 : s);
    const double angle = acos(s);
    double n = sqrt(__fma_rn(d.z, d.z, __fma_rn(d.y, d.y, d.x * d.x)));
    n = n < 1e-9 ? 1e-9 : n;
    av = {(d.x / n) * angle / td, (d.y / n) * angle / td, (d.z / n) * angle / td};
  }
  double* o = p.av_raw + idx * 3;
  o[0] = av.x;
  o[1] = av.y;
  o[2] = av.z;

  if (nf >= 2) {  // _compute_velocity, poselib_skeleton.py:1231-1236, before the filter
    const float dtf = (float)td;
    const float* col = p.gts + (f - k) * 72 + j * 3;
    float* v = p.vel_raw + idx * 3;
    v[0] = gradient_at(col, k, nf, dtf);
    v[1] = gradient_at(col + 1, k, nf, dtf);
    v[2] = gradient_at(col + 2, k, nf, dtf);
  }

  if (j == 0) return;
  // compute_motion_dof_vels_jit, motion_lib.py:120-142: frames (k, k+1); the last frame repeats
  float* dv = p.dvs + (f * (J - 1) + (j - 1)) * 3;
  if (nf < 2) {
    dv[0] = dv[1] = dv[2] = 0.0f;
    return;
  }
  const int64_t fa = f - k + (k < nf - 1 ? k : nf - 2);
  const float4 a4 = *reinterpret_cast<const float4*>(p.lrs + (fa * J + j) * 4);
  const float4 b4 = *reinterpret_cast<const float4*>(p.lrs + ((fa + 1) * J + j) * 4);
  const Q4<float> d = qmul(qconj(Q4<float>{a4.x, a4.y, a4.z, a4.w}), Q4<float>{b4.x, b4.y, b4.z, b4.w});
  // quat_to_angle_axis, torch_utils.py:86-106
  const float sin_theta = sqrtf(1.0f - d.w * d.w);
  float angle = 2.0f * acosf(d.w);
  angle = atan2f(sinf(angle), cosf(angle));
  const bool ok = fabsf(sin_theta) > 1e-5f;
  const float dtf = (float)td;
  V3<float> axis = ok ? V3<float>{d.x / sin_theta, d.y / sin_theta, d.z / sin_theta} : V3<float>{0.0f, 0.0f, 1.0f};
  angle = ok ? angle : 0.0f;
  dv[0] = axis.x * angle / dtf;
  dv[1] = axis.y * angle / dtf;
  dv[2] = axis.z * angle / dtf;
}

// ---------------------------------------------------------------------------------------------
// Kernel C — one thread per (frame, joint*3+c): gaussian_filter1d(sigma 2, mode "nearest") of the two
//            unfiltered velocities -> gvs, gavs.  scipy's correlate1d accumulates in fp64: centre tap
//            first, then the symmetric pairs from the outside in.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) build_filter_kernel(const BuildParams p) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.F * 72) return;
  const int64_t f = idx / 72;
  const int e = (int)(idx - f * 72);
  const int m = clip_of(p.starts, p.M, f);
  const int64_t s0 = __ldg(p.starts + m);
  const int64_t k = f - s0;
  const int64_t nf = __ldg(p.nf + m);
  if (nf < 2) {
    p.gvs[idx] = 0.0f;
    p.gavs[idx] = 0.0f;
    return;
  }
  const float* vel = p.vel_raw + s0 * 72 + e;
  const double* av = p.av_raw + s0 * 72 + e;
  auto clampi = [nf](int64_t i) { return i < 0 ? (int64_t)0 : (i > nf - 1 ? nf - 1 : i); };
  double acc_v = (double)vel[k * 72] * p.w[R];
  double acc_a = av[k * 72] * p.w[R];
#pragma unroll
  for (int jj = -R; jj < 0; ++jj) {
    const int64_t lo = clampi(k + jj), hi = clampi(k - jj);
    acc_v += ((double)vel[lo * 72] + (double)vel[hi * 72]) * p.w[R + jj];
    acc_a += (av[lo * 72] + av[hi * 72]) * p.w[R + jj];
  }
  p.gvs[idx] = (float)acc_v;
  p.gavs[idx] = (float)acc_a;
}

}  // namespace

extern "C" int phc_motion_build(const PhcBuildArgs* a, phc_stream_t stream) {
  if (!a) return PHC_ERR_NULL;
  if (a->total_frames < 0 || a->num_motions < 0) return PHC_ERR_SHAPE;
  if (a->total_frames == 0 || a->num_motions == 0) return a->total_frames == 0 ? PHC_OK : PHC_ERR_SHAPE;
  if (!a->pose_quat_global || !a->root_trans || !a->local_translation || !a->num_frames || !a->length_starts ||
      !a->fps || !a->parent_indices_host || !a->filter_weights_host || !a->gts || !a->grs || !a->lrs || !a->gvs ||
      !a->gavs || !a->dvs || !a->scratch)
    return PHC_ERR_NULL;
  if ((a->pose_aa == nullptr) != (a->motion_aa == nullptr)) return PHC_ERR_NULL;
  if (((uintptr_t)a->pose_quat_global & 31) || ((uintptr_t)a->grs & 15) || ((uintptr_t)a->lrs & 15)) return PHC_ERR_ALIGN;
  if (a->total_frames > (int64_t)1 << 40) return PHC_ERR_SHAPE;

  BuildParams p;
  p.quat = a->pose_quat_global;
  p.trans = a->root_trans;
  p.aa = a->pose_aa;
  p.lt = a->local_translation;
  p.nf = a->num_frames;
  p.starts = a->length_starts;
  p.fps = a->fps;
  p.heading = a->heading_zw;
  p.F = a->total_frames;
  p.M = a->num_motions;
  p.gts = a->gts;
  p.grs = a->grs;
  p.lrs = a->lrs;
  p.gvs = a->gvs;
  p.gavs = a->gavs;
  p.dvs = a->dvs;
  p.motion_aa = a->motion_aa;
  p.av_raw = a->scratch;
  p.vel_raw = reinterpret_cast<float*>(a->scratch + a->total_frames * J * 3);
  for (int i = 0; i < 2 * R + 1; ++i) p.w[i] = a->filter_weights_host[i];
  p.max_depth = 0;
  for (int j = 0; j < J; ++j) {  // topological order: a parent precedes its children (SkeletonTree.from_mjcf)
    const int par = a->parent_indices_host[j];
    if (j == 0 ? par != -1 : (par < 0 || par >= j)) return PHC_ERR_SHAPE;
    p.parent[j] = (int8_t)par;
    p.depth[j] = j == 0 ? 0 : (int8_t)(p.depth[par] + 1);
    if (p.depth[j] > p.max_depth) p.max_depth = p.depth[j];
  }

  const int64_t F = p.F;
  const int wpb = 8;
  build_pose_kernel<<<(unsigned)((F + wpb - 1) / wpb), wpb * 32, 0, stream>>>(p);
  build_diff_kernel<<<(unsigned)((F * J + 255) / 256), 256, 0, stream>>>(p);
  build_filter_kernel<<<(unsigned)((F * 72 + 255) / 256), 256, 0, stream>>>(p);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? PHC_OK : phc::record_cuda_error((int)e);
}
