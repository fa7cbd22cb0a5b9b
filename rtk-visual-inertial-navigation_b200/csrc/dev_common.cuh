// Device-side helpers shared by all solver kernels: window view, block reductions, small math.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "device_types.h"
#include "kernels.cuh"

namespace swgn {

// View of one window's arrays; built once per CTA from the descriptor staged in shared memory.
struct Win {
  const WinDesc* d;
  const int32_t* ip;
  const double* cp;
  double* wp;
  __device__ __forceinline__ const int32_t* I(int a) const { return ip + d->ioff[a]; }
  __device__ __forceinline__ const double* C(int a) const { return cp + d->coff[a]; }
  __device__ __forceinline__ double* W(int a) const { return wp + d->woff[a]; }
};

// Stage WinDesc[w] into shared memory (all threads), return the view.
__device__ __forceinline__ Win load_window(const DeviceBatch& b, int w, WinDesc* sd) {
  const int32_t* src = reinterpret_cast<const int32_t*>(b.desc + w);
  int32_t* dst = reinterpret_cast<int32_t*>(sd);
  for (int i = threadIdx.x; i < (int)(sizeof(WinDesc) / 4); i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  Win v;
  v.d = sd;
  v.ip = b.ipool;
  v.cp = b.cpool;
  v.wp = b.wpool;
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Deterministic block-wide sum; every thread gets the result.  red: >= 33 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = lane < nw ? red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}
__device__ __forceinline__ double block_max(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = lane < nw ? red[lane] : 0.0;
    t = warp_max(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}
// block-wide OR of a predicate
__device__ __forceinline__ int block_any(int p) { return __syncthreads_or(p); }

// ---- quaternions (w, x, y, z), formulas as in Eigen / RVI/utility/utility.h:11-49 --------------
struct Quat {
  double w, x, y, z;
};
__device__ __forceinline__ Quat qmul(const Quat& a, const Quat& b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ Quat qinv(const Quat& q) {
  const double n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
  return {q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
}
__device__ __forceinline__ Quat qnormalized(const Quat& q) {
  const double n = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  return {q.w / n, q.x / n, q.y / n, q.z / n};
}
__device__ __forceinline__ void qrot(const Quat& q, const double v[3], double out[3]) {
  const double uv0 = 2.0 * (q.y * v[2] - q.z * v[1]), uv1 = 2.0 * (q.z * v[0] - q.x * v[2]),
               uv2 = 2.0 * (q.x * v[1] - q.y * v[0]);
  out[0] = v[0] + q.w * uv0 + (q.y * uv2 - q.z * uv1);
  out[1] = v[1] + q.w * uv1 + (q.z * uv0 - q.x * uv2);
  out[2] = v[2] + q.w * uv2 + (q.x * uv1 - q.y * uv0);
}
__device__ __forceinline__ void qtoR(const Quat& q, double R[9]) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
__device__ __forceinline__ Quat pose_q(const double* p) { return {p[6], p[3], p[4], p[5]}; }
__device__ __forceinline__ void skew3(const double v[3], double S[9]) {
  S[0] = 0;     S[1] = -v[2]; S[2] = v[1];
  S[3] = v[2];  S[4] = 0;     S[5] = -v[0];
  S[6] = -v[1]; S[7] = v[0];  S[8] = 0;
}
__device__ __forceinline__ void m33_mul(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
__device__ __forceinline__ void m33_vec(const double* A, const double* v, double* o) {
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}
// x [+] delta for one parameter block (PoseLocalParameterization::Plus or identity),
// RVI/factor/pose_local_parameterization.cpp:5-20
__device__ __forceinline__ void block_plus(const double* x, const double* dl, double* out, int gsize, int lsize) {
  if (gsize == 7 && lsize == 6) {
    out[0] = x[0] + dl[0];
    out[1] = x[1] + dl[1];
    out[2] = x[2] + dl[2];
    const Quat q = pose_q(x);
    const Quat dq = {1.0, dl[3] / 2.0, dl[4] / 2.0, dl[5] / 2.0};
    const Quat r = qnormalized(qmul(q, dq));
    out[3] = r.x; out[4] = r.y; out[5] = r.z; out[6] = r.w;
  } else {
    for (int k = 0; k < gsize; ++k) out[k] = x[k] + dl[k];
  }
}

__device__ __forceinline__ bool finite_d(double v) { return isfinite(v); }

}  // namespace swgn
