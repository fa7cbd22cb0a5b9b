// Per-epoch GNSS preprocessing, device part: elevation of every observation
// (update_azel, RVI/gnss/src/common_function.cpp:394-408 -> ecef2pos :111-123, distance :126-134,
// satazel :84-100, xyz2enu :150-162) and the two cycle-slip gating residuals GnssPreprocess evaluates
// with RTKCarrierPhaseFactor(use_istd = false) (RVI/swf/swf_gnss.cpp:346-377, gnss_factor.cpp:105-119).
// One thread per observation, all epochs of the call in one launch; compiled with -fmad=false like the
// factor kernels so that the residuals round like the reference's scalar code.
#include <cuda_runtime.h>

#include <cstdint>

namespace swgn {
namespace {
constexpr double kPI = 3.1415926535897932;       // common_function.h PI
constexpr double kRE = 6378137.0;                // RE_WGS84
constexpr double kFE = 1.0 / 298.257223563;      // FE_WGS84
constexpr double kOMGE = 7.2921151467E-5;
constexpr double kCLIGHT = 299792458.0;

__device__ __forceinline__ double dot_hi(const double* a, const double* b, int n) {  // dot(): high index first
  double c = 0.0;
  while (--n >= 0) c += a[n] * b[n];
  return c;
}
__device__ double range_rtk(const double* rr, const double* rs, double* e) {  // distance()
  for (int i = 0; i < 3; ++i) e[i] = rr[i] - rs[i];
  const double r = sqrt(dot_hi(e, e, 3));
  for (int i = 0; i < 3; ++i) e[i] /= r;
  return r + kOMGE * (rs[0] * rr[1] - rs[1] * rr[0]) / kCLIGHT;
}
__device__ void ecef2pos_rtk(const double* r, double* pos) {
  const double e2 = kFE * (2.0 - kFE), r2 = dot_hi(r, r, 2);
  double z, zk, v = kRE, sinp;
  int guard = 0;
  for (z = r[2], zk = 0.0; fabs(z - zk) >= 1E-4 && guard < 64; ++guard) {
    zk = z;
    sinp = z / sqrt(r2 + z * z);
    v = kRE / sqrt(1.0 - e2 * sinp * sinp);
    z = r[2] + v * e2 * sinp;
  }
  pos[0] = r2 > 1E-12 ? atan(z / sqrt(r2)) : (r[2] > 0.0 ? kPI / 2.0 : -kPI / 2.0);
  pos[1] = r2 > 1E-12 ? atan2(r[1], r[0]) : 0.0;
  pos[2] = sqrt(r2 + z * z) - v;
}
__device__ double elevation_rtk(const double* pos, const double* e) {  // satazel(): el only
  double el = kPI / 2.0;
  if (pos[2] > -kRE) {
    const double sinp = sin(pos[0]), cosp = cos(pos[0]), sinl = sin(pos[1]), cosl = cos(pos[1]);
    // third row of xyz2enu's E (column-major E[2], E[5], E[8]) times e, summed in matmul's order
    double d = 0.0;
    d += (cosp * cosl) * e[0];
    d += (cosp * sinl) * e[1];
    d += sinp * e[2];
    el = asin(d);
  }
  return el;
}
}  // namespace

// rec: 16 doubles per observation {sat[3], receiver ECEF (pose + base)[3], unused[3], lam, L_rtk (cycles), N_rtk,
// clk_rtk, L_spp, N_spp, clk_spp}; flags bit0 = RTK residual wanted, bit1 = SPP residual wanted;
// out: {el, residual_rtk, residual_spp}
__global__ void k_gate_residuals(int n, const double* __restrict__ rec, const int32_t* __restrict__ flags, double azelmin,
                                 double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* q = rec + (size_t)16 * i;
  const double sat[3] = {q[0], q[1], q[2]}, rr[3] = {q[3], q[4], q[5]};
  double pos[3], e[3], e2[3];
  ecef2pos_rtk(rr, pos);
  const double rho = range_rtk(rr, sat, e);
  for (int k = 0; k < 3; ++k) e2[k] = -e[k];
  const double el = elevation_rtk(pos, e2);
  const double lam = q[9];
  const bool low = el < azelmin;  // swf_gnss.cpp:351-353: measurements of low satellites are zeroed first
  const int fl = flags ? flags[i] : 3;
  double r_rtk = 0.0, r_spp = 0.0;
  if (fl & 1) r_rtk = 1.0 * (rho - q[11] * lam - (low ? 0.0 : q[10]) * lam + q[12]);
  if (fl & 2) r_spp = 1.0 * (rho - q[14] * lam - (low ? 0.0 : q[13]) * lam + q[15]);
  out[3 * (size_t)i] = el;
  out[3 * (size_t)i + 1] = r_rtk;
  out[3 * (size_t)i + 2] = r_spp;
}

cudaError_t launch_gate_residuals(int n, const double* rec_dev, const int32_t* flags_dev, double azelmin, double* out_dev,
                                  cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  k_gate_residuals<<<(n + 127) / 128, 128, 0, s>>>(n, rec_dev, flags_dev, azelmin, out_dev);
  return cudaGetLastError();
}

}  // namespace swgn
