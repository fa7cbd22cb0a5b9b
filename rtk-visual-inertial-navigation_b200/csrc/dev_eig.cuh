// Symmetric eigen-decomposition of a small dense matrix by one CTA: two-sided Jacobi in round-robin
// (tournament) order, np/2 disjoint rotations per step -- columns of A and V, then rows of A.  One warp
// per rotation pair, lanes over the rows / columns; the annihilated element is set to zero exactly; stops
// when the squared off-diagonal norm drops below 1e-30 of the squared diagonal norm (or after 30 sweeps).
// Stands in for Eigen::SelfAdjointEigenSolver in IMUGNSSBase::UpdateSchurComponent
// (RVI/factor/gnss_imu_factor.cpp:458-494) and SWFOptimization::UpdateSchur (RVI/swf/swf_gnss.cpp:44-45).
// A (n x n, leading dimension ld, full symmetric) is overwritten by diag(eigenvalues) + rounding-level
// off-diagonal; V must hold the identity on entry and holds the eigenvectors (columns) on exit.  A and V
// may live in shared or global memory; cs: 4 * ((n + 1) / 2) + 4 doubles and red: 34 doubles of shared memory.
#pragma once
#include "dev_common.cuh"

namespace swgn {

template <int NT>
__device__ void jacobi_eig(double* A, double* V, int n, int ld, double* cs, double* red) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int np = (n + 1) & ~1;  // players (an odd n gets a bye)
  const int half = np / 2;
  const int wid = tid >> 5, nwarp = NT / 32;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int ra = wid; ra < n; ra += nwarp)
      for (int cb = lane; cb < n; cb += 32) {
        const double val = A[ra * ld + cb] * A[ra * ld + cb];
        if (ra == cb) dg += val;
        else if (cb > ra) off += val;
      }
    off = block_sum(off, red);
    dg = block_sum(dg, red);
    if (off <= 1e-30 * (dg + 1e-300)) break;
    for (int step = 0; step < np - 1; ++step) {
      // pair t of this step: players pa, pb (circle method, player np-1 fixed)
      for (int t = tid; t < half; t += NT) {
        const int pa = (t == 0) ? np - 1 : (step + t) % (np - 1);
        const int pb = (step + np - 1 - t) % (np - 1);
        int pp = pa < pb ? pa : pb;
        const int qq = pa < pb ? pb : pa;
        double cth = 1.0, sth = 0.0;
        if (qq < n) {
          const double apq = A[pp * ld + qq];
          if (apq != 0.0) {
            const double theta = (A[qq * ld + qq] - A[pp * ld + pp]) / (2.0 * apq);
            const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            cth = 1.0 / sqrt(tt * tt + 1.0);
            sth = tt * cth;
          } else {
            pp = -1;  // nothing to rotate
          }
        } else {
          pp = -1;    // the bye
        }
        cs[4 * t] = cth;
        cs[4 * t + 1] = sth;
        cs[4 * t + 2] = (double)pp;
        cs[4 * t + 3] = (double)qq;
      }
      __syncthreads();
      for (int t = wid; t < half; t += nwarp) {  // columns: A <- A R, V <- V R
        const int pp = (int)cs[4 * t + 2];
        if (pp < 0) continue;
        const int qq = (int)cs[4 * t + 3];
        const double cth = cs[4 * t], sth = cs[4 * t + 1];
        for (int kk = lane; kk < n; kk += 32) {
          const double akp = A[kk * ld + pp], akq = A[kk * ld + qq];
          A[kk * ld + pp] = cth * akp - sth * akq;
          A[kk * ld + qq] = sth * akp + cth * akq;
          const double vkp = V[kk * ld + pp], vkq = V[kk * ld + qq];
          V[kk * ld + pp] = cth * vkp - sth * vkq;
          V[kk * ld + qq] = sth * vkp + cth * vkq;
        }
      }
      __syncthreads();
      for (int t = wid; t < half; t += nwarp) {  // rows: A <- R' A
        const int pp = (int)cs[4 * t + 2];
        if (pp < 0) continue;
        const int qq = (int)cs[4 * t + 3];
        const double cth = cs[4 * t], sth = cs[4 * t + 1];
        for (int kk = lane; kk < n; kk += 32) {
          const double apk = A[pp * ld + kk], aqk = A[qq * ld + kk];
          A[pp * ld + kk] = cth * apk - sth * aqk;
          A[qq * ld + kk] = sth * apk + cth * aqk;
        }
        __syncwarp();
        if (lane == 0) A[pp * ld + qq] = A[qq * ld + pp] = 0.0;  // annihilated exactly
      }
      __syncthreads();
    }
  }
}

}  // namespace swgn
