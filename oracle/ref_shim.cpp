// TEST INFRASTRUCTURE.  extern "C" trampolines onto the reference's own C++-mangled functions
// (compiled from /root/reference by build_ref.sh) so ctypes can call them.
#include "common_function.h"
#include "lambda.h"
extern "C" {
int ref_lambda(int n, int m, const double* a, const double* Q, double* F, double* s) {
  return lambda(n, m, a, Q, F, s);
}
int ref_matinv(double* A, int n) { return matinv(A, n); }
double ref_distance(const double* rr, const double* rs, double* e) { return distance(rr, rs, e); }
double ref_velecitydistance(const double* rr, const double* rs, const double* vr, const double* vs,
                            double* e) {
  return velecitydistance(rr, rs, vr, vs, e);
}
double ref_dot(const double* a, const double* b, int n) { return dot(a, b, n); }
void ref_xyz2enu(const double* pos, double* E) { xyz2enu(pos, E); }
void ref_ecef2pos(const double* r, double* pos) { ecef2pos(r, pos); }
}
