// Synthetic sliding-window generator (see swgn_synth.h).  Workload tooling only.
#include "swgn_synth.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace {

// ---------------------------------------------------------------- RNG
struct Rng {
  uint64_t s;
  static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() { return splitmix(s); }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  double uni(double a, double b) { return a + (b - a) * uni(); }
  int uni_int(int a, int b) { return a + (int)(next() % (uint64_t)(b - a + 1)); }
  double normal() {
    double u1 = uni(), u2 = uni();
    if (u1 < 1e-300) u1 = 1e-300;
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
};

// ---------------------------------------------------------------- small math
struct V3 {
  double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
struct M3 {
  double m[9];
};
inline M3 mul(const M3& A, const M3& B) {
  M3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C.m[i * 3 + j] = A.m[i * 3] * B.m[j] + A.m[i * 3 + 1] * B.m[3 + j] + A.m[i * 3 + 2] * B.m[6 + j];
  return C;
}
inline M3 T(const M3& A) {
  M3 B;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) B.m[i * 3 + j] = A.m[j * 3 + i];
  return B;
}
inline V3 mul(const M3& A, V3 v) {
  return {A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z,
          A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z};
}
inline M3 skew(V3 v) { return {{0, -v.z, v.y, v.z, 0, -v.x, -v.y, v.x, 0}}; }
inline M3 I3() { return {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
inline M3 add(const M3& A, const M3& B, double sb = 1.0) {
  M3 C;
  for (int i = 0; i < 9; ++i) C.m[i] = A.m[i] + sb * B.m[i];
  return C;
}
inline M3 scale(const M3& A, double s) {
  M3 C;
  for (int i = 0; i < 9; ++i) C.m[i] = s * A.m[i];
  return C;
}
struct Q4 {
  double w, x, y, z;
};
inline Q4 qmul(Q4 a, Q4 b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
inline Q4 qnorm(Q4 q) {
  double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  return {q.w / n, q.x / n, q.y / n, q.z / n};
}
inline M3 qR(Q4 q) {
  double x = q.x, y = q.y, z = q.z, w = q.w;
  return {{1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w),
           1 - 2 * (x * x + z * z), 2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w),
           1 - 2 * (x * x + y * y)}};
}
inline V3 qrot(Q4 q, V3 v) { return mul(qR(q), v); }
Q4 R2q(const M3& R) {
  double tr = R.m[0] + R.m[4] + R.m[8];
  Q4 q;
  if (tr > 0) {
    double s = std::sqrt(tr + 1.0) * 2;
    q = {0.25 * s, (R.m[7] - R.m[5]) / s, (R.m[2] - R.m[6]) / s, (R.m[3] - R.m[1]) / s};
  } else if (R.m[0] > R.m[4] && R.m[0] > R.m[8]) {
    double s = std::sqrt(1.0 + R.m[0] - R.m[4] - R.m[8]) * 2;
    q = {(R.m[7] - R.m[5]) / s, 0.25 * s, (R.m[1] + R.m[3]) / s, (R.m[2] + R.m[6]) / s};
  } else if (R.m[4] > R.m[8]) {
    double s = std::sqrt(1.0 + R.m[4] - R.m[0] - R.m[8]) * 2;
    q = {(R.m[2] - R.m[6]) / s, (R.m[1] + R.m[3]) / s, 0.25 * s, (R.m[5] + R.m[7]) / s};
  } else {
    double s = std::sqrt(1.0 + R.m[8] - R.m[0] - R.m[4]) * 2;
    q = {(R.m[3] - R.m[1]) / s, (R.m[2] + R.m[6]) / s, (R.m[5] + R.m[7]) / s, 0.25 * s};
  }
  q = qnorm(q);
  if (q.w < 0) q = {-q.w, -q.x, -q.y, -q.z};
  return q;
}

// dense n x n helpers (row-major) for the 15x15 pre-integration algebra
typedef std::vector<double> Md;
Md mm(const Md& A, int ar, int ac, const Md& B, int bc) {
  Md C((size_t)ar * bc, 0.0);
  for (int i = 0; i < ar; ++i)
    for (int k = 0; k < ac; ++k) {
      double a = A[i * ac + k];
      if (a == 0.0) continue;
      for (int j = 0; j < bc; ++j) C[i * bc + j] += a * B[k * bc + j];
    }
  return C;
}
Md tr(const Md& A, int r, int c) {
  Md B((size_t)r * c);
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) B[j * r + i] = A[i * c + j];
  return B;
}
bool inv_lu(const Md& A, int n, Md* out) {
  Md lu = A;
  std::vector<int> piv(n);
  for (int i = 0; i < n; ++i) piv[i] = i;
  for (int k = 0; k < n; ++k) {
    int p = k;
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(lu[i * n + k]) > std::fabs(lu[p * n + k])) p = i;
    if (lu[p * n + k] == 0.0) return false;
    if (p != k) {
      for (int j = 0; j < n; ++j) std::swap(lu[k * n + j], lu[p * n + j]);
      std::swap(piv[k], piv[p]);
    }
    for (int i = k + 1; i < n; ++i) {
      lu[i * n + k] /= lu[k * n + k];
      for (int j = k + 1; j < n; ++j) lu[i * n + j] -= lu[i * n + k] * lu[k * n + j];
    }
  }
  out->assign((size_t)n * n, 0.0);
  std::vector<double> x(n);
  for (int c = 0; c < n; ++c) {
    for (int i = 0; i < n; ++i) x[i] = piv[i] == c ? 1.0 : 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < i; ++j) x[i] -= lu[i * n + j] * x[j];
    for (int i = n - 1; i >= 0; --i) {
      for (int j = i + 1; j < n; ++j) x[i] -= lu[i * n + j] * x[j];
      x[i] /= lu[i * n + i];
    }
    for (int i = 0; i < n; ++i) (*out)[i * n + c] = x[i];
  }
  return true;
}
bool chol_lower(const Md& A, int n, Md* L) {
  L->assign((size_t)n * n, 0.0);
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= (*L)[j * n + k] * (*L)[j * n + k];
    if (!(d > 0)) return false;
    d = std::sqrt(d);
    (*L)[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= (*L)[i * n + k] * (*L)[j * n + k];
      (*L)[i * n + j] = s / d;
    }
  }
  return true;
}

// ---------------------------------------------------------------- IMU pre-integration
// midpoint scheme with first-order bias Jacobian and covariance, following
// RVI/factor/integration_base.cpp:30-142
struct Preint {
  V3 acc0, gyr0, ba, bg;
  V3 dp{0, 0, 0}, dv{0, 0, 0};
  Q4 dq{1, 0, 0, 0};
  double sum_dt = 0;
  V3 gyri, gyrj;
  Md jac, cov, noise;
  Preint(V3 a0, V3 g0, V3 ba_, V3 bg_, double an, double gn, double aw, double gw)
      : acc0(a0), gyr0(g0), ba(ba_), bg(bg_), gyri(g0), gyrj(g0) {
    jac.assign(225, 0.0);
    for (int i = 0; i < 15; ++i) jac[i * 15 + i] = 1.0;
    cov.assign(225, 0.0);
    noise.assign(18 * 18, 0.0);
    double d[6] = {an * an, gn * gn, an * an, gn * gn, aw * aw, gw * gw};
    for (int b = 0; b < 6; ++b)
      for (int i = 0; i < 3; ++i) noise[(b * 3 + i) * 18 + b * 3 + i] = d[b];
  }
  static void put(Md& M, int ld, int r, int c, const M3& B) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) M[(r + i) * ld + c + j] = B.m[i * 3 + j];
  }
  void push(double dt, V3 acc1, V3 gyr1) {
    gyrj = gyr1;
    M3 Rq = qR(dq);
    V3 un_acc0 = mul(Rq, acc0 - ba);
    V3 un_gyr = 0.5 * (gyr0 + gyr1) - bg;
    Q4 ndq = qmul(dq, Q4{1, un_gyr.x * dt / 2, un_gyr.y * dt / 2, un_gyr.z * dt / 2});
    M3 Rn = qR(ndq);  // (not yet normalised, like the reference's toRotationMatrix on result_delta_q)
    {
      // Eigen's toRotationMatrix assumes a unit quaternion but is evaluated on the un-normalised
      // product in the reference; the 1e-6-level scale error is part of its covariance model.
      double x = ndq.x, y = ndq.y, z = ndq.z, w = ndq.w;
      Rn = {{1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w),
             1 - 2 * (x * x + z * z), 2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w),
             1 - 2 * (x * x + y * y)}};
    }
    // q * v for an un-normalised quaternion in Eigen also goes through toRotationMatrix
    V3 un_acc1 = mul(Rn, acc1 - ba);
    V3 un_acc = 0.5 * (un_acc0 + un_acc1);
    V3 ndp = dp + dt * dv + (0.5 * dt * dt) * un_acc;
    V3 ndv = dv + dt * un_acc;
    // Jacobian / covariance
    M3 Rw = skew(un_gyr), Ra0 = skew(acc0 - ba), Ra1 = skew(acc1 - ba);
    M3 ImRw = add(I3(), Rw, -dt);
    Md F(225, 0.0), V(15 * 18, 0.0);
    put(F, 15, 0, 0, I3());
    put(F, 15, 0, 3,
        add(scale(mul(Rq, Ra0), -0.25 * dt * dt), scale(mul(mul(Rn, Ra1), ImRw), -0.25 * dt * dt)));
    put(F, 15, 0, 6, scale(I3(), dt));
    put(F, 15, 0, 9, scale(add(Rq, Rn), -0.25 * dt * dt));
    put(F, 15, 0, 12, scale(mul(Rn, Ra1), -0.25 * dt * dt * -dt));
    put(F, 15, 3, 3, ImRw);
    put(F, 15, 3, 12, scale(I3(), -dt));
    put(F, 15, 6, 3, add(scale(mul(Rq, Ra0), -0.5 * dt), scale(mul(mul(Rn, Ra1), ImRw), -0.5 * dt)));
    put(F, 15, 6, 6, I3());
    put(F, 15, 6, 9, scale(add(Rq, Rn), -0.5 * dt));
    put(F, 15, 6, 12, scale(mul(Rn, Ra1), -0.5 * dt * -dt));
    put(F, 15, 9, 9, I3());
    put(F, 15, 12, 12, I3());
    put(V, 18, 0, 0, scale(Rq, 0.25 * dt * dt));
    M3 v03 = scale(mul(Rn, Ra1), -0.25 * dt * dt * 0.5 * dt);
    put(V, 18, 0, 3, v03);
    put(V, 18, 0, 6, scale(Rn, 0.25 * dt * dt));
    put(V, 18, 0, 9, v03);
    put(V, 18, 3, 3, scale(I3(), 0.5 * dt));
    put(V, 18, 3, 9, scale(I3(), 0.5 * dt));
    put(V, 18, 6, 0, scale(Rq, 0.5 * dt));
    M3 v63 = scale(mul(Rn, Ra1), -0.5 * dt * 0.5 * dt);
    put(V, 18, 6, 3, v63);
    put(V, 18, 6, 6, scale(Rn, 0.5 * dt));
    put(V, 18, 6, 9, v63);
    put(V, 18, 9, 12, scale(I3(), dt));
    put(V, 18, 12, 15, scale(I3(), dt));
    jac = mm(F, 15, 15, jac, 15);
    Md FC = mm(F, 15, 15, cov, 15);
    Md FCFt = mm(FC, 15, 15, tr(F, 15, 15), 15);
    Md VN = mm(V, 15, 18, noise, 18);
    Md VNVt = mm(VN, 15, 18, tr(V, 15, 18), 15);
    for (int i = 0; i < 225; ++i) cov[i] = FCFt[i] + VNVt[i];
    dp = ndp;
    dv = ndv;
    dq = qnorm(ndq);
    sum_dt += dt;
    acc0 = acc1;
    gyr0 = gyr1;
  }
  bool sqrt_info(Md* out) const {  // LLT(cov^-1).matrixL().transpose()
    Md inv, L;
    if (!inv_lu(cov, 15, &inv)) return false;
    for (int i = 0; i < 15; ++i)
      for (int j = 0; j < i; ++j) inv[i * 15 + j] = inv[j * 15 + i] = 0.5 * (inv[i * 15 + j] + inv[j * 15 + i]);
    if (!chol_lower(inv, 15, &L)) return false;
    *out = tr(L, 15, 15);
    return true;
  }
};

const double kPi = 3.14159265358979323846;
const double kClight = 299792458.0;
const double kOmge = 7.2921151467E-5;

}  // namespace

struct swgn_synth {
  swgn_synth_config cfg;
  swgn_graph g;
  swgn_options opt;
  std::vector<int32_t> block_size, block_manifold, block_const, block_group, block_offset;
  std::vector<double> state, truth;
  std::vector<int32_t> proj_blocks, imu_blocks, gnss_kind, gnss_blocks;
  std::vector<double> proj_uv, imu_data, gnss_data;
  std::vector<int32_t> prior_n, prior_blk_begin, prior_blocks, prior_blk_idx;
  std::vector<int64_t> prior_x0_begin, prior_J_begin, prior_r_begin;
  std::vector<double> prior_x0, prior_J, prior_r0;
  std::vector<int32_t> unit_block;
  std::vector<double> unit_istd;
  std::vector<int32_t> epoch_begin, obs_amb, obs_sysfreq;
  std::vector<int32_t> chain_blk_begin, chain_blocks, chain_frame_begin, chain_frame_block;
  std::vector<double> chain_frame_data, chain_frame_N, chain_N, chain_imu_data, chain_frame_truth;
  std::vector<double> true_N;
  int32_t info[8];
};

extern "C" {

void swgn_synth_default_config(int32_t which, swgn_synth_config* c) {
  std::memset(c, 0, sizeof(*c));
  c->seed0 = 20261017ull;
  c->state_noise = 1.0;
  c->bias_walk_scale = 1.0;
  if (which == 1) {
    c->n_keyframes = 5;
    c->n_landmarks = 50;
    c->n_gnss_epochs = 0;
    c->n_sats = 0;
  } else if (which == 4) {
    c->n_keyframes = 6;
    c->n_landmarks = 40;
    c->n_gnss_epochs = 2;
    c->n_sats = 8;
    c->composition = 1;
    c->hidden_per_gap = 2;
    c->bias_walk_scale = 30.0;
    c->hidden_bias_istd = 10.0;
  } else {
    c->n_keyframes = 20;
    c->n_landmarks = 300;
    c->n_gnss_epochs = 10;
    c->n_sats = 20;
    if (which == 3) {
      c->composition = 1;
      c->hidden_per_gap = 2;
      c->bias_walk_scale = 30.0;
      c->hidden_bias_istd = 10.0;
    }
  }
}

swgn_synth* swgn_synth_create(const swgn_synth_config* cfg, uint64_t window_id) {
  swgn_synth* S = new swgn_synth();
  S->cfg = *cfg;
  uint64_t sd = cfg->seed0 + window_id;
  Rng rng(Rng::splitmix(sd));
  const int nkf = cfg->n_keyframes, nep = cfg->n_gnss_epochs, nsat = nep > 0 ? cfg->n_sats : 0;
  const double sn = cfg->state_noise;
  const bool compA = cfg->composition == 1;
  const int hpg = std::max(1, std::min(3, (int)cfg->hidden_per_gap));

  // ---- yaml constants (YAML/rtk_visual_inertial_config.yaml:24-28,68-123)
  const double bws = cfg->bias_walk_scale > 0.0 ? cfg->bias_walk_scale : 1.0;
  const double ACC_N = 0.05, GYR_N = 0.005, ACC_W = 0.0005 * bws, GYR_W = 0.00005 * bws, GNORM = 9.8;
  const V3 Pbg = {-0.0051302024, 0.0091942546, 0.308739733};
  const M3 RIC = {{-1.1283524065062611e-02, 9.0570010831436121e-03, 9.9989532092917277e-01,
                   -9.9992100257025784e-01, -5.6404389398068133e-03, -1.1232723065088990e-02,
                   5.5381137189322582e-03, -9.9994307646982916e-01, 9.1199296318514866e-03}};
  const V3 TIC = {1.3224454035460147e-02, 5.7114724738452263e-02, -1.5241815653778757e-02};
  const V3 base = {-2323932.39454, 5387298.51324, 2493096.51920};
  const double lams[3] = {0.190293672798364871256993069437, 0.19203948631027648, 0.19029367279836487};

  // ---- ENU -> ECEF rotation at the anchor (Rwgw): geodetic latitude by fixed-point iteration
  double lat, lon;
  {
    const double a = 6378137.0, f = 1.0 / 298.257223563, e2 = f * (2.0 - f);
    double r2 = base.x * base.x + base.y * base.y, z = base.z, zk = 0.0, v = a, sinp;
    for (int it = 0; it < 50 && std::fabs(z - zk) >= 1e-4; ++it) {
      zk = z;
      sinp = z / std::sqrt(r2 + z * z);
      v = a / std::sqrt(1.0 - e2 * sinp * sinp);
      z = base.z + v * e2 * sinp;
    }
    lat = std::atan(z / std::sqrt(r2));
    lon = std::atan2(base.y, base.x);
  }
  const double sp = std::sin(lat), cp = std::cos(lat), sl = std::sin(lon), cl = std::cos(lon);
  // columns: East, North, Up expressed in ECEF
  const M3 Rwgw = {{-sl, -sp * cl, cp * cl, cl, -sp * sl, cp * sl, 0, cp, sp}};
  const V3 g_w = mul(Rwgw, V3{0, 0, GNORM});

  // ---- frames: keyframes every 0.25 s, a GNSS frame 0.05 s after every 2nd keyframe
  struct Frame {
    double t;
    int is_gnss, kf_index, epoch;
  };
  std::vector<Frame> frames;
  {
    int ep = 0;
    for (int k = 0; k < nkf; ++k) {
      frames.push_back({0.25 * k, 0, k, -1});
      if (nep > 0 && (k % 2 == 1) && ep < nep) {
        if (!compA) {
          frames.push_back({0.25 * k + 0.05, 1, -1, ep});
          ++ep;
        } else if (k + 1 < nkf) {  // hidden frames need a keyframe on both sides
          for (int h = 0; h < hpg; ++h) frames.push_back({0.25 * k + 0.05 * (h + 1), 1, -1, ep});
          ++ep;
        }
      }
    }
  }
  const int F = (int)frames.size();
  int n_epochs_real = 0;
  for (auto& fr : frames) n_epochs_real += fr.is_gnss;
  const int n_gnss_frames = n_epochs_real;
  (void)n_gnss_frames;
  if (compA) n_epochs_real = 0;  // no raw GNSS factors, clock or drift blocks in composition A

  // ---- trajectory (IMU origin, ENU): figure-8 of radius 10 m, ~2-3 m/s, +-3 deg roll/pitch
  const double Rr = 10.0, om = 0.2;
  const double t0 = rng.uni(0.0, 2 * kPi / om), ph_r = rng.uni(0, 2 * kPi), ph_p = rng.uni(0, 2 * kPi);
  auto pos_enu = [&](double t) {
    double s = t + t0;
    return V3{Rr * std::sin(om * s), 0.5 * Rr * std::sin(2 * om * s), 0.3 * std::sin(0.5 * s)};
  };
  auto vel_enu = [&](double t) {
    double s = t + t0;
    return V3{Rr * om * std::cos(om * s), Rr * om * std::cos(2 * om * s), 0.15 * std::cos(0.5 * s)};
  };
  auto acc_enu = [&](double t) {
    double s = t + t0;
    return V3{-Rr * om * om * std::sin(om * s), -2 * Rr * om * om * std::sin(2 * om * s),
              -0.075 * std::sin(0.5 * s)};
  };
  auto rot_enu = [&](double t) {
    V3 v = vel_enu(t);
    double yaw = std::atan2(v.y, v.x);
    double roll = 3.0 * kPi / 180 * std::sin(0.7 * (t + t0) + ph_r);
    double pitch = 3.0 * kPi / 180 * std::sin(0.9 * (t + t0) + ph_p);
    M3 Rz = {{std::cos(yaw), -std::sin(yaw), 0, std::sin(yaw), std::cos(yaw), 0, 0, 0, 1}};
    M3 Ry = {{std::cos(pitch), 0, std::sin(pitch), 0, 1, 0, -std::sin(pitch), 0, std::cos(pitch)}};
    M3 Rx = {{1, 0, 0, 0, std::cos(roll), -std::sin(roll), 0, std::sin(roll), std::cos(roll)}};
    return mul(mul(Rz, Ry), Rx);
  };
  auto rot_w = [&](double t) { return mul(Rwgw, rot_enu(t)); };
  auto omega_b = [&](double t) {
    const double h = 1e-5;
    M3 M = mul(T(rot_enu(t - h)), rot_enu(t + h));
    return V3{(M.m[7] - M.m[5]) / (4 * h), (M.m[2] - M.m[6]) / (4 * h), (M.m[3] - M.m[1]) / (4 * h)};
  };
  auto pimu_w = [&](double t) { return mul(Rwgw, pos_enu(t)); };
  auto vimu_w = [&](double t) { return mul(Rwgw, vel_enu(t)); };
  auto sf_b = [&](double t) { return mul(T(rot_w(t)), mul(Rwgw, acc_enu(t)) + g_w); };

  // ---- block layout
  const int b_pose = 0, b_sb = F, b_ext = 2 * F, b_lm = 2 * F + 1;
  const int b_N = b_lm + cfg->n_landmarks;
  const int b_clk = b_N + nsat;
  const int b_drift = b_clk + 3 * n_epochs_real;
  const int b_black = b_drift + n_epochs_real;               // only with GNSS
  const int b_black2 = b_black + (n_epochs_real > 0 ? 1 : 0);
  const int nb = b_black2 + 1;
  S->block_size.assign(nb, 1);
  S->block_manifold.assign(nb, SWGN_MANIFOLD_EUCLIDEAN);
  S->block_const.assign(nb, 0);
  S->block_group.assign(nb, -1);
  for (int f = 0; f < F; ++f) {
    S->block_size[b_pose + f] = 7;
    S->block_manifold[b_pose + f] = SWGN_MANIFOLD_POSE;
    S->block_size[b_sb + f] = 9;
  }
  S->block_size[b_ext] = 7;
  S->block_manifold[b_ext] = SWGN_MANIFOLD_POSE;
  S->block_const[b_ext] = (cfg->variant & SWGN_SYNTH_FREE_EXTRINSIC) ? 0 : 1;  // ESTIMATE_EXTRINSIC
  for (int l = 0; l < cfg->n_landmarks; ++l) S->block_size[b_lm + l] = 3;
  S->block_offset.resize(nb);
  int off = 0;
  for (int i = 0; i < nb; ++i) {
    S->block_offset[i] = off;
    off += S->block_size[i];
  }
  S->truth.assign(off, 0.0);
  S->state.assign(off, 0.0);
  double* X = S->truth.data();
  auto bp = [&](int b) { return X + S->block_offset[b]; };

  // ---- true states
  V3 ba_true = {0.05 * rng.normal(), 0.05 * rng.normal(), 0.05 * rng.normal()};
  V3 bg_true = {0.005 * rng.normal(), 0.005 * rng.normal(), 0.005 * rng.normal()};
  for (int f = 0; f < F; ++f) {
    double t = frames[f].t;
    M3 R = rot_w(t);
    V3 w = omega_b(t);
    V3 P = pimu_w(t) + mul(R, Pbg);             // antenna position
    V3 Vv = vimu_w(t) + mul(R, cross(w, Pbg));  // antenna velocity
    Q4 q = R2q(R);
    double* p = bp(b_pose + f);
    p[0] = P.x; p[1] = P.y; p[2] = P.z; p[3] = q.x; p[4] = q.y; p[5] = q.z; p[6] = q.w;
    double* s = bp(b_sb + f);
    s[0] = Vv.x; s[1] = Vv.y; s[2] = Vv.z;
    s[3] = ba_true.x; s[4] = ba_true.y; s[5] = ba_true.z;
    s[6] = bg_true.x; s[7] = bg_true.y; s[8] = bg_true.z;
  }
  {
    Q4 q = R2q(RIC);
    double* p = bp(b_ext);
    p[0] = TIC.x; p[1] = TIC.y; p[2] = TIC.z; p[3] = q.x; p[4] = q.y; p[5] = q.z; p[6] = q.w;
  }

  // ---- initial (perturbed) state for the frames, needed as bias linearisation point
  std::memcpy(S->state.data(), S->truth.data(), sizeof(double) * off);
  double* X0 = S->state.data();
  auto bp0 = [&](int b) { return X0 + S->block_offset[b]; };
  if (cfg->variant & SWGN_SYNTH_FREE_EXTRINSIC) {  // a fixed offset (no random draw: the default windows keep their seeds)
    double* p = bp0(b_ext);
    const double dt[3] = {0.01, -0.008, 0.005}, th[3] = {0.004, -0.003, 0.002};
    for (int k = 0; k < 3; ++k) p[k] += sn * dt[k];
    Q4 q = qnorm(qmul(Q4{p[6], p[3], p[4], p[5]}, Q4{1, sn * th[0] / 2, sn * th[1] / 2, sn * th[2] / 2}));
    p[3] = q.x; p[4] = q.y; p[5] = q.z; p[6] = q.w;
  }
  for (int f = 0; f < F; ++f) {
    double* p = bp0(b_pose + f);
    for (int k = 0; k < 3; ++k) p[k] += sn * 0.05 * rng.normal();
    double th[3] = {sn * 0.5 * kPi / 180 * rng.normal(), sn * 0.5 * kPi / 180 * rng.normal(),
                    sn * 0.5 * kPi / 180 * rng.normal()};
    Q4 q = qnorm(qmul(Q4{p[6], p[3], p[4], p[5]}, Q4{1, th[0] / 2, th[1] / 2, th[2] / 2}));
    p[3] = q.x; p[4] = q.y; p[5] = q.z; p[6] = q.w;
    double* s = bp0(b_sb + f);
    for (int k = 0; k < 3; ++k) s[k] += sn * 0.05 * rng.normal();
    for (int k = 3; k < 6; ++k) s[k] += sn * 0.02 * rng.normal();
    for (int k = 6; k < 9; ++k) s[k] += sn * 0.002 * rng.normal();
  }

  // ---- IMU factors between consecutive frames (400 Hz)
  const double dti = 0.0025;
  std::vector<double> imu_rec;  // record f: frame f -> f + 1
  for (int f = 0; f + 1 < F; ++f) {
    int k0 = (int)std::llround(frames[f].t / dti), k1 = (int)std::llround(frames[f + 1].t / dti);
    auto sample = [&](int k, V3* a, V3* w) {
      double t = k * dti;
      *a = sf_b(t) + ba_true + V3{ACC_N * rng.normal(), ACC_N * rng.normal(), ACC_N * rng.normal()};
      *w = omega_b(t) + bg_true + V3{GYR_N * rng.normal(), GYR_N * rng.normal(), GYR_N * rng.normal()};
    };
    V3 a0, w0;
    sample(k0, &a0, &w0);
    const double* s0 = bp0(b_sb + f);
    Preint pre(a0, w0, V3{s0[3], s0[4], s0[5]}, V3{s0[6], s0[7], s0[8]}, ACC_N, GYR_N, ACC_W, GYR_W);
    for (int k = k0 + 1; k <= k1; ++k) {
      V3 a, w;
      sample(k, &a, &w);
      pre.push(dti, a, w);
    }
    Md sq;
    if (!pre.sqrt_info(&sq)) {
      delete S;
      return nullptr;
    }
    size_t base_i = imu_rec.size();
    imu_rec.resize(base_i + SWGN_IMU_STRIDE, 0.0);
    double* r = imu_rec.data() + base_i;
    r[SWGN_IMU_DELTA_P] = pre.dp.x; r[SWGN_IMU_DELTA_P + 1] = pre.dp.y; r[SWGN_IMU_DELTA_P + 2] = pre.dp.z;
    r[SWGN_IMU_DELTA_Q] = pre.dq.x; r[SWGN_IMU_DELTA_Q + 1] = pre.dq.y;
    r[SWGN_IMU_DELTA_Q + 2] = pre.dq.z; r[SWGN_IMU_DELTA_Q + 3] = pre.dq.w;
    r[SWGN_IMU_DELTA_V] = pre.dv.x; r[SWGN_IMU_DELTA_V + 1] = pre.dv.y; r[SWGN_IMU_DELTA_V + 2] = pre.dv.z;
    r[SWGN_IMU_LIN_BA] = pre.ba.x; r[SWGN_IMU_LIN_BA + 1] = pre.ba.y; r[SWGN_IMU_LIN_BA + 2] = pre.ba.z;
    r[SWGN_IMU_LIN_BG] = pre.bg.x; r[SWGN_IMU_LIN_BG + 1] = pre.bg.y; r[SWGN_IMU_LIN_BG + 2] = pre.bg.z;
    r[SWGN_IMU_GYRI] = pre.gyri.x; r[SWGN_IMU_GYRI + 1] = pre.gyri.y; r[SWGN_IMU_GYRI + 2] = pre.gyri.z;
    r[SWGN_IMU_GYRJ] = pre.gyrj.x; r[SWGN_IMU_GYRJ + 1] = pre.gyrj.y; r[SWGN_IMU_GYRJ + 2] = pre.gyrj.z;
    r[SWGN_IMU_SUM_DT] = pre.sum_dt;
    std::memcpy(r + SWGN_IMU_JACOBIAN, pre.jac.data(), sizeof(double) * 225);
    std::memcpy(r + SWGN_IMU_SQRT_INFO, sq.data(), sizeof(double) * 225);
    if (compA && (frames[f].is_gnss || frames[f + 1].is_gnss)) continue;  // part of a chain factor
    int32_t ib[4] = {b_pose + f, b_sb + f, b_pose + f + 1, b_sb + f + 1};
    S->imu_blocks.insert(S->imu_blocks.end(), ib, ib + 4);
    S->imu_data.insert(S->imu_data.end(), r, r + SWGN_IMU_STRIDE);
  }

  // ---- landmarks and visual tracks
  std::vector<int> kf_frame;
  for (int f = 0; f < F; ++f)
    if (!frames[f].is_gnss) kf_frame.push_back(f);
  for (int l = 0; l < cfg->n_landmarks; ++l) {
    for (int attempt = 0; attempt < 100; ++attempt) {
      int len = rng.uni_int(2, 10);
      if (len > nkf) len = nkf;
      int k_first = rng.uni_int(0, nkf - len);
      int f0 = kf_frame[k_first];
      M3 Rwb = rot_w(frames[f0].t);
      V3 c = pimu_w(frames[f0].t) + mul(Rwb, TIC);
      // keep the depth observable: at most 10x the baseline spanned by the track
      double t_last = frames[kf_frame[k_first + len - 1]].t;
      double baseline = norm(pimu_w(t_last) + mul(rot_w(t_last), TIC) - c);
      double dmax = 10.0 * baseline;
      if (dmax > 30.0) dmax = 30.0;
      if (dmax < 3.5) dmax = 3.5;
      double d = rng.uni(3.0, dmax), u = rng.uni(-0.45, 0.45), v = rng.uni(-0.3, 0.3);
      V3 Xw = c + mul(mul(Rwb, RIC), V3{u * d, v * d, d});
      std::vector<int> fr;
      std::vector<double> uv;
      for (int k = k_first; k < nkf && (int)fr.size() < len; ++k) {
        int f = kf_frame[k];
        M3 Rb = rot_w(frames[f].t);
        V3 cc = pimu_w(frames[f].t) + mul(Rb, TIC);
        V3 pc = mul(T(mul(Rb, RIC)), Xw - cc);
        if (pc.z < 1.0 || std::fabs(pc.x / pc.z) > 1.5 || std::fabs(pc.y / pc.z) > 1.0) break;
        fr.push_back(f);
        uv.push_back(pc.x / pc.z + 1e-3 * rng.normal());
        uv.push_back(pc.y / pc.z + 1e-3 * rng.normal());
      }
      if ((int)fr.size() < 2 && attempt < 99) continue;
      double* p = bp(b_lm + l);
      p[0] = Xw.x; p[1] = Xw.y; p[2] = Xw.z;
      for (size_t i = 0; i < fr.size(); ++i) {
        int32_t pb[3] = {b_pose + fr[i], b_ext, b_lm + l};
        S->proj_blocks.insert(S->proj_blocks.end(), pb, pb + 3);
        S->proj_uv.push_back(uv[2 * i]);
        S->proj_uv.push_back(uv[2 * i + 1]);
      }
      break;
    }
    double* p0 = bp0(b_lm + l);
    const double* pt = bp(b_lm + l);
    for (int k = 0; k < 3; ++k) p0[k] = pt[k] + sn * 0.2 * rng.normal();
  }

  // ---- GNSS
  S->true_N.assign(nsat, 0.0);
  std::vector<int> sat_sys(nsat);
  std::vector<V3> sat_pos0(nsat), sat_vel(nsat);
  std::vector<double> sat_el(nsat);
  if (nsat > 0) {
    int n_gps = (nsat * 8 + 10) / 20, n_bds = (nsat * 7 + 10) / 20;
    if (n_gps + n_bds > nsat) n_bds = nsat - n_gps;
    for (int s = 0; s < nsat; ++s) {
      sat_sys[s] = s < n_gps ? 0 : (s < n_gps + n_bds ? 1 : 2);
      double el = rng.uni(30.0, 80.0) * kPi / 180, az = rng.uni(0, 2 * kPi);
      V3 dir = mul(Rwgw, V3{std::cos(el) * std::sin(az), std::cos(el) * std::cos(az), std::sin(el)});
      double rs = rng.uni(2.0e7, 2.6e7), bd = dot(base, dir);
      double rho = -bd + std::sqrt(bd * bd - dot(base, base) + rs * rs);
      sat_pos0[s] = base + rho * dir;
      V3 rnd = {rng.normal(), rng.normal(), rng.normal()};
      V3 perp = cross(sat_pos0[s], rnd);
      sat_vel[s] = (3000.0 / norm(perp)) * perp;
      sat_el[s] = el;
      S->true_N[s] = (double)rng.uni_int(-50, 50);
      bp(b_N + s)[0] = S->true_N[s];
      bp0(b_N + s)[0] = S->true_N[s] + sn * 0.3 * rng.normal();
    }
  }
  auto range_sagnac = [&](V3 rr, V3 rs) {
    return norm(rr - rs) + kOmge * (rs.x * rr.y - rs.y * rr.x) / kClight;
  };
  S->epoch_begin.push_back(0);
  if (compA) {
    // IMUGNSSFactor chains (RVI/factor/gnss_imu_factor.cpp:264-356 AddMargInfo): every hidden GNSS
    // frame carries A = J'J, b = J'r0 of its carrier-phase / pseudorange / Doppler rows linearised
    // at the frame's initial estimate (ambiguities at 0), split into pose_hessians (15x15),
    // pose_phase_biases_hessians (15xk), pose_rhses, and accumulated phase_biases_hessians / rhs.
    const int k = nsat;
    S->chain_blk_begin.push_back(0);
    S->chain_frame_begin.push_back(0);
    int f = 0;
    while (f < F) {
      if (!frames[f].is_gnss) { ++f; continue; }
      int f1 = f;
      while (f1 < F && frames[f1].is_gnss) ++f1;  // hidden run [f, f1), keyframes f-1 and f1
      const int m = f1 - f;
      std::vector<double> NN((size_t)k * k, 0.0), Nr(k, 0.0);
      for (int h = f; h < f1; ++h) {
        const double t = frames[h].t;
        const double* pl = bp0(b_pose + h);
        const double* sl = bp0(b_sb + h);
        const double* pt = bp(b_pose + h);
        const double* st = bp(b_sb + h);
        size_t o = S->chain_frame_data.size();
        S->chain_frame_data.resize(o + SWGN_CHAIN_FRAME_STRIDE, 0.0);
        double* fr = S->chain_frame_data.data() + o;
        for (int q = 0; q < 7; ++q) fr[SWGN_CHAIN_POSE + q] = fr[SWGN_CHAIN_POSE_LIN + q] = pl[q];
        for (int q = 0; q < 9; ++q) fr[SWGN_CHAIN_SB + q] = fr[SWGN_CHAIN_SB_LIN + q] = sl[q];
        for (int q = 0; q < 7; ++q) S->chain_frame_truth.push_back(pt[q]);
        for (int q = 0; q < 9; ++q) S->chain_frame_truth.push_back(st[q]);
        S->chain_frame_block.push_back(b_pose + h);
        std::vector<double> PN((size_t)15 * k, 0.0);
        double* Hp = fr + SWGN_CHAIN_HESSIAN;
        double* bpv = fr + SWGN_CHAIN_RHS;
        const V3 xt = V3{pt[0], pt[1], pt[2]} + base, xl = V3{pl[0], pl[1], pl[2]} + base;
        const V3 vt = {st[0], st[1], st[2]}, vl = {sl[0], sl[1], sl[2]};
        for (int s = 0; s < nsat; ++s) {
          const V3 sp = sat_pos0[s] + t * sat_vel[s];
          const double lam = lams[sat_sys[s]];
          const double sinel = std::sin(sat_el[s]);
          const V3 el = (1.0 / norm(xl - sp)) * (xl - sp);
          const double ev[3] = {el.x, el.y, el.z};
          const double rho_t = range_sagnac(xt, sp), rho_l = range_sagnac(xl, sp);
          // carrier phase: r = w (rho - N lam - L)
          {
            const double w = sinel / (0.003 * lam);
            const double L = rho_t - S->true_N[s] * lam + 0.003 * lam * rng.normal();
            const double r0 = w * (rho_l - L), jn = -w * lam;
            for (int a = 0; a < 3; ++a) {
              for (int c = 0; c < 3; ++c) Hp[a * 15 + c] += w * ev[a] * w * ev[c];
              PN[(size_t)a * k + s] += w * ev[a] * jn;
              bpv[a] += w * ev[a] * r0;
            }
            NN[(size_t)s * k + s] += jn * jn;
            Nr[s] += jn * r0;
          }
          // pseudorange: r = w (rho - P)
          {
            const double w = sinel / 0.3;
            const double P1 = rho_t + 0.3 * rng.normal();
            const double r0 = w * (rho_l - P1);
            for (int a = 0; a < 3; ++a) {
              for (int c = 0; c < 3; ++c) Hp[a * 15 + c] += w * ev[a] * w * ev[c];
              bpv[a] += w * ev[a] * r0;
            }
          }
          // Doppler: r = w ((v - v_sat).e + D)
          {
            const double w = sinel * sinel / 0.05;
            const V3 et = (1.0 / norm(xt - sp)) * (xt - sp);
            const double D = -dot(vt - sat_vel[s], et) + 0.05 * rng.normal();
            const double r0 = w * (dot(vl - sat_vel[s], el) + D);
            for (int a = 0; a < 3; ++a) {
              for (int c = 0; c < 3; ++c) Hp[(6 + a) * 15 + 6 + c] += w * ev[a] * w * ev[c];
              bpv[6 + a] += w * ev[a] * r0;
            }
          }
          S->obs_amb.push_back(s);
          S->obs_sysfreq.push_back(sat_sys[s] * 2);
        }
        if (cfg->hidden_bias_istd > 0.0) {  // weak absolute bias information at the linearisation point
          const double wa = cfg->hidden_bias_istd, wg = 10.0 * cfg->hidden_bias_istd;
          for (int a = 0; a < 3; ++a) {
            Hp[(9 + a) * 15 + 9 + a] += wa * wa;
            Hp[(12 + a) * 15 + 12 + a] += wg * wg;
          }
        }
        S->epoch_begin.push_back((int32_t)S->obs_amb.size());
        S->chain_frame_N.insert(S->chain_frame_N.end(), PN.begin(), PN.end());
      }
      S->chain_N.insert(S->chain_N.end(), NN.begin(), NN.end());
      S->chain_N.insert(S->chain_N.end(), Nr.begin(), Nr.end());
      for (int h = f - 1; h < f1; ++h)
        S->chain_imu_data.insert(S->chain_imu_data.end(), imu_rec.begin() + (size_t)SWGN_IMU_STRIDE * h,
                                 imu_rec.begin() + (size_t)SWGN_IMU_STRIDE * (h + 1));
      int32_t cb[4] = {b_pose + f - 1, b_sb + f - 1, b_pose + f1, b_sb + f1};
      S->chain_blocks.insert(S->chain_blocks.end(), cb, cb + 4);
      for (int s = 0; s < nsat; ++s) S->chain_blocks.push_back(b_N + s);
      S->chain_blk_begin.push_back((int32_t)S->chain_blocks.size());
      S->chain_frame_begin.push_back(S->chain_frame_begin.back() + m);
      f = f1;
    }
  }
  for (int f = 0; f < F && !compA; ++f) {
    if (!frames[f].is_gnss) continue;
    const int e = frames[f].epoch;
    const double t = frames[f].t;
    const double* pt = bp(b_pose + f);
    const double* st = bp(b_sb + f);
    V3 xg = V3{pt[0], pt[1], pt[2]} + base;
    V3 vr = {st[0], st[1], st[2]};
    double clk[3], drift = rng.normal();
    for (int k = 0; k < 3; ++k) {
      clk[k] = 3.0 * rng.normal();
      bp(b_clk + 3 * e + k)[0] = clk[k];
    }
    bp(b_drift + e)[0] = drift;
    const double br_dt = 0.1;
    for (int s = 0; s < nsat; ++s) {
      V3 sp = sat_pos0[s] + t * sat_vel[s];
      const int sys = sat_sys[s];
      const double lam = lams[sys];
      const double rho = range_sagnac(xg, sp);
      const double Lstd = 0.003, Pstd = 0.3;
      auto weight = [&](double var) {
        double b = kClight * 5e-12 * br_dt;
        double sinel = sinf(sat_el[s]);  // single precision, like gnss_factor.cpp:100
        return 1.0 / std::sqrt((var / sinel / sinel) + b * b);
      };
      auto add = [&](int kind, int b0, int b1, int b2, double meas, double w, double var) {
        S->gnss_kind.push_back(kind);
        int32_t gb[3] = {b0, b1, b2};
        S->gnss_blocks.insert(S->gnss_blocks.end(), gb, gb + 3);
        size_t o = S->gnss_data.size();
        S->gnss_data.resize(o + SWGN_GNSS_STRIDE, 0.0);
        double* r = S->gnss_data.data() + o;
        r[0] = sp.x; r[1] = sp.y; r[2] = sp.z;
        r[3] = sat_vel[s].x; r[4] = sat_vel[s].y; r[5] = sat_vel[s].z;
        r[6] = base.x; r[7] = base.y; r[8] = base.z;
        r[SWGN_GNSS_MEAS] = meas;
        r[SWGN_GNSS_LAM] = lam;
        r[SWGN_GNSS_WEIGHT] = w;
        r[SWGN_GNSS_EL] = sat_el[s];
        r[SWGN_GNSS_DT] = br_dt;
        r[SWGN_GNSS_VAR] = var;
      };
      // RB-SD carrier phase: r = w (rho - N lam - L1_lam + clk)
      double var_l = (Lstd * lam) * (Lstd * lam);
      double L1_lam = rho - S->true_N[s] * lam + clk[sys] + Lstd * lam * rng.normal();
      if (cfg->variant & SWGN_SYNTH_SPP)  // rover-only carrier phase: (pose, clk, N), r = istd (rho + clk - N lam - L1_lam)
        add(SWGN_GNSS_SPP_CARRIER, b_pose + f, b_clk + 3 * e + sys, b_N + s, L1_lam, weight(var_l), var_l);
      else
        add(SWGN_GNSS_RTK_CARRIER, b_pose + f, b_N + s, b_clk + 3 * e + sys, L1_lam, weight(var_l), var_l);
      // RB-SD pseudorange: r = w (rho - P1 + clk)
      double var_p = Pstd * Pstd;
      double P1 = rho + clk[sys] + Pstd * rng.normal();
      add((cfg->variant & SWGN_SYNTH_SPP) ? SWGN_GNSS_SPP_PSEUDORANGE : SWGN_GNSS_RTK_PSEUDORANGE, b_pose + f, b_clk + 3 * e + sys, -1, P1,
          weight(var_p), var_p);
      // Doppler: r = istd (rate + drift + D1_lam)
      V3 ee = (1.0 / norm(xg - sp)) * (xg - sp);
      double rate = dot(vr - sat_vel[s], ee) +
                    kOmge / kClight * (sat_vel[s].y * xg.x + sp.y * vr.x - sat_vel[s].x * xg.y - sp.x * vr.y);
      double sinel = std::sin(sat_el[s]);
      double istd = sinel * sinel / 0.05;
      double D1_lam = -(rate + drift) + 0.05 * rng.normal();
      add(SWGN_GNSS_DOPPLER, b_sb + f, b_drift + e, b_pose + f, D1_lam, istd, 0.05 * 0.05);
      S->obs_amb.push_back(s);
      S->obs_sysfreq.push_back(sys * 2);
    }
    S->epoch_begin.push_back((int32_t)S->obs_amb.size());
    // InitialBlackFactor on `blackvalue`, once per GNSS frame (RVI/swf/swf_core.cpp:103-105)
    S->unit_block.push_back(b_black);
    S->unit_istd.push_back(1.0);
  }
  // FixedIntegerFactor (gnss_factor.cpp:85-96): r = istd ((N_a - N_ref) - N21), as LambdaSearch adds them after a fix
  if (!compA && (cfg->variant & SWGN_SYNTH_FIXED_INTEGER) && nsat > 0 && n_epochs_real > 0) {
    int ref_of_sys[3] = {-1, -1, -1};
    for (int s = 0; s < nsat; ++s) {
      const int sys = sat_sys[s];
      if (ref_of_sys[sys] < 0) {
        ref_of_sys[sys] = s;
        continue;
      }
      S->gnss_kind.push_back(SWGN_GNSS_FIXED_INTEGER);
      int32_t gb[3] = {b_N + ref_of_sys[sys], b_N + s, -1};
      S->gnss_blocks.insert(S->gnss_blocks.end(), gb, gb + 3);
      size_t o = S->gnss_data.size();
      S->gnss_data.resize(o + SWGN_GNSS_STRIDE, 0.0);
      double* r = S->gnss_data.data() + o;
      r[SWGN_GNSS_MEAS] = S->true_N[s] - S->true_N[ref_of_sys[sys]];
      r[SWGN_GNSS_WEIGHT] = 100.0;
    }
  }
  // InitialBlackFactor on `blackvalue2` (RVI/swf/swf_core.cpp:553-556)
  S->unit_block.push_back(b_black2);
  S->unit_istd.push_back(1.0);

  // ---- dense prior on (pose0, sb0, N...)  [MarginalizationFactor(last_marg_info)]
  {
    std::vector<int> blks = {b_pose + 0, b_sb + 0};
    for (int s = 0; s < nsat; ++s) blks.push_back(b_N + s);
    int n = 15 + nsat;
    std::vector<double> w(n);
    const bool gn = nsat > 0;
    for (int k = 0; k < 3; ++k) w[k] = gn ? 20.0 : 2e2;          // position
    for (int k = 3; k < 6; ++k) w[k] = gn ? 180 / kPi / 0.5 : 2e2;  // orientation
    for (int k = 6; k < 9; ++k) w[k] = gn ? 20.0 : 1e1;          // velocity
    for (int k = 9; k < 12; ++k) w[k] = gn ? 50.0 : 1e1;         // acc bias
    for (int k = 12; k < 15; ++k) w[k] = gn ? 500.0 : 1e2;       // gyro bias
    for (int k = 15; k < n; ++k) w[k] = 1.0 / 0.3;               // ambiguities, sigma 0.3 cycle
    S->prior_n.push_back(n);
    S->prior_blk_begin.push_back(0);
    int idx = 0;
    S->prior_x0_begin.push_back(0);
    for (int b : blks) {
      S->prior_blocks.push_back(b);
      S->prior_blk_idx.push_back(idx);
      idx += (S->block_size[b] == 7) ? 6 : S->block_size[b];
      const double* p = bp0(b);
      for (int k = 0; k < S->block_size[b]; ++k) S->prior_x0.push_back(p[k]);
    }
    S->prior_blk_begin.push_back((int32_t)blks.size());
    S->prior_J_begin.push_back(0);
    S->prior_r_begin.push_back(0);
    S->prior_J.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double c = (i == j) ? 1.0 : (j > i ? 0.05 * rng.normal() : 0.0);
        S->prior_J[(size_t)i * n + j] = c * w[j];
      }
    S->prior_r0.assign(n, 0.0);
    for (int i = 0; i < n; ++i) S->prior_r0[i] = 0.1 * rng.normal();
  }

  // ---- elimination ordering, RVI/swf/swf_gnss.cpp:629-783 (keep-blocks of the prior and the
  // parameter_head ambiguities are held back to the end)
  {
    int ors = 1;
    S->block_group[b_black2] = 0;
    for (int l = 0; l < cfg->n_landmarks; ++l) S->block_group[b_lm + l] = 0;
    for (int e = 0; e < n_epochs_real; ++e)
      for (int k = 0; k < 3; ++k) S->block_group[b_clk + 3 * e + k] = 0;
    int index = 0;
    auto in_problem = [&](int f) { return !(compA && frames[f].is_gnss); };
    for (int f = 1; f < F; ++f)  // sb0 is a keep-block of the prior
      if (in_problem(f) && index++ % 2 == 0) S->block_group[b_sb + f] = 0;
    for (int f = 1; f < F; ++f)
      if (in_problem(f) && S->block_group[b_sb + f] < 0) S->block_group[b_sb + f] = ors++;
    for (int f = 1; f < F; ++f)
      if (in_problem(f)) S->block_group[b_pose + f] = ors++;
    if (cfg->variant & SWGN_SYNTH_FREE_EXTRINSIC) S->block_group[b_ext] = ors++;  // after the poses (swf_gnss.cpp:710-717)
    if (n_epochs_real > 0) S->block_group[b_black] = ors++;
    for (int e = 0; e < n_epochs_real; ++e) S->block_group[b_drift + e] = ors++;
    S->block_group[b_pose + 0] = ors++;
    S->block_group[b_sb + 0] = ors++;
    for (int s = 0; s < nsat; ++s) S->block_group[b_N + s] = ors++;
    if (!(cfg->variant & SWGN_SYNTH_FREE_EXTRINSIC)) S->block_group[b_ext] = ors++;  // constant: removed by the reduced program anyway
  }
  // clocks that no satellite of that system touches would be e-blocks without rows: make them
  // constant (the reference only adds the slots it uses)
  {
    std::vector<char> touched(nb, 0);
    for (size_t i = 0; i < S->gnss_blocks.size(); ++i)
      if (S->gnss_blocks[i] >= 0) touched[S->gnss_blocks[i]] = 1;
    for (int e = 0; e < n_epochs_real; ++e)
      for (int k = 0; k < 3; ++k)
        if (!touched[b_clk + 3 * e + k]) S->block_const[b_clk + 3 * e + k] = 1;
  }

  // ---- graph struct
  swgn_graph& g = S->g;
  std::memset(&g, 0, sizeof(g));
  g.n_blocks = nb;
  g.block_size = S->block_size.data();
  g.block_manifold = S->block_manifold.data();
  g.block_const = S->block_const.data();
  g.block_group = S->block_group.data();
  g.block_offset = S->block_offset.data();
  g.n_state = off;
  g.state = S->state.data();
  g.Pbg[0] = Pbg.x; g.Pbg[1] = Pbg.y; g.Pbg[2] = Pbg.z;
  g.gravity[0] = g_w.x; g.gravity[1] = g_w.y; g.gravity[2] = g_w.z;
  g.proj_sqrt_info[0] = g.proj_sqrt_info[3] = 1000.0 / 1.5;  // FOCAL_LENGTH / FEATUREWEIGHTINVERSE
  g.proj_cauchy_a = 1.0;
  g.n_proj = (int32_t)(S->proj_uv.size() / 2);
  g.proj_blocks = S->proj_blocks.data();
  g.proj_uv = S->proj_uv.data();
  g.n_imu = (int32_t)(S->imu_blocks.size() / 4);
  g.imu_blocks = S->imu_blocks.data();
  g.imu_data = S->imu_data.data();
  g.n_gnss = (int32_t)S->gnss_kind.size();
  g.gnss_kind = S->gnss_kind.data();
  g.gnss_blocks = S->gnss_blocks.data();
  g.gnss_data = S->gnss_data.data();
  g.n_prior = 1;
  g.prior_n = S->prior_n.data();
  g.prior_blk_begin = S->prior_blk_begin.data();
  g.prior_blocks = S->prior_blocks.data();
  g.prior_blk_idx = S->prior_blk_idx.data();
  g.prior_x0_begin = S->prior_x0_begin.data();
  g.prior_x0 = S->prior_x0.data();
  g.prior_J_begin = S->prior_J_begin.data();
  g.prior_J = S->prior_J.data();
  g.prior_r_begin = S->prior_r_begin.data();
  g.prior_r0 = S->prior_r0.data();
  g.n_unit = (int32_t)S->unit_block.size();
  g.unit_block = S->unit_block.data();
  g.unit_istd = S->unit_istd.data();
  g.n_order = 0;
  g.order = nullptr;
  g.is_use = nullptr;
  g.n_chain = (int32_t)S->chain_blk_begin.size() - (S->chain_blk_begin.empty() ? 0 : 1);
  if (g.n_chain > 0) {
    g.chain_blk_begin = S->chain_blk_begin.data();
    g.chain_blocks = S->chain_blocks.data();
    g.chain_frame_begin = S->chain_frame_begin.data();
    g.chain_frame_data = S->chain_frame_data.data();
    g.chain_frame_N = S->chain_frame_N.data();
    g.chain_N = S->chain_N.data();
    g.chain_imu_data = S->chain_imu_data.data();
  }

  std::memset(&S->opt, 0, sizeof(S->opt));
  S->opt.max_num_iterations = 8;
  S->opt.max_num_consecutive_invalid_steps = 5;
  S->opt.initial_trust_region_radius = 1e4;
  S->opt.max_trust_region_radius = 1e16;
  S->opt.min_trust_region_radius = 1e-32;
  S->opt.min_relative_decrease = 1e-3;
  S->opt.min_lm_diagonal = 1e-6;
  S->opt.max_lm_diagonal = 1e32;
  S->opt.function_tolerance = 1e-6;
  S->opt.gradient_tolerance = 1e-10;
  S->opt.parameter_tolerance = 1e-8;
  S->opt.dogleg_min_mu = 1e-12;
  S->opt.is_optimize = 1;
  S->opt.n_parameter_head = nsat;

  S->info[0] = F;
  S->info[1] = g.n_proj;
  S->info[2] = g.n_imu;
  S->info[3] = g.n_gnss;
  S->info[4] = nsat;
  S->info[5] = b_N;
  S->info[6] = S->prior_n[0];
  S->info[7] = nb;
  return S;
}

void swgn_synth_destroy(swgn_synth* s) { delete s; }
const swgn_graph* swgn_synth_graph(const swgn_synth* s) { return &s->g; }
const double* swgn_synth_truth(const swgn_synth* s) { return s->truth.data(); }
void swgn_synth_options(const swgn_synth* s, swgn_options* o) { *o = s->opt; }
void swgn_synth_info(const swgn_synth* s, int32_t* info8) { std::memcpy(info8, s->info, sizeof(s->info)); }
int32_t swgn_synth_ambiguity_epochs(const swgn_synth* s, int32_t* epoch_begin, int32_t* obs_amb,
                                    int32_t* obs_sysfreq, int32_t* n_obs) {
  int32_t ne = (int32_t)s->epoch_begin.size() - 1;
  if (n_obs) *n_obs = (int32_t)s->obs_amb.size();
  if (epoch_begin) std::memcpy(epoch_begin, s->epoch_begin.data(), sizeof(int32_t) * (ne + 1));
  if (obs_amb) std::memcpy(obs_amb, s->obs_amb.data(), sizeof(int32_t) * s->obs_amb.size());
  if (obs_sysfreq) std::memcpy(obs_sysfreq, s->obs_sysfreq.data(), sizeof(int32_t) * s->obs_sysfreq.size());
  return ne;
}
const double* swgn_synth_true_ambiguities(const swgn_synth* s) { return s->true_N.data(); }
int32_t swgn_synth_preintegrate(int32_t n_samples, const double* samples7, const double* bias6, const double* noise4, double* record) {
  // the generator's own pre-integration (Preint above) on caller-provided samples: an implementation
  // independent of oracle/ and of the device kernel, used by the tests as a cross-check
  if (n_samples < 1) return 1;
  Preint pre(V3{samples7[1], samples7[2], samples7[3]}, V3{samples7[4], samples7[5], samples7[6]}, V3{bias6[0], bias6[1], bias6[2]},
             V3{bias6[3], bias6[4], bias6[5]}, noise4[0], noise4[1], noise4[2], noise4[3]);
  for (int s = 1; s < n_samples; ++s)
    pre.push(samples7[7 * s], V3{samples7[7 * s + 1], samples7[7 * s + 2], samples7[7 * s + 3]},
             V3{samples7[7 * s + 4], samples7[7 * s + 5], samples7[7 * s + 6]});
  std::memset(record, 0, sizeof(double) * SWGN_IMU_STRIDE);
  double* r = record;
  r[SWGN_IMU_DELTA_P] = pre.dp.x; r[SWGN_IMU_DELTA_P + 1] = pre.dp.y; r[SWGN_IMU_DELTA_P + 2] = pre.dp.z;
  r[SWGN_IMU_DELTA_Q] = pre.dq.x; r[SWGN_IMU_DELTA_Q + 1] = pre.dq.y;
  r[SWGN_IMU_DELTA_Q + 2] = pre.dq.z; r[SWGN_IMU_DELTA_Q + 3] = pre.dq.w;
  r[SWGN_IMU_DELTA_V] = pre.dv.x; r[SWGN_IMU_DELTA_V + 1] = pre.dv.y; r[SWGN_IMU_DELTA_V + 2] = pre.dv.z;
  r[SWGN_IMU_SUM_DT] = pre.sum_dt;
  std::memcpy(r + SWGN_IMU_JACOBIAN, pre.jac.data(), sizeof(double) * 225);
  Md sq;
  if (!pre.sqrt_info(&sq)) return 2;
  std::memcpy(r + SWGN_IMU_SQRT_INFO, sq.data(), sizeof(double) * 225);
  return 0;
}
int32_t swgn_synth_chain_frame_blocks(const swgn_synth* s, int32_t* pose_block, int32_t* sb_block) {
  const int32_t n = (int32_t)s->chain_frame_block.size();
  for (int32_t i = 0; i < n; ++i) {
    if (pose_block) pose_block[i] = s->chain_frame_block[i];
    if (sb_block) sb_block[i] = s->chain_frame_block[i] + s->info[0];  // speed-bias blocks follow the pose blocks
  }
  return n;
}
int32_t swgn_synth_chain_truth(const swgn_synth* s, double* frames16) {
  const int32_t n = (int32_t)(s->chain_frame_truth.size() / 16);
  if (frames16) std::memcpy(frames16, s->chain_frame_truth.data(), sizeof(double) * s->chain_frame_truth.size());
  return n;
}

}  // extern "C"
