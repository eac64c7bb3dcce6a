// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product path
// (hso_b200/). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm use it.
//
// Small dependency-free math kit used by the CPU restatement of the HSO tracking hot path.
// It restates the *semantics* of the third-party types the reference uses:
//   - Sophus (vendored, non-templated) SO3/SE3:  /root/reference/thirdparty/Sophus/sophus/so3.cpp:64-85,179-202
//                                                /root/reference/thirdparty/Sophus/sophus/se3.cpp:59-95,170-196
//   - Eigen (system, NOT vendored, version unpinned: CMakeLists.txt:54): Quaterniond product /
//     normalize / _transformVector / toRotationMatrix, LDLT (diagonal-pivoted, robust Cholesky), and
//     fixed-size inverses. Eigen's published algorithms are restated; any SPD solve agrees to fp tolerance.
//   - hso::getMedian = nth_element at floor(n/2)  (include/hso/vikit/math_utils.h:119-126)
// Parity status of this file: SE3 is pinned by the known-answer cases of
// thirdparty/Sophus/sophus/test_se3.cpp:10-85 (tests/test_oracle_se3.py); LDLT is pinned against numpy.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

struct V2 { double x, y; };
struct V3 {
  double x, y, z;
  V3 operator+(const V3& o) const { return {x + o.x, y + o.y, z + o.z}; }
  V3 operator-(const V3& o) const { return {x - o.x, y - o.y, z - o.z}; }
  V3 operator*(double s) const { return {x * s, y * s, z * s}; }
  double dot(const V3& o) const { return x * o.x + y * o.y + z * o.z; }
  V3 cross(const V3& o) const { return {y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x}; }
  double norm() const { return std::sqrt(x * x + y * y + z * z); }
};

struct M3 {
  double m[3][3];
  static M3 identity() { M3 r{}; r.m[0][0] = r.m[1][1] = r.m[2][2] = 1.0; return r; }
  M3 operator*(const M3& o) const {
    M3 r{};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += m[i][k] * o.m[k][j];
        r.m[i][j] = s;
      }
    return r;
  }
  V3 operator*(const V3& v) const {
    return {m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z, m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
            m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z};
  }
  M3 operator+(const M3& o) const { M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = m[i][j] + o.m[i][j]; return r; }
  M3 operator*(double s) const { M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = m[i][j] * s; return r; }
};

// Unit quaternion, Eigen coefficient semantics (w + xi + yj + zk).
struct Quat {
  double w, x, y, z;
  static Quat identity() { return {1, 0, 0, 0}; }
  Quat mul(const Quat& b) const {  // Hamilton product, Eigen::Quaternion::operator*
    return {w * b.w - x * b.x - y * b.y - z * b.z, w * b.x + x * b.w + y * b.z - z * b.y,
            w * b.y + y * b.w + z * b.x - x * b.z, w * b.z + z * b.w + x * b.y - y * b.x};
  }
  void normalize() {
    double n = std::sqrt(w * w + x * x + y * y + z * z);
    w /= n; x /= n; y /= n; z /= n;
  }
  Quat conjugate() const { return {w, -x, -y, -z}; }
  V3 rotate(const V3& v) const {  // Eigen _transformVector: v + w*(2 q×v) + q×(2 q×v)
    V3 q{x, y, z};
    V3 uv = q.cross(v);
    uv = uv + uv;
    return v + uv * w + q.cross(uv);
  }
  M3 matrix() const {  // Eigen toRotationMatrix
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    M3 r;
    r.m[0][0] = 1 - (tyy + tzz); r.m[0][1] = txy - twz;       r.m[0][2] = txz + twy;
    r.m[1][0] = txy + twz;       r.m[1][1] = 1 - (txx + tzz); r.m[1][2] = tyz - twx;
    r.m[2][0] = txz - twy;       r.m[2][1] = tyz + twx;       r.m[2][2] = 1 - (txx + tyy);
    return r;
  }
};

static const double SMALL_EPS = 1e-10;  // thirdparty/Sophus/sophus/so3.h:35

inline M3 hat(const V3& v) {  // so3.cpp:204-212
  M3 o{};
  o.m[0][1] = -v.z; o.m[0][2] = v.y;
  o.m[1][0] = v.z;  o.m[1][2] = -v.x;
  o.m[2][0] = -v.y; o.m[2][1] = v.x;
  return o;
}

// SO3::expAndTheta, so3.cpp:179-202 (Taylor branch constants are the reference's truncated decimals).
inline Quat so3_exp(const V3& omega, double* theta_out) {
  double theta = omega.norm();
  double half_theta = 0.5 * theta;
  double imag_factor;
  double real_factor = std::cos(half_theta);
  if (theta < SMALL_EPS) {
    double theta_sq = theta * theta;
    double theta_po4 = theta_sq * theta_sq;
    imag_factor = 0.5 - 0.0208333 * theta_sq + 0.000260417 * theta_po4;
  } else {
    imag_factor = std::sin(half_theta) / theta;
  }
  Quat q{real_factor, imag_factor * omega.x, imag_factor * omega.y, imag_factor * omega.z};
  q.normalize();  // SO3(Quaterniond) ctor normalizes, so3.cpp:42-47
  *theta_out = theta;
  return q;
}

// SO3::logAndTheta, so3.cpp:124-170 (the |w|<eps branch is dead in the reference: its result is
// overwritten by the unconditional atan line; restated as such).
inline V3 so3_log(const Quat& q, double* theta_out) {
  double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
  double w = q.w;
  double two_atan_nbyw_by_n;
  if (n < SMALL_EPS) {
    two_atan_nbyw_by_n = 2. / w - 2. * (n * n) / (w * w * w);
  } else {
    two_atan_nbyw_by_n = 2 * std::atan(n / w) / n;
  }
  *theta_out = two_atan_nbyw_by_n * n;
  return V3{q.x, q.y, q.z} * two_atan_nbyw_by_n;
}

struct SE3 {
  Quat q = Quat::identity();
  V3 t{0, 0, 0};

  SE3 mul(const SE3& o) const {  // se3.cpp:59-66: translation += R*other.t ; q *= other.q ; normalize
    SE3 r;
    r.t = t + q.rotate(o.t);
    r.q = q.mul(o.q);
    r.q.normalize();
    return r;
  }
  SE3 inverse() const {  // se3.cpp:78-85 ; SO3::inverse goes through SO3(Quaterniond) → normalize
    SE3 r;
    r.q = q.conjugate();
    r.q.normalize();
    r.t = r.q.rotate(t * -1.);
    return r;
  }
  V3 apply(const V3& p) const { return q.rotate(p) + t; }  // se3.cpp:93-97
  M3 rotation() const { return q.matrix(); }

  // SE3::exp, se3.cpp:170-196 ; tangent order [upsilon(3), omega(3)].
  static SE3 exp(const double u[6]) {
    V3 upsilon{u[0], u[1], u[2]};
    V3 omega{u[3], u[4], u[5]};
    double theta;
    SE3 r;
    r.q = so3_exp(omega, &theta);
    M3 Omega = hat(omega);
    M3 Omega_sq = Omega * Omega;
    M3 V;
    if (theta < SMALL_EPS) {
      V = r.q.matrix();
    } else {
      double theta_sq = theta * theta;
      V = M3::identity() + Omega * ((1 - std::cos(theta)) / theta_sq) + Omega_sq * ((theta - std::sin(theta)) / (theta_sq * theta));
    }
    r.t = V * upsilon;
    return r;
  }

  // SE3::log, se3.cpp:198-221.
  void log(double out[6]) const {
    double theta;
    V3 om = so3_log(q, &theta);
    M3 Omega = hat(om);
    M3 V_inv;
    if (theta < SMALL_EPS) {
      V_inv = M3::identity() + Omega * -0.5 + (Omega * Omega) * (1. / 12.);
    } else {
      V_inv = M3::identity() + Omega * -0.5 + (Omega * Omega) * ((1 - theta / (2 * std::tan(theta / 2))) / (theta * theta));
    }
    V3 up = V_inv * t;
    out[0] = up.x; out[1] = up.y; out[2] = up.z;
    out[3] = om.x; out[4] = om.y; out[5] = om.z;
  }

  // 12-double row-major [R | t] (3x4) — the layout the C-ABI uses for poses.
  static SE3 from_rt(const double* rt) {
    // Eigen Quaternion(Matrix3) conversion (Shepperd-style, Eigen/src/Geometry/Quaternion.h).
    const double m00 = rt[0], m01 = rt[1], m02 = rt[2], m10 = rt[4], m11 = rt[5], m12 = rt[6], m20 = rt[8], m21 = rt[9], m22 = rt[10];
    Quat q;
    double tr = m00 + m11 + m22;
    if (tr > 0) {
      double s = std::sqrt(tr + 1.0);
      q.w = 0.5 * s;
      s = 0.5 / s;
      q.x = (m21 - m12) * s; q.y = (m02 - m20) * s; q.z = (m10 - m01) * s;
    } else {
      const double mm[3][3] = {{m00, m01, m02}, {m10, m11, m12}, {m20, m21, m22}};
      int i = 0;
      if (m11 > m00) i = 1;
      if (m22 > mm[i][i]) i = 2;
      int j = (i + 1) % 3, k = (j + 1) % 3;
      double s = std::sqrt(mm[i][i] - mm[j][j] - mm[k][k] + 1.0);
      double v[3];
      v[i] = 0.5 * s;
      s = 0.5 / s;
      q.w = (mm[k][j] - mm[j][k]) * s;
      v[j] = (mm[j][i] + mm[i][j]) * s;
      v[k] = (mm[k][i] + mm[i][k]) * s;
      q.x = v[0]; q.y = v[1]; q.z = v[2];
    }
    q.normalize();
    SE3 r;
    r.q = q;
    r.t = {rt[3], rt[7], rt[11]};
    return r;
  }
  void to_rt(double* rt) const {
    M3 R = q.matrix();
    for (int i = 0; i < 3; ++i) { rt[4 * i + 0] = R.m[i][0]; rt[4 * i + 1] = R.m[i][1]; rt[4 * i + 2] = R.m[i][2]; }
    rt[3] = t.x; rt[7] = t.y; rt[11] = t.z;
  }
};

// Diagonal-pivoted LDL^T ("robust Cholesky") solve, restating the published algorithm of Eigen::LDLT
// (Eigen/src/Cholesky/LDLT.h: unblocked in-place factorisation with largest-|diagonal| pivoting; solve =
// P^T L^-T D^+ L^-1 P b where D^+ zeroes entries with |d| <= min positive double). N <= 8.
template <int N>
inline void ldlt_solve(const double* A /*NxN row-major, symmetric*/, const double* b, double* x) {
  double m[N][N];
  for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) m[i][j] = A[i * N + j];
  int tr[N];
  bool all_zero = false;
  for (int k = 0; k < N; ++k) {
    int big = k;
    double bigv = std::fabs(m[k][k]);
    for (int i = k + 1; i < N; ++i) if (std::fabs(m[i][i]) > bigv) { bigv = std::fabs(m[i][i]); big = i; }
    tr[k] = big;
    if (big != k) {  // symmetric row/col swap on the lower triangle
      int s = N - big - 1;
      for (int j = 0; j < k; ++j) std::swap(m[k][j], m[big][j]);
      for (int i = 0; i < s; ++i) std::swap(m[big + 1 + i][k], m[big + 1 + i][big]);
      for (int i = k + 1; i < big; ++i) std::swap(m[i][k], m[big][i]);
      std::swap(m[k][k], m[big][big]);
    }
    int rs = N - k - 1;
    if (k > 0) {
      double temp[N];
      for (int j = 0; j < k; ++j) temp[j] = m[j][j] * m[k][j];
      double s = 0;
      for (int j = 0; j < k; ++j) s += m[k][j] * temp[j];
      m[k][k] -= s;
      for (int i = 0; i < rs; ++i) {
        double s2 = 0;
        for (int j = 0; j < k; ++j) s2 += m[k + 1 + i][j] * temp[j];
        m[k + 1 + i][k] -= s2;
      }
    }
    double akk = m[k][k];
    bool valid = std::fabs(akk) > 0.0;
    if (k == 0 && !valid) {
      for (int j = 0; j < N; ++j) tr[j] = j;
      all_zero = true;
      break;
    }
    if (rs > 0 && valid) for (int i = 0; i < rs; ++i) m[k + 1 + i][k] /= akk;
  }
  double y[N];
  for (int i = 0; i < N; ++i) y[i] = b[i];
  if (all_zero) {
    for (int i = 0; i < N; ++i) x[i] = 0.0;  // D^+ = 0
    return;
  }
  for (int k = 0; k < N; ++k) std::swap(y[k], y[tr[k]]);         // P b
  for (int i = 0; i < N; ++i) for (int j = 0; j < i; ++j) y[i] -= m[i][j] * y[j];  // L^-1
  const double tol = 2.2250738585072014e-308;                                     // (numeric_limits<double>::min)()
  for (int i = 0; i < N; ++i) { if (std::fabs(m[i][i]) > tol) y[i] /= m[i][i]; else y[i] = 0.0; }
  for (int i = N - 1; i >= 0; --i) for (int j = i + 1; j < N; ++j) y[i] -= m[j][i] * y[j];  // L^-T
  for (int k = N - 1; k >= 0; --k) std::swap(y[k], y[tr[k]]);   // P^T
  for (int i = 0; i < N; ++i) x[i] = y[i];
}

// hso::getMedian (include/hso/vikit/math_utils.h:119-126): nth_element at floor(n/2) — upper median.
template <class T>
inline T median_inplace(std::vector<T>& v) {
  auto it = v.begin() + (v.size() / 2);
  std::nth_element(v.begin(), it, v.end());
  return *it;
}

// Closed-form inverses (Eigen fixed-size inverse() uses cofactors for 2x2/3x3).
inline void inv2f(const float* H, float* Hi) {
  float det = H[0] * H[3] - H[1] * H[2];
  float id = 1.0f / det;
  Hi[0] = H[3] * id; Hi[1] = -H[1] * id; Hi[2] = -H[2] * id; Hi[3] = H[0] * id;
}
inline void inv3f(const float* H, float* Hi) {
  float c00 = H[4] * H[8] - H[5] * H[7];
  float c10 = H[5] * H[6] - H[3] * H[8];
  float c20 = H[3] * H[7] - H[4] * H[6];
  float det = H[0] * c00 + H[1] * c10 + H[2] * c20;
  float id = 1.0f / det;
  Hi[0] = c00 * id; Hi[1] = (H[2] * H[7] - H[1] * H[8]) * id; Hi[2] = (H[1] * H[5] - H[2] * H[4]) * id;
  Hi[3] = c10 * id; Hi[4] = (H[0] * H[8] - H[2] * H[6]) * id; Hi[5] = (H[2] * H[3] - H[0] * H[5]) * id;
  Hi[6] = c20 * id; Hi[7] = (H[1] * H[6] - H[0] * H[7]) * id; Hi[8] = (H[0] * H[4] - H[1] * H[3]) * id;
}

}  // namespace orc
