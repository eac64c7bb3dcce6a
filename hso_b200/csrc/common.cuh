// Shared device helpers for the hso_b200 kernels (sm_100a). Product code: independent of oracle/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace hso {

#define HSO_DEV __device__ __forceinline__

// ---- small vector / quaternion SE3 (semantics of the reference's vendored Sophus, thirdparty/Sophus/sophus/se3.cpp,
// so3.cpp: quaternion-backed rotation that is re-normalised after every product) -------------------------------------
struct Quatd { double w, x, y, z; };
struct Se3d { Quatd q; double tx, ty, tz; };

HSO_DEV Quatd quat_mul(const Quatd& a, const Quatd& b) {
  Quatd r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
HSO_DEV void quat_normalize(Quatd& q) {
  const double inv = 1.0 / sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);  // one division (serial control path)
  q.w *= inv; q.x *= inv; q.y *= inv; q.z *= inv;
}
HSO_DEV void quat_rotate(const Quatd& q, double vx, double vy, double vz, double& ox, double& oy, double& oz) {
  double ux = q.y * vz - q.z * vy, uy = q.z * vx - q.x * vz, uz = q.x * vy - q.y * vx;
  ux += ux; uy += uy; uz += uz;
  ox = vx + q.w * ux + (q.y * uz - q.z * uy);
  oy = vy + q.w * uy + (q.z * ux - q.x * uz);
  oz = vz + q.w * uz + (q.x * uy - q.y * ux);
}
HSO_DEV void quat_to_R(const Quatd& q, double* R /*9 row-major*/) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
// Rotation matrix (row-major 3x4 [R|t]) -> quaternion, the standard trace/largest-diagonal branch method.
HSO_DEV Se3d se3_from_rt(const double* rt) {
  Se3d s;
  const double m00 = rt[0], m01 = rt[1], m02 = rt[2], m10 = rt[4], m11 = rt[5], m12 = rt[6], m20 = rt[8], m21 = rt[9], m22 = rt[10];
  double tr = m00 + m11 + m22;
  if (tr > 0) {
    double r = sqrt(tr + 1.0);
    s.q.w = 0.5 * r;
    r = 0.5 / r;
    s.q.x = (m21 - m12) * r; s.q.y = (m02 - m20) * r; s.q.z = (m10 - m01) * r;
  } else if (m00 >= m11 && m00 >= m22) {
    double r = sqrt(m00 - m11 - m22 + 1.0);
    s.q.x = 0.5 * r; r = 0.5 / r;
    s.q.w = (m21 - m12) * r; s.q.y = (m10 + m01) * r; s.q.z = (m20 + m02) * r;
  } else if (m11 >= m22) {
    double r = sqrt(m11 - m22 - m00 + 1.0);
    s.q.y = 0.5 * r; r = 0.5 / r;
    s.q.w = (m02 - m20) * r; s.q.z = (m21 + m12) * r; s.q.x = (m01 + m10) * r;
  } else {
    double r = sqrt(m22 - m00 - m11 + 1.0);
    s.q.z = 0.5 * r; r = 0.5 / r;
    s.q.w = (m10 - m01) * r; s.q.x = (m02 + m20) * r; s.q.y = (m12 + m21) * r;
  }
  quat_normalize(s.q);
  s.tx = rt[3]; s.ty = rt[7]; s.tz = rt[11];
  return s;
}
HSO_DEV void se3_to_rt(const Se3d& s, double* rt) {
  double R[9];
  quat_to_R(s.q, R);
  rt[0] = R[0]; rt[1] = R[1]; rt[2] = R[2];  rt[3] = s.tx;
  rt[4] = R[3]; rt[5] = R[4]; rt[6] = R[5];  rt[7] = s.ty;
  rt[8] = R[6]; rt[9] = R[7]; rt[10] = R[8]; rt[11] = s.tz;
}
HSO_DEV Se3d se3_mul(const Se3d& a, const Se3d& b) {
  Se3d r;
  double ox, oy, oz;
  quat_rotate(a.q, b.tx, b.ty, b.tz, ox, oy, oz);
  r.tx = a.tx + ox; r.ty = a.ty + oy; r.tz = a.tz + oz;
  r.q = quat_mul(a.q, b.q);
  quat_normalize(r.q);
  return r;
}
HSO_DEV Se3d se3_inverse(const Se3d& a) {
  Se3d r;
  r.q.w = a.q.w; r.q.x = -a.q.x; r.q.y = -a.q.y; r.q.z = -a.q.z;
  quat_normalize(r.q);
  double ox, oy, oz;
  quat_rotate(r.q, -a.tx, -a.ty, -a.tz, ox, oy, oz);
  r.tx = ox; r.ty = oy; r.tz = oz;
  return r;
}
// exp of a twist [upsilon(3), omega(3)] (thirdparty/Sophus/sophus/se3.cpp:170-196, so3.cpp:179-202).
HSO_DEV Se3d se3_exp(const double* u) {
  const double ox = u[3], oy = u[4], oz = u[5];
  const double theta = sqrt(ox * ox + oy * oy + oz * oz);
  const double half = 0.5 * theta;
  double imag;
  double sh, ch;
  sincos(half, &sh, &ch);  // one sincos: sin/cos(theta) below follow from the half angle (serial control path)
  const double real = ch;
  const double inv_theta = theta < 1e-10 ? 0.0 : 1.0 / theta;
  if (theta < 1e-10) {
    const double t2 = theta * theta, t4 = t2 * t2;
    imag = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
  } else {
    imag = sh * inv_theta;
  }
  Se3d r;
  r.q.w = real; r.q.x = imag * ox; r.q.y = imag * oy; r.q.z = imag * oz;
  quat_normalize(r.q);
  // V = I + A*Omega + B*Omega^2 ; Omega = hat(omega)
  double V[9];
  if (theta < 1e-10) {
    quat_to_R(r.q, V);
  } else {
    const double t2 = theta * theta;
    const double sn = 2.0 * sh * ch;          // sin(theta)
    const double one_m_cs = 2.0 * sh * sh;    // 1 - cos(theta), without cancellation
    const double inv_t2 = inv_theta * inv_theta;
    const double A = one_m_cs * inv_t2;
    const double B = (theta - sn) * inv_t2 * inv_theta;
    // Omega^2 = omega omega^T - |omega|^2 I
    V[0] = 1 + B * (ox * ox - t2);      V[1] = -A * oz + B * ox * oy;   V[2] = A * oy + B * ox * oz;
    V[3] = A * oz + B * ox * oy;        V[4] = 1 + B * (oy * oy - t2);  V[5] = -A * ox + B * oy * oz;
    V[6] = -A * oy + B * ox * oz;       V[7] = A * ox + B * oy * oz;    V[8] = 1 + B * (oz * oz - t2);
  }
  r.tx = V[0] * u[0] + V[1] * u[1] + V[2] * u[2];
  r.ty = V[3] * u[0] + V[4] * u[1] + V[5] * u[2];
  r.tz = V[6] * u[0] + V[7] * u[1] + V[8] * u[2];
  return r;
}

// ---- diagonal-pivoted LDL^T solve for tiny SPD systems (N = 6, 7), double. Matches the behaviour the reference gets
// from Eigen::LDLT on these systems: pivoting on the largest remaining diagonal, zero pivots give zero components. --------
template <int N>
HSO_DEV void ldlt_solve(const double* A /*NxN row-major symmetric*/, const double* b, double* x) {
  double m[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) m[i][j] = A[i * N + j];
  int perm[N];
  bool zero = false;
  for (int k = 0; k < N; ++k) {
    int big = k;
    double bv = fabs(m[k][k]);
    for (int i = k + 1; i < N; ++i) {
      double v = fabs(m[i][i]);
      if (v > bv) { bv = v; big = i; }
    }
    perm[k] = big;
    if (big != k) {
      for (int j = 0; j < k; ++j) { double t = m[k][j]; m[k][j] = m[big][j]; m[big][j] = t; }
      for (int i = big + 1; i < N; ++i) { double t = m[i][k]; m[i][k] = m[i][big]; m[i][big] = t; }
      for (int i = k + 1; i < big; ++i) { double t = m[i][k]; m[i][k] = m[big][i]; m[big][i] = t; }
      double t = m[k][k]; m[k][k] = m[big][big]; m[big][big] = t;
    }
    if (k > 0) {
      double tmp[N];
      double s = 0;
      for (int j = 0; j < k; ++j) { tmp[j] = m[j][j] * m[k][j]; s += m[k][j] * tmp[j]; }
      m[k][k] -= s;
      for (int i = k + 1; i < N; ++i) {
        double s2 = 0;
        for (int j = 0; j < k; ++j) s2 += m[i][j] * tmp[j];
        m[i][k] -= s2;
      }
    }
    const double akk = m[k][k];
    const bool valid = fabs(akk) > 0.0;
    if (k == 0 && !valid) { zero = true; break; }
    if (valid) for (int i = k + 1; i < N; ++i) m[i][k] /= akk;
  }
  if (zero) {
    for (int i = 0; i < N; ++i) x[i] = 0.0;
    return;
  }
  double y[N];
  for (int i = 0; i < N; ++i) y[i] = b[i];
  for (int k = 0; k < N; ++k) { double t = y[k]; y[k] = y[perm[k]]; y[perm[k]] = t; }
  for (int i = 0; i < N; ++i) for (int j = 0; j < i; ++j) y[i] -= m[i][j] * y[j];
  for (int i = 0; i < N; ++i) y[i] = (fabs(m[i][i]) > 2.2250738585072014e-308) ? y[i] / m[i][i] : 0.0;
  for (int i = N - 1; i >= 0; --i) for (int j = i + 1; j < N; ++j) y[i] -= m[j][i] * y[j];
  for (int k = N - 1; k >= 0; --k) { double t = y[k]; y[k] = y[perm[k]]; y[perm[k]] = t; }
  for (int i = 0; i < N; ++i) x[i] = y[i];
}

// Unpivoted LDL^T of a small SPD system entirely in registers (compile-time indices). Returns false when a pivot is not safely
// positive — the caller then takes the pivoted path above, which reproduces Eigen::LDLT's handling of semi-definite systems.
// On SPD input both give the same solution up to rounding (order cond(A) * 1e-16).
template <int N>
HSO_DEV bool ldlt_solve_spd_fast(const double* A /*NxN row-major symmetric*/, const double* b, double* x) {
  double L[N][N], d[N], dinv[N];
  double maxdiag = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) maxdiag = fmax(maxdiag, fabs(A[i * N + i]));
  const double tiny = 1e-11 * maxdiag;
  bool ok = isfinite(maxdiag) && maxdiag > 0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double dk = A[k * N + k];
#pragma unroll
    for (int j = 0; j < k; ++j) dk -= L[k][j] * L[k][j] * d[j];
    d[k] = dk;
    ok = ok && (dk > tiny);
    const double inv = 1.0 / dk;
    dinv[k] = inv;
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      double v = A[i * N + k];
#pragma unroll
      for (int j = 0; j < k; ++j) v -= L[i][j] * L[k][j] * d[j];
      L[i][k] = v * inv;
    }
  }
  if (!ok) return false;
  double y[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double v = b[i];
#pragma unroll
    for (int j = 0; j < i; ++j) v -= L[i][j] * y[j];
    y[i] = v;
  }
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] *= dinv[i];
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    double v = y[i];
#pragma unroll
    for (int j = i + 1; j < N; ++j) v -= L[j][i] * y[j];
    y[i] = v;
  }
  bool fin = true;
#pragma unroll
  for (int i = 0; i < N; ++i) { x[i] = y[i]; fin = fin && isfinite(y[i]); }
  return fin;
}

// ---- camera projection, double (src/camera.cpp:94-125,199-221,307-315) -------------------------------------------------
struct CamDev {
  int model, width, height, undistort;
  double fx, fy, cx, cy;
  double d[5];
  int distortion;  // |d0| > 1e-7 (pinhole only)
};

HSO_DEV void world2cam(const CamDev& c, double X, double Y, double Z, double& pu, double& pv) {
  const double zi = 1.0 / Z;  // one division; X/Z and Y/Z of src/camera.cpp:96 agree to 1 ulp (the pixel is rounded to float next)
  const double u = X * zi, v = Y * zi;
  if (c.model == 0 && c.distortion) {
    const double r2 = u * u + v * v, r4 = r2 * r2, r6 = r4 * r2;
    const double a1 = 2 * u * v, a2 = r2 + 2 * u * u, a3 = r2 + 2 * v * v;
    const double cdist = 1 + c.d[0] * r2 + c.d[1] * r4 + c.d[4] * r6;
    const double xd = u * cdist + c.d[2] * a1 + c.d[3] * a2;
    const double yd = v * cdist + c.d[2] * a3 + c.d[3] * a1;
    pu = xd * c.fx + c.cx;
    pv = yd * c.fy + c.cy;
  } else if (c.model == 1 && !c.undistort) {
    const double omega = c.d[0];
    const double dist = sqrt(u * u + v * v);
    const double ratio = (omega == 0 || dist == 0) ? 1.0 : atan(2 * dist * tan(omega / 2)) / (dist * omega);
    pu = ratio * c.fx * u + c.cx;
    pv = ratio * c.fy * v + c.cy;
  } else {
    pu = c.fx * u + c.cx;
    pv = c.fy * v + c.cy;
  }
}

// PinholeCamera/FOVCamera/EquidistantCamera::world2cam with the reference's true divisions (src/camera.cpp:94-125,199-221,307-315):
// the pixel is truncated to an integer for the in-frame test and the cell index, so no reciprocal shortcut here.
HSO_DEV void world2cam_exact(const CamDev& c, double X, double Y, double Z, double& pu, double& pv) {
  const double u = X / Z, v = Y / Z;
  if (c.model == 0 && c.distortion) {
    const double r2 = u * u + v * v, r4 = r2 * r2, r6 = r4 * r2;
    const double a1 = 2 * u * v, a2 = r2 + 2 * u * u, a3 = r2 + 2 * v * v;
    const double cdist = 1 + c.d[0] * r2 + c.d[1] * r4 + c.d[4] * r6;
    const double xd = u * cdist + c.d[2] * a1 + c.d[3] * a2;
    const double yd = v * cdist + c.d[2] * a3 + c.d[3] * a1;
    pu = xd * c.fx + c.cx;
    pv = yd * c.fy + c.cy;
  } else if (c.model == 1 && !c.undistort) {
    const double omega = c.d[0];
    const double dist = sqrt(u * u + v * v);
    const double ratio = (omega == 0 || dist == 0) ? 1.0 : atan(2 * dist * tan(omega / 2)) / (dist * omega);
    pu = ratio * c.fx * u + c.cx;
    pv = ratio * c.fy * v + c.cy;
  } else {
    pu = c.fx * u + c.cx;
    pv = c.fy * v + c.cy;
  }
}

// cam2world of the three models, unit-norm bearing (src/camera.cpp:66-87,169-190,297-300). The radtan branch is cv::undistortPoints on one
// CV_32FC2 point with the float camera matrix / coefficients the reference builds (camera.cpp:43-45): 5 fixed-point iterations in double,
// float in, float out.
HSO_DEV void cam2world(const CamDev& c, double u, double v, double& ox, double& oy, double& oz) {
  double x, y;
  if (c.model == 0 && c.distortion) {
    const float uf = (float)u, vf = (float)v;
    const double fx = (double)(float)c.fx, fy = (double)(float)c.fy, cx = (double)(float)c.cx, cy = (double)(float)c.cy;
    const double k0 = (double)(float)c.d[0], k1 = (double)(float)c.d[1], k2 = (double)(float)c.d[2], k3 = (double)(float)c.d[3],
                 k4 = (double)(float)c.d[4];
    const double ifx = 1. / fx, ify = 1. / fy;
    double xx = ((double)uf - cx) * ifx, yy = ((double)vf - cy) * ify;
    const double x0 = xx, y0 = yy;
    for (int j = 0; j < 5; ++j) {
      const double r2 = xx * xx + yy * yy;
      const double icdist = 1.0 / (1 + ((k4 * r2 + k1) * r2 + k0) * r2);
      if (icdist < 0) { xx = x0; yy = y0; break; }
      const double deltaX = 2 * k2 * xx * yy + k3 * (r2 + 2 * xx * xx);
      const double deltaY = k2 * (r2 + 2 * yy * yy) + 2 * k3 * xx * yy;
      xx = (x0 - deltaX) * icdist;
      yy = (y0 - deltaY) * icdist;
    }
    x = (double)(float)xx;
    y = (double)(float)yy;
  } else if (c.model == 1 && !c.undistort) {
    const double ud = (u - c.cx) / c.fx, vd = (v - c.cy) / c.fy;
    const double dist = sqrt(ud * ud + vd * vd);
    const double omega = c.d[0];
    const double rd = tan(dist * omega) / (2 * dist * tan(omega / 2));
    x = rd * ud;
    y = rd * vd;
  } else {
    x = (u - c.cx) / c.fx;
    y = (v - c.cy) / c.fy;
  }
  const double n = sqrt(x * x + y * y + 1.0);
  ox = x / n; oy = y / n; oz = 1.0 / n;
}

HSO_DEV void se3_apply(const Se3d& T, double x, double y, double z, double& ox, double& oy, double& oz) {
  quat_rotate(T.q, x, y, z, ox, oy, oz);
  ox += T.tx; oy += T.ty; oz += T.tz;
}


// ---- warp / block reductions -------------------------------------------------------------------------------------------
HSO_DEV double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
HSO_DEV float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
HSO_DEV int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- TMA 1-D bulk copy global -> shared with an mbarrier (SASS: UBLKCP) ------------------------------------------------
HSO_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
HSO_DEV void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
HSO_DEV void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
HSO_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
HSO_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
HSO_DEV bool mbar_try_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(phase)
      : "memory");
  return ok != 0;
}
HSO_DEV void mbar_wait(uint64_t* bar, uint32_t phase) {
  while (!mbar_try_wait(bar, phase)) {}
}
// bytes must be a multiple of 16; src/dst 16-byte aligned.
HSO_DEV void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace hso
