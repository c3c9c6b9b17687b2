// oracle/geom.h -- TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into the product.
//
// Eigen-free restatement of the rotation / rigid-motion arithmetic that sits on the reference's
// hot path.  Every function cites the reference (or third-party) code it follows.  Third-party
// formulas that are NOT under /root/reference are marked [ext-knowledge]: Eigen 3.3 (unpinned,
// CMakeLists.txt:56) and Ceres 1.13 (unpinned, CMakeLists.txt:67) -- "parity unpinned" for these.
//
// Conventions: 3x3 matrices are row-major T[9] (m[3*i+j] = M(i,j)); quaternions are Eigen
// coefficient order xyzw (se3.hpp:917, eigen_quaternion.h:108-114); T may be double or orc::Jet<N>.
#pragma once
#include <cmath>
#include <limits>
#include "jet.h"

namespace orc {

// ---- tiny helpers -----------------------------------------------------------------------------
template <typename T> inline void mat3_mul_vec(const T* m, const T* v, T* out) {
  // Eigen fixed-size Matrix3*Vector3 coefficient product, left fold (SURVEY A2 op order).
  T r0 = (m[0] * v[0] + m[1] * v[1]) + m[2] * v[2];
  T r1 = (m[3] * v[0] + m[4] * v[1]) + m[5] * v[2];
  T r2 = (m[6] * v[0] + m[7] * v[1]) + m[8] * v[2];
  out[0] = r0; out[1] = r1; out[2] = r2;
}
template <typename T> inline void mat3_mul(const T* a, const T* b, T* out) {
  T r[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      r[3 * i + j] = (a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j]) + a[3 * i + 2] * b[6 + j];
  for (int i = 0; i < 9; ++i) out[i] = r[i];
}
template <typename T> inline void cross3(const T* a, const T* b, T* out) {
  T r0 = a[1] * b[2] - a[2] * b[1];
  T r1 = a[2] * b[0] - a[0] * b[2];
  T r2 = a[0] * b[1] - a[1] * b[0];
  out[0] = r0; out[1] = r1; out[2] = r2;
}

// General 3x3 inverse by cofactors: what `dstCloud.pose.linear().inverse()` evaluates to
// (src/internal/frame.cpp:118; Eigen 3.3 compute_inverse_size3_helper [ext-knowledge]).
inline void mat3_inverse_cofactor(const double* m, double* inv) {
  auto M = [&](int i, int j) { return m[3 * i + j]; };
  auto cof = [&](int i, int j) {  // cofactor_3x3<i,j>
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return M(i1, j1) * M(i2, j2) - M(i1, j2) * M(i2, j1);
  };
  const double c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
  const double det = (c00 * M(0, 0) + c10 * M(1, 0)) + c20 * M(2, 0);
  const double invdet = 1.0 / det;
  // result(r,c) = cofactor(c,r) * invdet
  inv[0] = c00 * invdet; inv[1] = c10 * invdet; inv[2] = c20 * invdet;
  inv[3] = cof(0, 1) * invdet; inv[4] = cof(1, 1) * invdet; inv[5] = cof(2, 1) * invdet;
  inv[6] = cof(0, 2) * invdet; inv[7] = cof(1, 2) * invdet; inv[8] = cof(2, 2) * invdet;
}

// ---- Eigen quaternion arithmetic [ext-knowledge Eigen 3.3 Geometry/Quaternion.h] -------------
// Quaterniond(Matrix3d): used by icp-ceres.cpp:236 and by Sophus SO3Group(R) (so3.hpp:666-668).
// NOT normalised afterwards.
inline void quat_from_matrix(const double* m, double* q /*xyzw*/) {
  auto M = [&](int i, int j) { return m[3 * i + j]; };
  double t = (M(0, 0) + M(1, 1)) + M(2, 2);
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (M(2, 1) - M(1, 2)) * t;
    q[1] = (M(0, 2) - M(2, 0)) * t;
    q[2] = (M(1, 0) - M(0, 1)) * t;
  } else {
    int i = 0;
    if (M(1, 1) > M(0, 0)) i = 1;
    if (M(2, 2) > M(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(M(i, i) - M(j, j) - M(k, k) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (M(k, j) - M(j, k)) * t;
    q[j] = (M(j, i) + M(i, j)) * t;
    q[k] = (M(k, i) + M(i, k)) * t;
  }
}

// Quaternion::_transformVector (what `q * v` means in icp-ceres.h:80,128,261,307 and so3.hpp:293-295).
// Assumes unit q; for non-unit q it is still this (linear-in-v) map, which is what the reference runs.
template <typename T> inline void quat_transform(const T* q, const T* v, T* out) {
  T uv[3]; cross3(q, v, uv);
  uv[0] = uv[0] + uv[0]; uv[1] = uv[1] + uv[1]; uv[2] = uv[2] + uv[2];
  T c[3]; cross3(q, uv, c);
  T r0 = v[0] + q[3] * uv[0] + c[0];
  T r1 = v[1] + q[3] * uv[1] + c[1];
  T r2 = v[2] + q[3] * uv[2] + c[2];
  out[0] = r0; out[1] = r1; out[2] = r2;
}

// Quaternion::toRotationMatrix (icp-ceres.h:132-134, icp-ceres.cpp:119, se3.hpp rotationMatrix()).
template <typename T> inline void quat_to_matrix(const T* q, T* m) {
  const T tx = T(2.0) * q[0], ty = T(2.0) * q[1], tz = T(2.0) * q[2];
  const T twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const T txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const T tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  m[0] = T(1.0) - (tyy + tzz); m[1] = txy - twz;            m[2] = txz + twy;
  m[3] = txy + twz;            m[4] = T(1.0) - (txx + tzz); m[5] = tyz - twx;
  m[6] = txz - twy;            m[7] = tyz + twx;            m[8] = T(1.0) - (txx + tyy);
}

// Quaternion product a*b (Hamilton), Eigen xyzw storage.
template <typename T> inline void quat_mul(const T* a, const T* b, T* out) {
  T w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  T x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  T y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  T z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  out[0] = x; out[1] = y; out[2] = z; out[3] = w;
}

// ---- Sophus SO3 / SE3 (ext/sophus-ceres/sophus) -----------------------------------------------
inline constexpr double kSophusEps = 1e-10;  // sophus.hpp:37-41

// SO3Group::normalize (so3.hpp:234-240)
template <typename T> inline void so3_normalize(T* q) {
  T len = sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
  q[0] = q[0] / len; q[1] = q[1] / len; q[2] = q[2] / len; q[3] = q[3] / len;
}

// SO3Group::expAndTheta (so3.hpp:382-408). Quaternion constructed from (real, imag*omega) and then
// normalised by the explicit SO3Group(Quaternion) constructor (so3.hpp:675-678).
template <typename T> inline void so3_exp_and_theta(const T* omega, T* q, T* theta) {
  T theta_sq = (omega[0] * omega[0] + omega[1] * omega[1]) + omega[2] * omega[2];
  *theta = sqrt(theta_sq);  // derivative parts are non-finite at 0 for Jets -- unused in that branch
  T half_theta = T(0.5) * (*theta);
  T imag_factor, real_factor;
  if ((*theta) < kSophusEps) {
    T theta_po4 = theta_sq * theta_sq;
    imag_factor = T(0.5) - T(1.0 / 48.0) * theta_sq + T(1.0 / 3840.0) * theta_po4;
    real_factor = T(1.0) - T(0.5) * theta_sq + T(1.0 / 384.0) * theta_po4;
  } else {
    T sin_half_theta = sin(half_theta);
    imag_factor = sin_half_theta / (*theta);
    real_factor = cos(half_theta);
  }
  q[0] = imag_factor * omega[0]; q[1] = imag_factor * omega[1]; q[2] = imag_factor * omega[2];
  q[3] = real_factor;
  so3_normalize(q);
}

// 7-parameter SE3 element, layout [qx qy qz qw tx ty tz] (se3.hpp:108-111,917).
// SE3Group::exp (se3.hpp:468-488); tangent = (upsilon, omega).
template <typename T> inline void se3_exp(const T* a, T* out7) {
  const T* ups = a; const T* omega = a + 3;
  T theta; T q[4];
  so3_exp_and_theta(omega, q, &theta);
  T Om[9] = {T(0.0), -omega[2], omega[1], omega[2], T(0.0), -omega[0], -omega[1], omega[0], T(0.0)};
  T V[9];
  if (theta < kSophusEps) {
    quat_to_matrix(q, V);  // V = so3.matrix()
  } else {
    T Om2[9]; mat3_mul(Om, Om, Om2);
    T theta_sq = theta * theta;
    T c1 = (T(1.0) - cos(theta)) / theta_sq;
    T c2 = (theta - sin(theta)) / (theta_sq * theta);
    for (int i = 0; i < 9; ++i) V[i] = c1 * Om[i] + c2 * Om2[i];
    V[0] = V[0] + T(1.0); V[4] = V[4] + T(1.0); V[8] = V[8] + T(1.0);
  }
  T t[3]; mat3_mul_vec(V, ups, t);
  out7[0] = q[0]; out7[1] = q[1]; out7[2] = q[2]; out7[3] = q[3];
  out7[4] = t[0]; out7[5] = t[1]; out7[6] = t[2];
}

// SE3 operator* = fastMultiply + normalize (se3.hpp:169-173, 288-321).
template <typename T> inline void se3_mul(const T* a7, const T* b7, T* out7) {
  T rt[3]; quat_transform(a7, b7 + 4, rt);
  T t0 = a7[4] + rt[0], t1 = a7[5] + rt[1], t2 = a7[6] + rt[2];
  T q[4]; quat_mul(a7, b7, q);
  so3_normalize(q);
  out7[0] = q[0]; out7[1] = q[1]; out7[2] = q[2]; out7[3] = q[3];
  out7[4] = t0; out7[5] = t1; out7[6] = t2;
}

// SophusSE3Plus / LocalParameterizationSE3::Plus (sophus_se3.h:10-19, 31-38): x * exp(delta).
template <typename T> inline void se3_plus(const T* x7, const T* delta6, T* out7) {
  T e[7]; se3_exp(delta6, e);
  se3_mul(x7, e, out7);
}

// LocalParameterizationSE3::ComputeJacobian (sophus_se3.h:45-51) = internalJacobian()^T
// (se3.hpp:183-212, generators 553-571).  Returned as 7x6, J[7*?]: jac[r*6+c] = d x_r / d delta_c.
inline void se3_internal_jacobian(const double* x7, double* jac /*7x6 row-major*/) {
  for (int c = 0; c < 6; ++c) {
    double gq[4] = {0, 0, 0, 0}, gt[3] = {0, 0, 0};
    if (c < 3) gt[c] = 1.0; else gq[c - 3] = 0.5;
    double rq[4]; quat_mul(x7, gq, rq);
    double rt[3]; quat_transform(x7, gt, rt);
    for (int r = 0; r < 4; ++r) jac[r * 6 + c] = rq[r];
    for (int r = 0; r < 3; ++r) jac[(4 + r) * 6 + c] = rt[r];
  }
}

// ---- Eigen-quaternion local parameterisation (include/eigen_quaternion.h:89-117) --------------
inline void eigen_quat_plus(const double* x /*xyzw*/, const double* delta3, double* out) {
  const double n = std::sqrt((delta3[0] * delta3[0] + delta3[1] * delta3[1]) + delta3[2] * delta3[2]);
  if (n > 0.0) {
    const double s = std::sin(n) / n;
    const double tmp[4] = {s * delta3[0], s * delta3[1], s * delta3[2], std::cos(n)};
    quat_mul(tmp, x, out);
  } else {
    for (int i = 0; i < 4; ++i) out[i] = x[i];
  }
}
inline void eigen_quat_jacobian(const double* x, double* jac /*4x3 row-major, rows xyzw*/) {
  jac[0] = x[3];  jac[1] = x[2];   jac[2] = -x[1];
  jac[3] = -x[2]; jac[4] = x[3];   jac[5] = x[0];
  jac[6] = x[1];  jac[7] = -x[0];  jac[8] = x[3];
  jac[9] = -x[0]; jac[10] = -x[1]; jac[11] = -x[2];
}

// ---- Ceres rotation.h [ext-knowledge, Ceres 1.13] ---------------------------------------------
// ceres::AngleAxisRotatePoint (icp-ceres.h:162,165,...).  In-place safe, as the reference calls it.
template <typename T> inline void angle_axis_rotate_point(const T* aa, const T* pt, T* result) {
  const T theta2 = (aa[0] * aa[0] + aa[1] * aa[1]) + aa[2] * aa[2];
  if (theta2 > std::numeric_limits<double>::epsilon()) {
    const T theta = sqrt(theta2);
    const T costheta = cos(theta);
    const T sintheta = sin(theta);
    const T theta_inverse = T(1.0) / theta;
    const T w[3] = {aa[0] * theta_inverse, aa[1] * theta_inverse, aa[2] * theta_inverse};
    const T wxp[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = ((w[0] * pt[0] + w[1] * pt[1]) + w[2] * pt[2]) * (T(1.0) - costheta);
    T r0 = pt[0] * costheta + wxp[0] * sintheta + w[0] * tmp;
    T r1 = pt[1] * costheta + wxp[1] * sintheta + w[1] * tmp;
    T r2 = pt[2] * costheta + wxp[2] * sintheta + w[2] * tmp;
    result[0] = r0; result[1] = r1; result[2] = r2;
  } else {
    const T wxp[3] = {aa[1] * pt[2] - aa[2] * pt[1], aa[2] * pt[0] - aa[0] * pt[2], aa[0] * pt[1] - aa[1] * pt[0]};
    T r0 = pt[0] + wxp[0], r1 = pt[1] + wxp[1], r2 = pt[2] + wxp[2];
    result[0] = r0; result[1] = r1; result[2] = r2;
  }
}

// ceres::RotationMatrixToAngleAxis via quaternion (icp-ceres.cpp:101, isoToAngleAxis).
inline void rotation_matrix_to_angle_axis(const double* m, double* aa) {
  auto R = [&](int i, int j) { return m[3 * i + j]; };
  double q[4];  // wxyz (Ceres order)
  const double trace = R(0, 0) + R(1, 1) + R(2, 2);
  if (trace >= 0.0) {
    double t = std::sqrt(trace + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (R(2, 1) - R(1, 2)) * t;
    q[2] = (R(0, 2) - R(2, 0)) * t;
    q[3] = (R(1, 0) - R(0, 1)) * t;
  } else {
    int i = 0;
    if (R(1, 1) > R(0, 0)) i = 1;
    if (R(2, 2) > R(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0);
    q[i + 1] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R(k, j) - R(j, k)) * t;
    q[j + 1] = (R(j, i) + R(i, j)) * t;
    q[k + 1] = (R(k, i) + R(i, k)) * t;
  }
  const double q1 = q[1], q2 = q[2], q3 = q[3];
  const double sin_squared_theta = q1 * q1 + q2 * q2 + q3 * q3;
  if (sin_squared_theta > 0.0) {
    const double sin_theta = std::sqrt(sin_squared_theta);
    const double cos_theta = q[0];
    const double two_theta = 2.0 * ((cos_theta < 0.0) ? std::atan2(-sin_theta, -cos_theta)
                                                       : std::atan2(sin_theta, cos_theta));
    const double k = two_theta / sin_theta;
    aa[0] = q1 * k; aa[1] = q2 * k; aa[2] = q3 * k;
  } else {
    aa[0] = q1 * 2.0; aa[1] = q2 * 2.0; aa[2] = q3 * 2.0;
  }
}

// ceres::AngleAxisToRotationMatrix (icp-ceres.cpp:111, axisAngleToIso).
inline void angle_axis_to_rotation_matrix(const double* aa, double* m) {
  auto R = [&](int i, int j) -> double& { return m[3 * i + j]; };
  const double theta2 = (aa[0] * aa[0] + aa[1] * aa[1]) + aa[2] * aa[2];
  if (theta2 > std::numeric_limits<double>::epsilon()) {
    const double theta = std::sqrt(theta2);
    const double wx = aa[0] / theta, wy = aa[1] / theta, wz = aa[2] / theta;
    const double costheta = std::cos(theta), sintheta = std::sin(theta);
    R(0, 0) = costheta + wx * wx * (1.0 - costheta);
    R(1, 0) = wz * sintheta + wx * wy * (1.0 - costheta);
    R(2, 0) = -wy * sintheta + wx * wz * (1.0 - costheta);
    R(0, 1) = wx * wy * (1.0 - costheta) - wz * sintheta;
    R(1, 1) = costheta + wy * wy * (1.0 - costheta);
    R(2, 1) = wx * sintheta + wy * wz * (1.0 - costheta);
    R(0, 2) = wy * sintheta + wx * wz * (1.0 - costheta);
    R(1, 2) = -wx * sintheta + wy * wz * (1.0 - costheta);
    R(2, 2) = costheta + wz * wz * (1.0 - costheta);
  } else {
    R(0, 0) = 1.0;    R(1, 0) = aa[2];  R(2, 0) = -aa[1];
    R(0, 1) = -aa[2]; R(1, 1) = 1.0;    R(2, 1) = aa[0];
    R(0, 2) = aa[1];  R(1, 2) = -aa[0]; R(2, 2) = 1.0;
  }
}

// ---- pose16 (Isometry3d = 4x4 column-major double[16]) <-> pieces -----------------------------
inline void pose16_split(const double* P, double* R /*row-major 3x3*/, double* t) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R[3 * i + j] = P[4 * j + i];
    t[i] = P[12 + i];
  }
}
inline void pose16_join(const double* R, const double* t, double* P) {
  for (int i = 0; i < 16; ++i) P[i] = 0.0;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) P[4 * j + i] = R[3 * i + j];
    P[12 + i] = t[i];
  }
  P[15] = 1.0;
}

}  // namespace orc
