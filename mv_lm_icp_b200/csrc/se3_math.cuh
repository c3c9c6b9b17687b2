// se3_math.cuh -- rigid-motion arithmetic of the engine (host + device).
//
// Pose <-> parameter conversions and the Plus operators of the three parameterisations the
// reference offers:
//   AA   ceres RotationMatrixToAngleAxis / AngleAxisToRotationMatrix   icp-ceres.cpp:97-115
//   QUAT Eigen Quaterniond(R) / toRotationMatrix + EigenQuaternionParameterization::Plus
//                                                  icp-ceres.cpp:117-122,236; eigen_quaternion.h:89-106
//   SE3  Sophus::SE3d(pose) / rotationMatrix + x*exp(delta)   icp-ceres.cpp:124-134; sophus_se3.h:31-38;
//                                                  se3.hpp:468-488,288-321; so3.hpp:382-408
// Row-major 3x3 (m[3*i+j]); quaternions xyzw.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define MV_HD __host__ __device__ __forceinline__
#else
#define MV_HD inline
#endif

namespace mv {

struct Rt { double R[9]; double t[3]; };   // frame -> world: y = R p + t

MV_HD void cross(const double* a, const double* b, double* c) {
  const double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  c[0] = x; c[1] = y; c[2] = z;
}
MV_HD void matvec(const double* m, const double* v, double* o) {
  const double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  const double y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  const double z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
MV_HD void matTvec(const double* m, const double* v, double* o) {
  const double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  const double y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  const double z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
MV_HD void matmul(const double* a, const double* b, double* o) {   // o = a b (o distinct from a,b)
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
MV_HD void matTmul(const double* a, const double* b, double* o) {  // o = a^T b
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}

// ---- quaternion (xyzw) -------------------------------------------------------------------------
MV_HD void quat_of_matrix(const double* m, double* q) {   // Eigen's branchy formula, not normalised
  const double tr = m[0] + m[4] + m[8];
  if (tr > 0.0) {
    double s = sqrt(tr + 1.0);
    q[3] = 0.5 * s; s = 0.5 / s;
    q[0] = (m[7] - m[5]) * s; q[1] = (m[2] - m[6]) * s; q[2] = (m[3] - m[1]) * s;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    q[i] = 0.5 * s; s = 0.5 / s;
    q[3] = (m[3 * k + j] - m[3 * j + k]) * s;
    q[j] = (m[3 * j + i] + m[3 * i + j]) * s;
    q[k] = (m[3 * k + i] + m[3 * i + k]) * s;
  }
}
MV_HD void matrix_of_quat(const double* q, double* m) {
  const double x2 = 2.0 * q[0], y2 = 2.0 * q[1], z2 = 2.0 * q[2];
  const double wx = x2 * q[3], wy = y2 * q[3], wz = z2 * q[3];
  const double xx = x2 * q[0], xy = y2 * q[0], xz = z2 * q[0], yy = y2 * q[1], yz = z2 * q[1], zz = z2 * q[2];
  m[0] = 1.0 - (yy + zz); m[1] = xy - wz;         m[2] = xz + wy;
  m[3] = xy + wz;         m[4] = 1.0 - (xx + zz); m[5] = yz - wx;
  m[6] = xz - wy;         m[7] = yz + wx;         m[8] = 1.0 - (xx + yy);
}
MV_HD void quat_prod(const double* a, const double* b, double* o) {
  const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  const double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  const double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  const double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
MV_HD void quat_rotate(const double* q, const double* v, double* o) {   // v + w*2(u x v) + u x 2(u x v)
  double uv[3]; cross(q, v, uv);
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  double c[3]; cross(q, uv, c);
  const double x = v[0] + q[3] * uv[0] + c[0], y = v[1] + q[3] * uv[1] + c[1], z = v[2] + q[3] * uv[2] + c[2];
  o[0] = x; o[1] = y; o[2] = z;
}
MV_HD void quat_normalize(double* q) {
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

// ---- angle-axis (Ceres rotation.h semantics) ------------------------------------------------------
MV_HD void aa_of_matrix(const double* m, double* aa) {
  double w, x, y, z;
  const double tr = m[0] + m[4] + m[8];
  if (tr >= 0.0) {
    double s = sqrt(tr + 1.0);
    w = 0.5 * s; s = 0.5 / s;
    x = (m[7] - m[5]) * s; y = (m[2] - m[6]) * s; z = (m[3] - m[1]) * s;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    double v[3];
    v[i] = 0.5 * s; s = 0.5 / s;
    w = (m[3 * k + j] - m[3 * j + k]) * s;
    v[j] = (m[3 * j + i] + m[3 * i + j]) * s;
    v[k] = (m[3 * k + i] + m[3 * i + k]) * s;
    x = v[0]; y = v[1]; z = v[2];
  }
  const double s2 = x * x + y * y + z * z;
  if (s2 > 0.0) {
    const double sn = sqrt(s2);
    const double two_theta = 2.0 * ((w < 0.0) ? atan2(-sn, -w) : atan2(sn, w));
    const double kk = two_theta / sn;
    aa[0] = x * kk; aa[1] = y * kk; aa[2] = z * kk;
  } else {
    aa[0] = x * 2.0; aa[1] = y * 2.0; aa[2] = z * 2.0;
  }
}
MV_HD void matrix_of_aa(const double* aa, double* m) {
  const double th2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (th2 > 2.220446049250313e-16) {
    const double th = sqrt(th2);
    const double wx = aa[0] / th, wy = aa[1] / th, wz = aa[2] / th;
    const double c = cos(th), s = sin(th), oc = 1.0 - c;
    m[0] = c + wx * wx * oc;       m[1] = wx * wy * oc - wz * s;  m[2] = wy * s + wx * wz * oc;
    m[3] = wz * s + wx * wy * oc;  m[4] = c + wy * wy * oc;       m[5] = -wx * s + wy * wz * oc;
    m[6] = -wy * s + wx * wz * oc; m[7] = wx * s + wy * wz * oc;  m[8] = c + wz * wz * oc;
  } else {
    m[0] = 1.0;    m[1] = -aa[2]; m[2] = aa[1];
    m[3] = aa[2];  m[4] = 1.0;    m[5] = -aa[0];
    m[6] = -aa[1]; m[7] = aa[0];  m[8] = 1.0;
  }
}
// Rotation actually applied by ceres::AngleAxisRotatePoint for this vector (Rodrigues, or I + [w]x when tiny).
MV_HD void rotation_of_aa_functor(const double* aa, double* m) {
  const double th2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (th2 > 2.220446049250313e-16) { matrix_of_aa(aa, m); return; }
  m[0] = 1.0;    m[1] = -aa[2]; m[2] = aa[1];
  m[3] = aa[2];  m[4] = 1.0;    m[5] = -aa[0];
  m[6] = -aa[1]; m[7] = aa[0];  m[8] = 1.0;
}
// Right Jacobian of SO(3): R(w + dw) = R(w) Exp(Jr(w) dw).  Identity in Ceres' first-order branch.
MV_HD void so3_right_jacobian(const double* w, double* J) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  for (int i = 0; i < 9; ++i) J[i] = 0.0;
  J[0] = J[4] = J[8] = 1.0;
  if (!(th2 > 2.220446049250313e-16)) return;
  const double th = sqrt(th2);
  double a, b;
  if (th < 1e-4) { a = 0.5 - th2 / 24.0; b = 1.0 / 6.0 - th2 / 120.0; }
  else { a = (1.0 - cos(th)) / th2; b = (th - sin(th)) / (th2 * th); }
  const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double W2[9]; matmul(W, W, W2);
  for (int i = 0; i < 9; ++i) J[i] += -a * W[i] + b * W2[i];
}

// ---- Sophus SE3 exp and group product --------------------------------------------------------------
MV_HD void se3_exp(const double* tg /*upsilon, omega*/, double* out7) {
  const double* u = tg; const double* w = tg + 3;
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = sqrt(th2);
  double im, re;
  if (th < 1e-10) {
    const double th4 = th2 * th2;
    im = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
    re = 1.0 - 0.5 * th2 + (1.0 / 384.0) * th4;
  } else {
    im = sin(0.5 * th) / th; re = cos(0.5 * th);
  }
  double q[4] = {im * w[0], im * w[1], im * w[2], re};
  quat_normalize(q);
  double V[9];
  if (th < 1e-10) {
    matrix_of_quat(q, V);
  } else {
    const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double W2[9]; matmul(W, W, W2);
    const double c1 = (1.0 - cos(th)) / th2, c2 = (th - sin(th)) / (th2 * th);
    for (int i = 0; i < 9; ++i) V[i] = c1 * W[i] + c2 * W2[i];
    V[0] += 1.0; V[4] += 1.0; V[8] += 1.0;
  }
  double t[3]; matvec(V, u, t);
  out7[0] = q[0]; out7[1] = q[1]; out7[2] = q[2]; out7[3] = q[3]; out7[4] = t[0]; out7[5] = t[1]; out7[6] = t[2];
}
MV_HD void se3_compose(const double* a7, const double* b7, double* out7) {   // a * b, renormalised
  double rt[3]; quat_rotate(a7, b7 + 4, rt);
  double q[4]; quat_prod(a7, b7, q);
  quat_normalize(q);
  const double t0 = a7[4] + rt[0], t1 = a7[5] + rt[1], t2 = a7[6] + rt[2];
  out7[0] = q[0]; out7[1] = q[1]; out7[2] = q[2]; out7[3] = q[3]; out7[4] = t0; out7[5] = t1; out7[6] = t2;
}

// ---- parameter blocks ---------------------------------------------------------------------------------
enum { PARAM_AA = 0, PARAM_QUAT = 1, PARAM_SE3 = 2 };
MV_HD int ambient_size(int param) { return param == PARAM_AA ? 6 : 7; }

MV_HD void pose16_to_Rt(const double* P, Rt* o) {
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) o->R[3 * i + j] = P[4 * j + i]; o->t[i] = P[12 + i]; }
}
MV_HD void Rt_to_pose16(const Rt* a, double* P) {
  for (int i = 0; i < 16; ++i) P[i] = 0.0;
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) P[4 * j + i] = a->R[3 * i + j]; P[12 + i] = a->t[i]; }
  P[15] = 1.0;
}
MV_HD void param_of_pose(int param, const double* P16, double* x) {
  Rt a; pose16_to_Rt(P16, &a);
  if (param == PARAM_AA) { aa_of_matrix(a.R, x); x[3] = a.t[0]; x[4] = a.t[1]; x[5] = a.t[2]; }
  else { quat_of_matrix(a.R, x); x[4] = a.t[0]; x[5] = a.t[1]; x[6] = a.t[2]; }
}
// The rotation/translation the cost functors apply for these parameters.
MV_HD void Rt_of_param(int param, const double* x, Rt* o) {
  if (param == PARAM_AA) { rotation_of_aa_functor(x, o->R); o->t[0] = x[3]; o->t[1] = x[4]; o->t[2] = x[5]; }
  else { matrix_of_quat(x, o->R); o->t[0] = x[4]; o->t[1] = x[5]; o->t[2] = x[6]; }
}
// What the reference writes back into Frame::pose (axisAngleToIso / eigenQuaternionToIso / sophusToIso).
MV_HD void pose_of_param(int param, const double* x, double* P16) {
  Rt a;
  if (param == PARAM_AA) { matrix_of_aa(x, a.R); a.t[0] = x[3]; a.t[1] = x[4]; a.t[2] = x[5]; }
  else { matrix_of_quat(x, a.R); a.t[0] = x[4]; a.t[1] = x[5]; a.t[2] = x[6]; }
  Rt_to_pose16(&a, P16);
}
// x (+) delta, delta in the parameterisation's own 6-dof tangent.
MV_HD void param_plus(int param, const double* x, const double* d, double* o) {
  if (param == PARAM_AA) { for (int i = 0; i < 6; ++i) o[i] = x[i] + d[i]; return; }
  if (param == PARAM_QUAT) {   // [sin|d| d/|d|, cos|d|] (x) q  ;  t + dt      (eigen_quaternion.h:89-106)
    const double n = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (n > 0.0) {
      const double s = sin(n) / n;
      const double dq[4] = {s * d[0], s * d[1], s * d[2], cos(n)};
      double q[4]; quat_prod(dq, x, q);
      o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = q[3];
    } else { o[0] = x[0]; o[1] = x[1]; o[2] = x[2]; o[3] = x[3]; }
    o[4] = x[4] + d[3]; o[5] = x[5] + d[4]; o[6] = x[6] + d[5];
    return;
  }
  double e[7]; se3_exp(d, e);   // x * exp(delta)  (sophus_se3.h:31-38)
  double r[7]; se3_compose(x, e, r);
  for (int i = 0; i < 7; ++i) o[i] = r[i];
}
// K (6x6 row-major): canonical body tangent (upsilon, omega) of T*exp(xi) = K * parameterisation tangent.
//   SE3 : identity.                    QUAT: delta = (dq, dt): omega = 2 R^T dq, upsilon = R^T dt.
//   AA  : delta = (dw, dt): omega = Jr(w) dw, upsilon = R^T dt.
MV_HD void tangent_map(int param, const double* x, const Rt* a, double* K) {
  for (int i = 0; i < 36; ++i) K[i] = 0.0;
  if (param == PARAM_SE3) { for (int i = 0; i < 6; ++i) K[7 * i] = 1.0; return; }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) K[6 * i + 3 + j] = a->R[3 * j + i];                 // upsilon = R^T dt
  if (param == PARAM_QUAT) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) K[6 * (3 + i) + j] = 2.0 * a->R[3 * j + i];       // omega = 2 R^T dq
  } else {
    double J[9]; so3_right_jacobian(x, J);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) K[6 * (3 + i) + j] = J[3 * i + j];                // omega = Jr dw
  }
}


// ---- general (non-unit quaternion) frame model ---------------------------------------------------------------------
// With a non-unit quaternion (a non-rigid Isometry squeezed through Quaterniond(Matrix3d), icp-ceres.cpp:236 /
// so3.hpp:666-668) the reference's functors still apply the polynomial map F(q) v = v + 2w(u x v) + 2u x (u x v)
// (Eigen _transformVector; toRotationMatrix() is the same polynomial), which is linear in v but no rotation, and Ceres
// differentiates it as such.  For one frame: y(v) = F v + t and, per local tangent direction j of the active
// parameterisation,  d y(v) / d delta_j = D_j v + c_j,  d (F n) / d delta_j = D_j n,  with
//   D_j = sum_c P_q[c][j] dF/dq_c,   c_j = P_t[:, j],   P = d Plus(x, delta) / d delta at 0  (7x6, quaternion rows xyzw).
struct FrameGen { double F[9]; double t[3]; double D[6][9]; double c[6][3]; };

MV_HD void frame_general(int param, const double* x, FrameGen* o) {
  const double qx = x[0], qy = x[1], qz = x[2], qw = x[3];
  matrix_of_quat(x, o->F);
  o->t[0] = x[4]; o->t[1] = x[5]; o->t[2] = x[6];
  // dF/dq_c: F = I + 2w[u]x + 2([u]x)^2
  double M[4][9];
  const double u[3] = {qx, qy, qz};
  const double U[9] = {0, -qz, qy, qz, 0, -qx, -qy, qx, 0};
  for (int i = 0; i < 9; ++i) M[3][i] = 2.0 * U[i];                      // d/dw
  for (int c = 0; c < 3; ++c) {
    double E[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const int a = (c + 1) % 3, b = (c + 2) % 3;
    E[3 * b + a] = 1.0; E[3 * a + b] = -1.0;                             // [e_c]x
    double EU[9], UE[9]; matmul(E, U, EU); matmul(U, E, UE);
    for (int i = 0; i < 9; ++i) M[c][i] = 2.0 * qw * E[i] + 2.0 * (EU[i] + UE[i]);
  }
  (void)u;
  double Pq[4][6], Pt[3][6];
  for (int r = 0; r < 4; ++r) for (int j = 0; j < 6; ++j) Pq[r][j] = 0.0;
  for (int r = 0; r < 3; ++r) for (int j = 0; j < 6; ++j) Pt[r][j] = 0.0;
  if (param == PARAM_QUAT) {   // EigenQuaternionParameterization::ComputeJacobian (eigen_quaternion.h:108-114), rows xyzw; t additive
    Pq[0][0] = qw;  Pq[0][1] = qz;  Pq[0][2] = -qy;
    Pq[1][0] = -qz; Pq[1][1] = qw;  Pq[1][2] = qx;
    Pq[2][0] = qy;  Pq[2][1] = -qx; Pq[2][2] = qw;
    Pq[3][0] = -qx; Pq[3][1] = -qy; Pq[3][2] = -qz;
    for (int i = 0; i < 3; ++i) Pt[i][3 + i] = 1.0;
  } else {                     // SophusSE3Plus differentiated through x * exp(delta) INCLUDING the renormalisation
                               // (AutoDiffLocalParameterization, the multiview default: sophus_se3.h:64-68, icp-ceres.h:42)
    const double n2 = qx * qx + qy * qy + qz * qz + qw * qw, nn = sqrt(n2);
    for (int i = 0; i < 3; ++i) {
      const double h[4] = {i == 0 ? 0.5 : 0.0, i == 1 ? 0.5 : 0.0, i == 2 ? 0.5 : 0.0, 0.0};
      double g[4]; quat_prod(x, h, g);                                   // d(q (x) q_delta)/d omega_i
      const double dot = (g[0] * qx + g[1] * qy + g[2] * qz + g[3] * qw) / n2;
      Pq[0][3 + i] = (g[0] - qx * dot) / nn; Pq[1][3 + i] = (g[1] - qy * dot) / nn;
      Pq[2][3 + i] = (g[2] - qz * dot) / nn; Pq[3][3 + i] = (g[3] - qw * dot) / nn;
      for (int r = 0; r < 3; ++r) Pt[r][i] = o->F[3 * r + i];            // d t / d upsilon_i = F e_i
    }
  }
  for (int j = 0; j < 6; ++j) {
    for (int i = 0; i < 9; ++i) o->D[j][i] = Pq[0][j] * M[0][i] + Pq[1][j] * M[1][i] + Pq[2][j] * M[2][i] + Pq[3][j] * M[3][i];
    for (int r = 0; r < 3; ++r) o->c[j][r] = Pt[r][j];
  }
}

}  // namespace mv
