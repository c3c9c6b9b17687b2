// oracle/ref_functors.cpp -- TEST INFRASTRUCTURE ONLY (the checker's checker).
//
// Compiles the REFERENCE's own cost-functor text -- include/icp-ceres.h (all twelve functors, :47-554) and the quaternion local
// parameterisation include/eigen_quaternion.h:54-119 -- UNMODIFIED, from where it lies under /root/reference, against
// oracle/stubs/ (a ~150-line subset of Eigen's interface, class shells for Ceres, three members of Sophus::SE3Group) and
// differentiates it with the oracle's forward-mode Jet exactly as ceres::AutoDiffCostFunction would.  Output:
// oracle/_ref/libref_functors.so (recipe: oracle/Makefile `ref`).  tests/test_oracle_functor_pin.py asserts that the oracle's
// restated functors (oracle_icp.cpp: functor_p2p / functor_p2plane) give the same residuals and Jacobians.
//
// What this pins: the functor layer of the LM step (which pose is applied to which point, where toRotationMatrix() is used
// instead of q*v, that normals are rotated but not translated, residual order, parameter layouts) and the hand-written
// quaternion Plus / Jacobian.  What stays unpinned: Ceres' solver loop, its Jet and AngleAxisRotatePoint (restated in the stubs),
// Eigen's own arithmetic (restated in mini_eigen.h), Sophus' exp / internalJacobian (include/sophus_se3.h needs the real Sophus:
// skipped by pre-defining its include guard; oracle/geom.h restates them and tests/test_oracle_math.py checks them against
// the Sophus test vectors).
#include <cmath>
#include <cstdint>
#include <iostream>
#include <memory>
#include <vector>
#include "jet.h"


#define FRAME_H      // frame.h pulls nanoflann + gflags + common.h (directory listing, PCA, ...): nothing the functors use
#define SOPHUS_SE3   // sophus_se3.h needs the real Sophus (exp, internalJacobian): see header comment
class Frame;
#include <Eigen/Dense>
#include <sophus/se3.hpp>
#include "icp-ceres.h"   // -I/root/reference/include : the reference's file, unmodified

using namespace ICPCostFunctions;

namespace {
template <int N> using J = orc::Jet<N>;

// F(cam1[G], cam2[G]) -> r[NR], 2 blocks of G scalars each (aa / se3 global functors)
template <class F, int G, int NR> void eval2(const F& f, const double* c1, const double* c2, double* r, double* jac) {
  J<2 * G> a[G], b[G], out[NR];
  for (int i = 0; i < G; ++i) { a[i] = J<2 * G>(c1[i], i); b[i] = J<2 * G>(c2[i], G + i); }
  f(a, b, out);
  for (int q = 0; q < NR; ++q) { r[q] = out[q].a; for (int k = 0; k < 2 * G; ++k) jac[q * 2 * G + k] = out[q].v[k]; }
}
// quaternion global functors: four blocks (q1[4], t1[3], q2[4], t2[3]); derivative columns in the order q1 t1 q2 t2
template <class F, int NR> void eval4(const F& f, const double* c1, const double* c2, double* r, double* jac) {
  J<14> q1[4], t1[3], q2[4], t2[3], out[NR];
  for (int i = 0; i < 4; ++i) { q1[i] = J<14>(c1[i], i); q2[i] = J<14>(c2[i], 7 + i); }
  for (int i = 0; i < 3; ++i) { t1[i] = J<14>(c1[4 + i], 4 + i); t2[i] = J<14>(c2[4 + i], 11 + i); }
  f(q1, t1, q2, t2, out);
  for (int q = 0; q < NR; ++q) { r[q] = out[q].a; for (int k = 0; k < 14; ++k) jac[q * 14 + k] = out[q].v[k]; }
}
template <class F, int G, int NR> void eval1(const F& f, const double* c1, double* r, double* jac) {
  J<G> a[G], out[NR];
  for (int i = 0; i < G; ++i) a[i] = J<G>(c1[i], i);
  f(a, out);
  for (int q = 0; q < NR; ++q) { r[q] = out[q].a; for (int k = 0; k < G; ++k) jac[q * G + k] = out[q].v[k]; }
}
template <class F, int NR> void eval1q(const F& f, const double* c1, double* r, double* jac) {
  J<7> q1[4], t1[3], out[NR];
  for (int i = 0; i < 4; ++i) q1[i] = J<7>(c1[i], i);
  for (int i = 0; i < 3; ++i) t1[i] = J<7>(c1[4 + i], 4 + i);
  f(q1, t1, out);
  for (int q = 0; q < NR; ++q) { r[q] = out[q].a; for (int k = 0; k < 7; ++k) jac[q * 7 + k] = out[q].v[k]; }
}
}  // namespace

extern "C" {
// Multiview ("Global") functors, icp-ceres.h:49-316.  param: 0 aa [w t] (6+6 columns), 1 quat / 2 se3 [qx qy qz qw tx ty tz]
// (7+7 columns).  plane: 0 point-to-point (3 residuals), 1 point-to-plane (1).  jac is row-major [NR][2G].
void ref_functor_global(int param, int plane, const double* cam1, const double* cam2, const double* src, const double* dst,
                        const double* nor, double* r, double* jac) {
  const Eigen::Vector3d s(src[0], src[1], src[2]), d(dst[0], dst[1], dst[2]), n(nor ? nor[0] : 0, nor ? nor[1] : 0, nor ? nor[2] : 0);
  if (param == 0) {
    if (plane) eval2<PointToPlaneErrorGlobal_CeresAngleAxis, 6, 1>(PointToPlaneErrorGlobal_CeresAngleAxis(d, s, n), cam1, cam2, r, jac);
    else eval2<PointToPointErrorGlobal_CeresAngleAxis, 6, 3>(PointToPointErrorGlobal_CeresAngleAxis(d, s), cam1, cam2, r, jac);
  } else if (param == 1) {
    if (plane) eval4<PointToPlaneErrorGlobal, 1>(PointToPlaneErrorGlobal(d, s, n), cam1, cam2, r, jac);
    else eval4<PointToPointErrorGlobal, 3>(PointToPointErrorGlobal(d, s), cam1, cam2, r, jac);
  } else {
    if (plane) eval2<PointToPlaneErrorGlobal_SophusSE3, 7, 1>(PointToPlaneErrorGlobal_SophusSE3(d, s, n), cam1, cam2, r, jac);
    else eval2<PointToPointErrorGlobal_SophusSE3, 7, 3>(PointToPointErrorGlobal_SophusSE3(d, s), cam1, cam2, r, jac);
  }
}
// Pairwise functors, icp-ceres.h:320-552 (one pose; dst / nor in the fixed frame).  jac is row-major [NR][G].
void ref_functor_pairwise(int param, int plane, const double* cam, const double* src, const double* dst, const double* nor,
                          double* r, double* jac) {
  const Eigen::Vector3d s(src[0], src[1], src[2]), d(dst[0], dst[1], dst[2]), n(nor ? nor[0] : 0, nor ? nor[1] : 0, nor ? nor[2] : 0);
  if (param == 0) {
    if (plane) eval1<PointToPlaneError_CeresAngleAxis, 6, 1>(PointToPlaneError_CeresAngleAxis(d, s, n), cam, r, jac);
    else eval1<PointToPointError_CeresAngleAxis, 6, 3>(PointToPointError_CeresAngleAxis(d, s), cam, r, jac);
  } else if (param == 1) {
    if (plane) eval1q<PointToPlaneError_EigenQuaternion, 1>(PointToPlaneError_EigenQuaternion(d, s, n), cam, r, jac);
    else eval1q<PointToPointError_EigenQuaternion, 3>(PointToPointError_EigenQuaternion(d, s), cam, r, jac);
  } else {
    if (plane) eval1<PointToPlaneError_SophusSE3, 7, 1>(PointToPlaneError_SophusSE3(d, s, n), cam, r, jac);
    else eval1<PointToPointError_SophusSE3, 7, 3>(PointToPointError_SophusSE3(d, s), cam, r, jac);
  }
}
// eigen_quaternion::EigenQuaternionParameterization (eigen_quaternion.h:89-117): Plus and the 4x3 Jacobian (row-major)
void ref_quat_plus(const double* x, const double* delta, double* out) { eigen_quaternion::EigenQuaternionParameterization p; p.Plus(x, delta, out); }
void ref_quat_jacobian(const double* x, double* jac12) { eigen_quaternion::EigenQuaternionParameterization p; p.ComputeJacobian(x, jac12); }
int ref_functors_abi(void) { return 1; }
}
