// oracle/stubs/ceres/stub_ceres.h -- TEST INFRASTRUCTURE ONLY.  What include/icp-ceres.h and include/eigen_quaternion.h name of
// Ceres, so that they compile without it: class shells (never solved with) and ceres::AngleAxisRotatePoint, restated from
// Ceres 1.13 rotation.h [ext-knowledge: Ceres is not under /root/reference] -- Rodrigues with the first-order branch for
// theta^2 <= DBL_EPSILON.
#pragma once
#include <cmath>
#include <limits>
#define CHECK_NE(a, b) ((void)0)
namespace ceres {
struct CostFunction { virtual ~CostFunction() {} };
template <typename F, int R, int... N> struct AutoDiffCostFunction : CostFunction {
  std::unique_ptr<F> f;
  explicit AutoDiffCostFunction(F* p) : f(p) {}
};
struct LocalParameterization {
  virtual ~LocalParameterization() {}
  virtual bool Plus(const double* x, const double* delta, double* x_plus_delta) const = 0;
  virtual bool ComputeJacobian(const double* x, double* jacobian) const = 0;
  virtual int GlobalSize() const = 0;
  virtual int LocalSize() const = 0;
};
template <typename F, int G, int L> struct AutoDiffLocalParameterization : LocalParameterization {};
template <typename T, int row_stride, int col_stride> struct MatrixAdapter {
  T* p; explicit MatrixAdapter(T* q) : p(q) {}
  T& operator()(int r, int c) const { return p[r * row_stride + c * col_stride]; }
};
template <typename T> MatrixAdapter<T, 3, 1> RowMajorAdapter3x3(T* p) { return MatrixAdapter<T, 3, 1>(p); }

template <typename T> inline T DotProduct(const T x[3], const T y[3]) { return (x[0] * y[0] + x[1] * y[1] + x[2] * y[2]); }

template <typename T> inline void AngleAxisRotatePoint(const T angle_axis[3], const T pt[3], T result[3]) {
  using std::sqrt; using std::cos; using std::sin;
  const T theta2 = DotProduct(angle_axis, angle_axis);
  if (theta2 > T(std::numeric_limits<double>::epsilon())) {
    const T theta = sqrt(theta2);
    const T costheta = cos(theta);
    const T sintheta = sin(theta);
    const T theta_inverse = T(1.0) / theta;
    const T w[3] = {angle_axis[0] * theta_inverse, angle_axis[1] * theta_inverse, angle_axis[2] * theta_inverse};
    const T w_cross_pt[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
    result[0] = pt[0] * costheta + w_cross_pt[0] * sintheta + w[0] * tmp;
    result[1] = pt[1] * costheta + w_cross_pt[1] * sintheta + w[1] * tmp;
    result[2] = pt[2] * costheta + w_cross_pt[2] * sintheta + w[2] * tmp;
  } else {
    const T w_cross_pt[3] = {angle_axis[1] * pt[2] - angle_axis[2] * pt[1], angle_axis[2] * pt[0] - angle_axis[0] * pt[2],
                             angle_axis[0] * pt[1] - angle_axis[1] * pt[0]};
    result[0] = pt[0] + w_cross_pt[0];
    result[1] = pt[1] + w_cross_pt[1];
    result[2] = pt[2] + w_cross_pt[2];
  }
}
}  // namespace ceres
