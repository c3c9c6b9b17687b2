// oracle/stubs/sophus/se3.hpp -- TEST INFRASTRUCTURE ONLY.  The three members of Sophus::SE3Group<T> that the reference's cost
// functors touch (include/icp-ceres.h:258-262,304-309,434-438,528-532): Map from 7 scalars [qx qy qz qw tx ty tz]
// (ext/sophus-ceres/sophus/se3.hpp:108-111,917), unit_quaternion(), translation().  The real header needs Eigen.
#pragma once
#include "../mini_eigen.h"
namespace Sophus {
template <typename T> struct SE3Group {
  Eigen::Quaternion<T> q; Eigen::Matrix<T, 3, 1> t;
  static const int num_parameters = 7; static const int DoF = 6;
  const Eigen::Quaternion<T>& unit_quaternion() const { return q; }
  const Eigen::Matrix<T, 3, 1>& translation() const { return t; }
};
typedef SE3Group<double> SE3d;
}  // namespace Sophus
namespace Eigen {
template <typename T> struct Map<const Sophus::SE3Group<T>> : Sophus::SE3Group<T> {
  explicit Map(const T* p) { for (int i = 0; i < 4; ++i) this->q.c[i] = p[i]; for (int i = 0; i < 3; ++i) this->t.v[i] = p[4 + i]; }
};
}  // namespace Eigen
