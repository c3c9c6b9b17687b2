// oracle/stubs/mini_eigen.h -- TEST INFRASTRUCTURE ONLY.
// The ~150 lines of Eigen's interface that the reference's cost functors (include/icp-ceres.h:47-554) and quaternion local
// parameterisation (include/eigen_quaternion.h:54-119) use, so that those two headers compile UNMODIFIED, from where they lie
// under /root/reference, into oracle/_ref/libref_functors.so (recipe: oracle/Makefile).  Eigen itself is not in this image.
// Arithmetic follows Eigen 3.3 [ext-knowledge]: Quaternion * vector is QuaternionBase::_transformVector
// (uv = q.vec x v; uv += uv; v + w uv + q.vec x uv), toRotationMatrix is QuaternionBase::toRotationMatrix (tx = 2x ...),
// quaternion product is quat_product<>, and 3-term sums (dot, norm, matrix * vector coefficients) associate as
// x0 + (x1 + x2) (redux_novec_unroller splits a length-3 reduction into halves of 1 and 2).
#pragma once
#include <cmath>
#include <memory>
#include <vector>

namespace Eigen {

template <typename T, int R, int C> struct Matrix;

template <typename T> struct CommaInit {
  T* p; int i;
  CommaInit& operator,(const T& v) { p[i++] = v; return *this; }
};

template <typename T, int R> struct Matrix<T, R, 1> {
  T v[R];
  Matrix() { for (int i = 0; i < R; ++i) v[i] = T(0.0); }
  Matrix(const T& a, const T& b, const T& c) { static_assert(R == 3, "3-vector"); v[0] = a; v[1] = b; v[2] = c; }
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
  T& operator()(int i) { return v[i]; }
  const T& operator()(int i) const { return v[i]; }
  CommaInit<T> operator<<(const T& a) { v[0] = a; return CommaInit<T>{v, 1}; }
  Matrix& operator+=(const Matrix& o) { for (int i = 0; i < R; ++i) v[i] = v[i] + o.v[i]; return *this; }
  Matrix operator+(const Matrix& o) const { Matrix r; for (int i = 0; i < R; ++i) r.v[i] = v[i] + o.v[i]; return r; }
  Matrix operator-(const Matrix& o) const { Matrix r; for (int i = 0; i < R; ++i) r.v[i] = v[i] - o.v[i]; return r; }
  T dot(const Matrix& o) const { static_assert(R == 3, "3-vector"); return v[0] * o.v[0] + (v[1] * o.v[1] + v[2] * o.v[2]); }
  T squaredNorm() const { return dot(*this); }
  T norm() const { using std::sqrt; return sqrt(squaredNorm()); }
  Matrix cross(const Matrix& o) const {
    static_assert(R == 3, "3-vector");
    return Matrix(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
  }
  const T* data() const { return v; }
};
template <typename T, int R> inline Matrix<T, R, 1> operator*(const T& s, const Matrix<T, R, 1>& m) { Matrix<T, R, 1> r; for (int i = 0; i < R; ++i) r.v[i] = s * m.v[i]; return r; }

template <typename T> struct Matrix<T, 3, 3> {
  T m[3][3];   // m[row][col]
  T& coeffRef(int r, int c) { return m[r][c]; }
  const T& operator()(int r, int c) const { return m[r][c]; }
  Matrix<T, 3, 1> operator*(const Matrix<T, 3, 1>& x) const {
    Matrix<T, 3, 1> r;
    for (int i = 0; i < 3; ++i) r.v[i] = m[i][0] * x.v[0] + (m[i][1] * x.v[1] + m[i][2] * x.v[2]);
    return r;
  }
};

typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 3, 3> Matrix3d;
struct Isometry3d { double m[16]; };   // only named in declarations of icp-ceres.h:30-42
struct Isometry3f { float m[16]; };

template <typename T> struct Quaternion {
  T c[4];   // x y z w (coeffs() order)
  Quaternion() {}
  Quaternion(const T& w, const T& x, const T& y, const T& z) { c[0] = x; c[1] = y; c[2] = z; c[3] = w; }
  const T& x() const { return c[0]; } const T& y() const { return c[1]; } const T& z() const { return c[2]; } const T& w() const { return c[3]; }
  Matrix<T, 3, 1> vec() const { return Matrix<T, 3, 1>(c[0], c[1], c[2]); }
  Matrix<T, 3, 1> operator*(const Matrix<T, 3, 1>& v) const {   // _transformVector
    Matrix<T, 3, 1> uv = vec().cross(v);
    uv += uv;
    return v + w() * uv + vec().cross(uv);
  }
  Quaternion operator*(const Quaternion& b) const {   // quat_product
    const Quaternion& a = *this;
    return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                      a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                      a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                      a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
  }
  Matrix<T, 3, 3> toRotationMatrix() const {
    Matrix<T, 3, 3> res;
    const T tx = T(2) * x(), ty = T(2) * y(), tz = T(2) * z();
    const T twx = tx * w(), twy = ty * w(), twz = tz * w();
    const T txx = tx * x(), txy = ty * x(), txz = tz * x();
    const T tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
    res.coeffRef(0, 0) = T(1) - (tyy + tzz); res.coeffRef(0, 1) = txy - twz; res.coeffRef(0, 2) = txz + twy;
    res.coeffRef(1, 0) = txy + twz; res.coeffRef(1, 1) = T(1) - (txx + tzz); res.coeffRef(1, 2) = tyz - twx;
    res.coeffRef(2, 0) = txz - twy; res.coeffRef(2, 1) = tyz + twx; res.coeffRef(2, 2) = T(1) - (txx + tyy);
    return res;
  }
};
typedef Quaternion<double> Quaterniond;

// Map<const X>(ptr) converts to X by copying the coefficients; Map<X>(ptr) = value writes them back.
template <typename X> struct Map;
template <typename T> struct Map<const Quaternion<T>> : Quaternion<T> {
  explicit Map(const T* p) { for (int i = 0; i < 4; ++i) this->c[i] = p[i]; }
};
template <typename T> struct Map<Quaternion<T>> {
  T* p; explicit Map(T* q) : p(q) {}
  Map& operator=(const Quaternion<T>& q) { for (int i = 0; i < 4; ++i) p[i] = q.c[i]; return *this; }
};
template <typename T, int R> struct Map<const Matrix<T, R, 1>> : Matrix<T, R, 1> {
  explicit Map(const T* p) { for (int i = 0; i < R; ++i) this->v[i] = p[i]; }
};

}  // namespace Eigen
