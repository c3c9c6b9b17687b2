// apps/closed_form.hpp -- the `pairwise` benchmark's comparison baseline "closed form" (src/internal/icp-closedform.cpp:9-54),
// on the host: an O(N) accumulation plus a 3x3 / 6x6 solve.  Not part of the engine (it never runs in the multiview loop);
// it is here so that apps/pairwise_b200 prints the same table as the reference binary.
#pragma once
#include <cmath>
#include <vector>
#include "mini_types.hpp"

namespace closed_form {
// symmetric 3x3 eigen-decomposition by cyclic Jacobi rotations: A = V diag(w) V^T
inline void eig_sym3(double A[3][3], double w[3], double V[3][3]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = (i == j);
  for (int sweep = 0; sweep < 60; ++sweep) {
    if (A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2] < 1e-300) break;
    for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
      if (A[p][q] == 0.0) continue;
      const double th = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
      const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0)), c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < 3; ++k) { const double a = A[k][p], b = A[k][q]; A[k][p] = c * a - s * b; A[k][q] = s * a + c * b; }
      for (int k = 0; k < 3; ++k) { const double a = A[p][k], b = A[q][k]; A[p][k] = c * a - s * b; A[q][k] = s * a + c * b; }
      for (int k = 0; k < 3; ++k) { const double a = V[k][p], b = V[k][q]; V[k][p] = c * a - s * b; V[k][q] = s * a + c * b; }
    }
  }
  for (int i = 0; i < 3; ++i) w[i] = A[i][i];
}

// ICP_Closedform::pointToPoint (icp-closedform.cpp:9-26): R = U V^T of K = sum (q - qbar)(p - pbar)^T -- the orthogonal polar
// factor of K, obtained as K (K^T K)^(-1/2) -- with the reference's `R.col(2) *= -1` when det R < 0; t = qbar - R pbar.
inline Eigen::Isometry3d pointToPoint(const std::vector<Eigen::Vector3d>& src, const std::vector<Eigen::Vector3d>& dst) {
  const size_t n = src.size();
  double pb[3] = {0, 0, 0}, qb[3] = {0, 0, 0}, K[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (size_t i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) { pb[a] += src[i][a]; qb[a] += dst[i][a]; }
  for (int a = 0; a < 3; ++a) { pb[a] /= (double)n; qb[a] /= (double)n; }
  for (size_t i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) K[a][b] += (dst[i][a] - qb[a]) * (src[i][b] - pb[b]);
  double S[3][3], w[3], V[3][3], R[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { S[a][b] = 0; for (int c = 0; c < 3; ++c) S[a][b] += K[c][a] * K[c][b]; }
  eig_sym3(S, w, V);
  for (int j = 0; j < 3; ++j) {
    const double sg = std::sqrt(w[j] > 0 ? w[j] : 0.0);
    for (int a = 0; a < 3; ++a) { double u = 0; for (int c = 0; c < 3; ++c) u += K[a][c] * V[c][j]; u /= sg; for (int b = 0; b < 3; ++b) R[a][b] += u * V[b][j]; }
  }
  const double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) + R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
  if (det < 0) for (int a = 0; a < 3; ++a) R[a][2] = -R[a][2];
  Eigen::Isometry3d T;
  for (int a = 0; a < 3; ++a) { for (int b = 0; b < 3; ++b) T(a, b) = R[a][b]; T(a, 3) = qb[a] - (R[a][0] * pb[0] + R[a][1] * pb[1] + R[a][2] * pb[2]); }
  return T;
}

// ICP_Closedform::pointToPlane (icp-closedform.cpp:30-54): small-angle normal equations over rows [p x n ; n], LDL^T,
// R = Rx(x0) Ry(x1) Rz(x2), t = x[3..5].
inline Eigen::Isometry3d pointToPlane(const std::vector<Eigen::Vector3d>& src, const std::vector<Eigen::Vector3d>& dst, const std::vector<Eigen::Vector3d>& nor) {
  double C[6][6] = {{0}}, d[6] = {0, 0, 0, 0, 0, 0};
  for (size_t i = 0; i < src.size(); ++i) {
    const Eigen::Vector3d &p = src[i], &q = dst[i], &m = nor[i];
    const double a[6] = {p[1] * m[2] - p[2] * m[1], p[2] * m[0] - p[0] * m[2], p[0] * m[1] - p[1] * m[0], m[0], m[1], m[2]};
    const double e = (p[0] - q[0]) * m[0] + (p[1] - q[1]) * m[1] + (p[2] - q[2]) * m[2];
    for (int r = 0; r < 6; ++r) { for (int c = 0; c < 6; ++c) C[r][c] += a[r] * a[c]; d[r] -= a[r] * e; }
  }
  double L[6][6] = {{0}}, D[6], y[6], x[6];
  for (int j = 0; j < 6; ++j) {
    double dj = C[j][j]; for (int k = 0; k < j; ++k) dj -= L[j][k] * L[j][k] * D[k];
    D[j] = dj; L[j][j] = 1;
    for (int i = j + 1; i < 6; ++i) { double v = C[i][j]; for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k] * D[k]; L[i][j] = v / dj; }
  }
  for (int i = 0; i < 6; ++i) { y[i] = d[i]; for (int k = 0; k < i; ++k) y[i] -= L[i][k] * y[k]; }
  for (int i = 5; i >= 0; --i) { x[i] = y[i] / D[i]; for (int k = i + 1; k < 6; ++k) x[i] -= L[k][i] * x[k]; }
  const double ca = std::cos(x[0]), sa = std::sin(x[0]), cb = std::cos(x[1]), sb = std::sin(x[1]), cg = std::cos(x[2]), sg = std::sin(x[2]);
  Eigen::Isometry3d T;
  T(0, 0) = cb * cg;                T(0, 1) = -cb * sg;               T(0, 2) = sb;
  T(1, 0) = sa * sb * cg + ca * sg; T(1, 1) = -sa * sb * sg + ca * cg; T(1, 2) = -sa * cb;
  T(2, 0) = -ca * sb * cg + sa * sg; T(2, 1) = ca * sb * sg + sa * cg; T(2, 2) = ca * cb;
  for (int a = 0; a < 3; ++a) T(a, 3) = x[3 + a];
  return T;
}
}  // namespace closed_form
