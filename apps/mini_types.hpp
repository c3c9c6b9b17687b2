// apps/mini_types.hpp -- the two Eigen types and the three reference structs that the headless drivers need, with the
// reference's memory layout (include/frame.h:18-46), so that compat/mvicp_compat.hpp binds to them unchanged.  With the
// real Eigen + the reference's frame.h on the include path these definitions are simply not used.
#pragma once
#include <memory>
#include <vector>
namespace Eigen {
struct Vector3d { double v[3]; double* data() { return v; } const double* data() const { return v; } double& operator[](int i) { return v[i]; } double operator[](int i) const { return v[i]; } };
struct Isometry3d {   // 4x4 column-major
  double m[16];
  Isometry3d() { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0 : 0.0; }
  double* data() { return m; } const double* data() const { return m; }
  double& operator()(int r, int c) { return m[4 * c + r]; } double operator()(int r, int c) const { return m[4 * c + r]; }
};
}
struct Correspondance { int first; int second; double dist; };
struct OutgoingEdge { int neighbourIdx; float weight; std::vector<Correspondance> correspondances; };
struct Frame {
  std::vector<Eigen::Vector3d> pts, nor;
  bool fixed = false;
  Eigen::Isometry3d pose, poseGroundTruth;
  std::vector<OutgoingEdge> neighbours;
};
