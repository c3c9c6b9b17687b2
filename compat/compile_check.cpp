// Compile-and-link check of compat/mvicp_compat.hpp against the reference's container shapes (include/frame.h:18-46).
#include <Eigen/Dense>
#include <memory>
#include <vector>
#include "mvicp_compat.hpp"
struct Correspondance { int first; int second; double dist; };
struct OutgoingEdge { int neighbourIdx; float weight; std::vector<Correspondance> correspondances; };
struct Frame {
  std::vector<Eigen::Vector3d> pts, nor; bool fixed = false; Eigen::Isometry3d pose; std::vector<OutgoingEdge> neighbours;
};
int main() {
  std::vector<std::shared_ptr<Frame>> frames;
  mvicp_compat::Session<Frame> s;
  if (frames.empty()) return 0;            // never runs without data; instantiates every template below
  mvicp_compat::recomputeNormals(s, frames, 10);
  mvicp_compat::computeClosestPoints(s, frames, 0.05f, true);
  mvicp_compat::optimize(s, frames, MVICP_PARAM_SE3, true, true);
  double P[16]; mvicp_compat::pairwise(MVICP_PARAM_AA, false, frames[0]->pts, frames[1]->pts, (const std::vector<Eigen::Vector3d>*)nullptr, P);
  return 0;
}
