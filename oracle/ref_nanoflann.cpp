// oracle/ref_nanoflann.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Thin shim that drives the REFERENCE's own vendored nanoflann.hpp (v1.1.9), compiled from where
// it lies (/root/reference/include/nanoflann.hpp, -I on the command line; never copied here), the
// way Frame::getClosestPoint does (src/internal/frame.cpp:187-206): KDTreeSingleIndexAdaptor over
// an L2_Simple_Adaptor<double>, leaf_max_size = 1, KNNResultSet(1), SearchParams(32, 0, false).
// The dataset adaptor restates Frame's (include/frame.h:67-92).  Output lands in oracle/_ref/
// (git-ignored, travels to the GPU box).  Used (a) to pin oracle_icp.cpp's correspondence
// restatement and (b) as the "reference" CPU baseline of the correspondence step in bench.py.
#include <cstdint>
#include <vector>
#include <nanoflann.hpp>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "geom.h"

namespace {
struct Cloud {
  const double* pts; size_t n;
  inline size_t kdtree_get_point_count() const { return n; }
  inline double kdtree_distance(const double* p1, const size_t idx_p2, size_t) const {
    const double d0 = p1[0] - pts[3 * idx_p2], d1 = p1[1] - pts[3 * idx_p2 + 1], d2 = p1[2] - pts[3 * idx_p2 + 2];
    return d0 * d0 + d1 * d1 + d2 * d2;
  }
  inline double kdtree_get_pt(const size_t idx, int dim) const { return pts[3 * idx + dim]; }
  template <class BBOX> bool kdtree_get_bbox(BBOX&) const { return false; }
};
typedef nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<double, Cloud>, Cloud, 3> tree_t;
struct Index { Cloud cloud; tree_t* tree; };
}  // namespace

extern "C" {
void* ref_kd_build(const double* pts, int64_t n) {
  Index* ix = new Index{{pts, (size_t)n}, nullptr};
  ix->tree = new tree_t(3, ix->cloud, nanoflann::KDTreeSingleIndexAdaptorParams(1 /* max leaf */));
  ix->tree->buildIndex();
  return ix;
}
void ref_kd_free(void* h) { Index* ix = (Index*)h; delete ix->tree; delete ix; }
void ref_kd_query(void* h, const double* q, int64_t* idx, double* d2) {
  size_t ret_index = 0; double out = 0;
  nanoflann::KNNResultSet<double> rs(1);
  rs.init(&ret_index, &out);
  ((Index*)h)->tree->findNeighbors(rs, q, nanoflann::SearchParams(32, 0, false));
  *idx = (int64_t)ret_index; *d2 = out;
}
// Frame::getNeighbours' knnSearch (frame.cpp:208-225): k nearest to a query, ascending distance
void ref_kd_knn(void* h, const double* q, int64_t k, int64_t* idx, double* d2) {
  std::vector<size_t> ri(k); std::vector<double> rd(k);
  ((Index*)h)->tree->knnSearch(q, (size_t)k, ri.data(), rd.data());
  for (int64_t i = 0; i < k; ++i) { idx[i] = (int64_t)ri[i]; d2[i] = rd[i]; }
}
// all src points of one edge (frame.cpp:129-138), transform restated in geom.h / oracle_icp.cpp
void ref_closest_points(void* h, const double* src_pts, int64_t n_src, const double* pose_src16,
                        const double* pose_dst16, int32_t* nn_idx, double* nn_d2, int num_threads) {
  double Rs[9], ts[3], Rd[9], td[3], Rinv[9];
  orc::pose16_split(pose_src16, Rs, ts); orc::pose16_split(pose_dst16, Rd, td);
  orc::mat3_inverse_cofactor(Rd, Rinv);
  tree_t* tree = ((Index*)h)->tree;
#pragma omp parallel for num_threads(num_threads > 0 ? num_threads : 1) schedule(static)
  for (int64_t k = 0; k < n_src; ++k) {
    double g[3], q[3];
    orc::mat3_mul_vec(Rs, src_pts + 3 * k, g);
    g[0] += ts[0]; g[1] += ts[1]; g[2] += ts[2];
    const double d[3] = {g[0] - td[0], g[1] - td[1], g[2] - td[2]};
    orc::mat3_mul_vec(Rinv, d, q);
    size_t ret_index = 0; double out = 0;
    nanoflann::KNNResultSet<double> rs(1);
    rs.init(&ret_index, &out);
    tree->findNeighbors(rs, q, nanoflann::SearchParams(32, 0, false));
    nn_idx[k] = (int32_t)ret_index; nn_d2[k] = out;
  }
}
}
