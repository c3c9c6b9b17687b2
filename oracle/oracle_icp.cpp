// oracle/oracle_icp.cpp -- TEST INFRASTRUCTURE ONLY.  CPU oracle for the mv-lm-icp hot path.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library; the product (mv_lm_icp_b200/csrc) never links or calls it.
//
// What is restated here (reference file:line):
//   * correspondence step  Frame::computeClosestPointsToNeighbours   src/internal/frame.cpp:91-185
//                          Frame::getClosestPoint                    src/internal/frame.cpp:187-206
//                          Frame::kdtree_distance                    include/frame.h:70-76
//   * problem builders     ICP_Ceres::ceresOptimizer{,_ceresAngleAxis,_sophusSE3}
//                                                                    src/internal/icp-ceres.cpp:220-475
//     pairwise solvers     pointToPoint_* / pointToPlane_*           src/internal/icp-ceres.cpp:137-218,525-565
//   * cost functors        ICPCostFunctions::*                       include/icp-ceres.h:49-552
//   * parameterisations    sophus_se3.h:10-60, eigen_quaternion.h:89-117 (see geom.h)
//   * the LM loop          ceres::Solve with getOptionsMedium        see lm.h (PARITY UNPINNED: Ceres absent)
//
// PINNING STATUS.  Correspondence step: pinned against the reference's own nanoflann.hpp compiled
// from /root/reference (oracle/_ref/libref_nanoflann.so, see ref_nanoflann.cpp and
// tests/test_oracle_corr.py).  LM step: PARITY UNPINNED -- Ceres/Eigen are not in /root/reference
// and not installable here; the functors are restated and differentiated with forward-mode Jets as
// Ceres' AutoDiffCostFunction does, the result is cross-checked against finite differences, scipy
// and the pairwise known-answer of README.md:141-150, but never against a Ceres binary.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "geom.h"
#include "jet.h"
#include "lm.h"

namespace orc {

// =================================================================================================
// Correspondence step
// =================================================================================================

// Frame::kdtree_distance (include/frame.h:70-76): d0*d0+d1*d1+d2*d2, left to right, no FMA
// (the whole oracle is compiled with -ffp-contract=off).
static inline double dist_sq(const double* q, const double* p) {
  const double d0 = q[0] - p[0], d1 = q[1] - p[1], d2 = q[2] - p[2];
  return d0 * d0 + d1 * d1 + d2 * d2;
}

// Per-edge constants of frame.cpp:117-118,131,136.
struct EdgeXform {
  double Rs[9], ts[3], Rinv[9], td[3];
  EdgeXform(const double* pose_src16, const double* pose_dst16) {
    double Rd[9];
    pose16_split(pose_src16, Rs, ts);
    pose16_split(pose_dst16, Rd, td);
    mat3_inverse_cofactor(Rd, Rinv);
  }
  // srcPtInGlobalFrame = srcCloud.pose * p (:131); srcPtinDstFrame = preInvRot * (g - preTra) (:136)
  inline void apply(const double* p, double* q) const {
    double g[3];
    mat3_mul_vec(Rs, p, g);
    g[0] = g[0] + ts[0]; g[1] = g[1] + ts[1]; g[2] = g[2] + ts[2];
    const double d[3] = {g[0] - td[0], g[1] - td[1], g[2] - td[2]};
    mat3_mul_vec(Rinv, d, q);
  }
};

// An independent exact 1-NN structure (NOT nanoflann's algorithm: median-split KD-tree with bucket
// leaves).  The *result* it must reproduce is nanoflann's: the point with the minimal fp64
// kdtree_distance.  Tie rule of this oracle: lowest index (nanoflann's is traversal order,
// nanoflann.hpp:1210; ties are counted by the tests, see SURVEY.md section 7 "hard parts").
struct KdTree {
  struct Node { int32_t left, right; int32_t begin, end; int axis; double split; double lo[3], hi[3]; };
  const double* pts; int64_t n;
  std::vector<int32_t> idx; std::vector<Node> nodes;
  static constexpr int kLeaf = 8;
  KdTree(const double* p, int64_t n_) : pts(p), n(n_), idx(n_) {
    for (int64_t i = 0; i < n; ++i) idx[i] = (int32_t)i;
    nodes.reserve(2 * (n / kLeaf + 1));
    if (n > 0) build(0, (int32_t)n);
  }
  int32_t build(int32_t b, int32_t e) {
    Node nd; nd.begin = b; nd.end = e; nd.left = nd.right = -1; nd.axis = 0; nd.split = 0;
    for (int a = 0; a < 3; ++a) { nd.lo[a] = std::numeric_limits<double>::infinity(); nd.hi[a] = -nd.lo[a]; }
    for (int32_t i = b; i < e; ++i)
      for (int a = 0; a < 3; ++a) {
        const double v = pts[3 * (int64_t)idx[i] + a];
        nd.lo[a] = std::min(nd.lo[a], v); nd.hi[a] = std::max(nd.hi[a], v);
      }
    const int32_t me = (int32_t)nodes.size();
    nodes.push_back(nd);
    if (e - b > kLeaf) {
      int ax = 0; double ext = nd.hi[0] - nd.lo[0];
      for (int a = 1; a < 3; ++a) if (nd.hi[a] - nd.lo[a] > ext) { ext = nd.hi[a] - nd.lo[a]; ax = a; }
      const int32_t mid = b + (e - b) / 2;
      std::nth_element(idx.begin() + b, idx.begin() + mid, idx.begin() + e, [&](int32_t x, int32_t y) {
        const double vx = pts[3 * (int64_t)x + ax], vy = pts[3 * (int64_t)y + ax];
        return vx < vy || (vx == vy && x < y);
      });
      const int32_t l = build(b, mid);
      const int32_t r = build(mid, e);
      nodes[me].left = l; nodes[me].right = r; nodes[me].axis = ax;
    }
    return me;
  }
  // lower bound with the same op order as dist_sq => monotone => safe strict pruning
  inline double box_lb(const Node& nd, const double* q) const {
    double d[3];
    for (int a = 0; a < 3; ++a) {
      const double lo = nd.lo[a] - q[a], hi = q[a] - nd.hi[a];
      d[a] = lo > 0 ? lo : (hi > 0 ? hi : 0.0);
    }
    return d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  }
  void search(int32_t ni, const double* q, double& best, int32_t& bi) const {
    const Node& nd = nodes[ni];
    if (nd.left < 0) {
      for (int32_t i = nd.begin; i < nd.end; ++i) {
        const int32_t j = idx[i];
        const double d = dist_sq(q, pts + 3 * (int64_t)j);
        if (d < best || (d == best && j < bi)) { best = d; bi = j; }
      }
      return;
    }
    const double ll = box_lb(nodes[nd.left], q), lr = box_lb(nodes[nd.right], q);
    const int32_t first = ll <= lr ? nd.left : nd.right, second = ll <= lr ? nd.right : nd.left;
    const double lf = ll <= lr ? ll : lr, ls = ll <= lr ? lr : ll;
    if (lf <= best) search(first, q, best, bi);
    if (ls <= best) search(second, q, best, bi);
  }
  // k nearest (ascending distance; equal distances by ascending index) -- Frame::getNeighbours' knnSearch (frame.cpp:208-225)
  void knn_rec(int32_t ni, const double* q, int k, std::vector<std::pair<double, int32_t>>& heap) const {
    const Node& nd = nodes[ni];
    const double worst = (int)heap.size() < k ? std::numeric_limits<double>::infinity() : heap.back().first;
    if (nd.left < 0) {
      for (int32_t i = nd.begin; i < nd.end; ++i) {
        const int32_t j = idx[i];
        const std::pair<double, int32_t> c{dist_sq(q, pts + 3 * (int64_t)j), j};
        if ((int)heap.size() < k || c < heap.back()) {
          heap.insert(std::upper_bound(heap.begin(), heap.end(), c), c);
          if ((int)heap.size() > k) heap.pop_back();
        }
      }
      return;
    }
    (void)worst;
    const double ll = box_lb(nodes[nd.left], q), lr = box_lb(nodes[nd.right], q);
    const int32_t first = ll <= lr ? nd.left : nd.right, second = ll <= lr ? nd.right : nd.left;
    const double lf = ll <= lr ? ll : lr, ls = ll <= lr ? lr : ll;
    if ((int)heap.size() < k || lf <= heap.back().first) knn_rec(first, q, k, heap);
    if ((int)heap.size() < k || ls <= heap.back().first) knn_rec(second, q, k, heap);
  }
  void query(const double* q, int32_t* out_idx, double* out_d2) const {
    double best = std::numeric_limits<double>::infinity(); int32_t bi = std::numeric_limits<int32_t>::max();
    if (n > 0) search(0, q, best, bi);
    *out_idx = bi; *out_d2 = best;
  }
};

static void brute_query(const double* pts, int64_t n, const double* q, int32_t* out_idx, double* out_d2) {
  double best = std::numeric_limits<double>::infinity(); int32_t bi = -1;
  for (int64_t j = 0; j < n; ++j) {
    const double d = dist_sq(q, pts + 3 * j);
    if (d < best) { best = d; bi = (int32_t)j; }   // strict: lowest index wins ties
  }
  *out_idx = bi; *out_d2 = best;
}

// =================================================================================================
// Cost functors (include/icp-ceres.h), restated on raw arrays; T = double or Jet.
// Ambient layouts: AA [wx wy wz tx ty tz]; QUAT/SE3 [qx qy qz qw tx ty tz].
// =================================================================================================
enum { PARAM_AA = 0, PARAM_QUAT = 1, PARAM_SE3 = 2 };
enum { COST_P2P = 0, COST_P2PLANE = 1, COST_MIXED = 2 };

template <typename T> static inline T dot3(const T* a, const T* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

// PointToPointErrorGlobal{,_CeresAngleAxis,_SophusSE3}  icp-ceres.h:49-99,143-185,236-275
template <int param, typename T> static inline void functor_p2p(const T* cam1, const T* cam2, const double* src,
                                                                const double* dst, T* r) {
  T s[3] = {T(src[0]), T(src[1]), T(src[2])}, d[3] = {T(dst[0]), T(dst[1]), T(dst[2])};
  T p[3], p2[3];
  if constexpr (param == PARAM_AA) {
    angle_axis_rotate_point(cam1, s, p); angle_axis_rotate_point(cam2, d, p2);
    for (int i = 0; i < 3; ++i) { p[i] += cam1[3 + i]; p2[i] += cam2[3 + i]; }
  } else {  // QUAT: q * src; p += t (:80-83).  SE3: q.unit_quaternion()*src + q.translation() (:261-262)
    quat_transform(cam1, s, p); quat_transform(cam2, d, p2);
    for (int i = 0; i < 3; ++i) { p[i] = p[i] + cam1[4 + i]; p2[i] = p2[i] + cam2[4 + i]; }
  }
  r[0] = p[0] - p2[0]; r[1] = p[1] - p2[1]; r[2] = p[2] - p2[2];
}

// PointToPlaneErrorGlobal{,_CeresAngleAxis,_SophusSE3}  icp-ceres.h:101-141,187-234,277-316
template <int param, typename T> static inline void functor_p2plane(const T* cam1, const T* cam2, const double* src,
                                                                    const double* dst, const double* nor, T* r) {
  T s[3] = {T(src[0]), T(src[1]), T(src[2])}, d[3] = {T(dst[0]), T(dst[1]), T(dst[2])};
  T n[3] = {T(nor[0]), T(nor[1]), T(nor[2])};
  T p[3], p2[3], n2[3];
  if constexpr (param == PARAM_AA) {
    angle_axis_rotate_point(cam1, s, p); angle_axis_rotate_point(cam2, d, p2); angle_axis_rotate_point(cam2, n, n2);
    for (int i = 0; i < 3; ++i) { p[i] += cam1[3 + i]; p2[i] += cam2[3 + i]; }
    r[0] = (p[0] - p2[0]) * n2[0] + (p[1] - p2[1]) * n2[1] + (p[2] - p2[2]) * n2[2];
  } else {
    quat_transform(cam1, s, p);
    for (int i = 0; i < 3; ++i) p[i] = p[i] + cam1[4 + i];
    if constexpr (param == PARAM_QUAT) {  // dst side goes through toRotationMatrix() (:132-134)
      T Rk[9]; quat_to_matrix(cam2, Rk);
      mat3_mul_vec(Rk, d, p2); mat3_mul_vec(Rk, n, n2);
    } else {
      quat_transform(cam2, d, p2); quat_transform(cam2, n, n2);
    }
    for (int i = 0; i < 3; ++i) p2[i] = p2[i] + cam2[4 + i];
    T df[3] = {p[0] - p2[0], p[1] - p2[1], p[2] - p2[2]};
    r[0] = dot3(df, n2);
  }
}

// ceres::SoftLOneLoss(a)::Evaluate [ext-knowledge Ceres 1.13 loss_function.cc]
static inline void soft_l1(double a, double s, double rho[3]) {
  const double b = a * a, c = 1.0 / b;
  const double sum = 1.0 + s * c, tmp = std::sqrt(sum);
  rho[0] = 2.0 * b * (tmp - 1.0);
  rho[1] = std::max(std::numeric_limits<double>::min(), 1.0 / tmp);
  rho[2] = -(c * rho[1]) / (2.0 * sum);
}

// =================================================================================================
// Multiview problem  (icp-ceres.cpp:220-475)
// =================================================================================================
struct MvProblem {
  int M = 0, param = PARAM_SE3, cost = COST_P2PLANE, robust = 1, se3_autodiff = 1, threads = 1;
  std::vector<const double*> pts, nor;
  std::vector<uint8_t> fixed;
  int E = 0;
  const int32_t *e_src = nullptr, *e_dst = nullptr, *first = nullptr, *second = nullptr;
  const int64_t* e_off = nullptr;
  const float* e_weight = nullptr;
  int G() const { return param == PARAM_AA ? 6 : 7; }
  std::vector<double> params;      // M*G ambient parameters (all frames, fixed ones included)
  std::vector<int> col;            // frame -> first local column, -1 if constant
  std::vector<int> xoff;           // frame -> offset in reduced x, -1 if constant
  int n_local = 0, n_ambient = 0;

  void pose_to_param(const double* P16, double* x) const {
    double R[9], t[3]; pose16_split(P16, R, t);
    if (param == PARAM_AA) { rotation_matrix_to_angle_axis(R, x); x[3] = t[0]; x[4] = t[1]; x[5] = t[2]; }
    else { quat_from_matrix(R, x); x[4] = t[0]; x[5] = t[1]; x[6] = t[2]; }  // no normalisation (so3.hpp:666-668, icp-ceres.cpp:236)
  }
  void param_to_pose(const double* x, double* P16) const {
    double R[9];
    if (param == PARAM_AA) { angle_axis_to_rotation_matrix(x, R); pose16_join(R, x + 3, P16); }
    else { quat_to_matrix(x, R); pose16_join(R, x + 4, P16); }   // eigenQuaternionToIso / sophusToIso
  }
  void frame_plus(const double* x, const double* d, double* out) const {
    if (param == PARAM_AA) { for (int i = 0; i < 6; ++i) out[i] = x[i] + d[i]; }
    else if (param == PARAM_QUAT) { eigen_quat_plus(x, d, out); for (int i = 0; i < 3; ++i) out[4 + i] = x[4 + i] + d[3 + i]; }
    else se3_plus(x, d, out);
  }
  // local parameterisation Jacobian, G x 6 row-major
  void frame_local_jac(const double* x, double* P) const {
    const int g = G();
    for (int i = 0; i < g * 6; ++i) P[i] = 0.0;
    if (param == PARAM_AA) { for (int i = 0; i < 6; ++i) P[i * 6 + i] = 1.0; }
    else if (param == PARAM_QUAT) {
      double J[12]; eigen_quat_jacobian(x, J);
      for (int r = 0; r < 4; ++r) for (int c = 0; c < 3; ++c) P[r * 6 + c] = J[r * 3 + c];
      for (int i = 0; i < 3; ++i) P[(4 + i) * 6 + 3 + i] = 1.0;
    } else if (se3_autodiff) {   // ceres::AutoDiffLocalParameterization<SophusSE3Plus,7,6> (sophus_se3.h:64-68)
      Jet<6> xj[7], dj[6], out[7];
      for (int i = 0; i < 7; ++i) xj[i] = Jet<6>(x[i]);
      for (int i = 0; i < 6; ++i) dj[i] = Jet<6>(0.0, i);
      se3_plus(xj, dj, out);
      for (int r = 0; r < 7; ++r) for (int c = 0; c < 6; ++c) P[r * 6 + c] = out[r].v[c];
    } else se3_internal_jacobian(x, P);
  }

  void reduced_to_params(const double* x, std::vector<double>& p) const {
    const int g = G(); p = params;
    for (int f = 0; f < M; ++f) if (xoff[f] >= 0) std::memcpy(&p[(size_t)f * g], x + xoff[f], sizeof(double) * g);
  }

  template <int GG, int PP> bool evaluate_t(const double* x, double* cost_out, double* H, double* g) const {
    const int n = n_local;
    std::vector<double> p; reduced_to_params(x, p);
    std::vector<double> P((size_t)M * GG * 6);
    if (H) for (int f = 0; f < M; ++f) frame_local_jac(&p[(size_t)f * GG], &P[(size_t)f * GG * 6]);
    int nth = std::max(1, threads);
    constexpr int CPAD = 16;   // one cache line (and its neighbour) per thread: the accumulators are written every iteration
    std::vector<double> costs((size_t)nth * CPAD, 0.0);
    std::vector<std::vector<double>> Hs, gs;
    if (H) { Hs.assign(nth, std::vector<double>((size_t)n * n, 0.0)); gs.assign(nth, std::vector<double>(n, 0.0)); }
    for (int e = 0; e < E; ++e) {
      const int s = e_src[e], k = e_dst[e];
      if (fixed[s]) continue;   // `if(srcCloud.fixed) continue;` icp-ceres.cpp:255,353,426
      const double a = robust ? (double)e_weight[e] : 0.0;   // SoftLOneLoss(dstEdge.weight) :284,374,449
      const double* xs = &p[(size_t)s * GG]; const double* xk = &p[(size_t)k * GG];
      const double* Ps = &P[(size_t)s * GG * 6]; const double* Pk = &P[(size_t)k * GG * 6];
      const int cs = col[s], ck = col[k];
      const int64_t c0 = e_off[e], c1 = e_off[e + 1];
#pragma omp parallel for num_threads(nth) schedule(static)
      for (int64_t c = c0; c < c1; ++c) {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        const double* ps = pts[s] + 3 * (int64_t)first[c];
        const double* pd = pts[k] + 3 * (int64_t)second[c];
        const double* nd = nor[k] ? nor[k] + 3 * (int64_t)second[c] : nullptr;
        // blocks: p2p (3 residuals) and/or p2plane (1 residual); each with its own loss instance
        for (int blk = 0; blk < 2; ++blk) {
          const bool is_plane = (blk == 1);
          if (is_plane && cost == COST_P2P) continue;
          if (!is_plane && cost == COST_P2PLANE) continue;
          const int nr = is_plane ? 1 : 3;
          double r[3]; double J[3][12];
          if (H) {
            Jet<2 * GG> c1j[GG], c2j[GG], rj[3];
            for (int i = 0; i < GG; ++i) { c1j[i] = Jet<2 * GG>(xs[i], i); c2j[i] = Jet<2 * GG>(xk[i], GG + i); }
            if (is_plane) functor_p2plane<PP>(c1j, c2j, ps, pd, nd, rj); else functor_p2p<PP>(c1j, c2j, ps, pd, rj);
            for (int q = 0; q < nr; ++q) {
              r[q] = rj[q].a;
              for (int l = 0; l < 6; ++l) {   // J_local = J_global * P   (Ceres applies the local parameterisation)
                double as = 0, ak = 0;
                for (int gi = 0; gi < GG; ++gi) { as += rj[q].v[gi] * Ps[gi * 6 + l]; ak += rj[q].v[GG + gi] * Pk[gi * 6 + l]; }
                J[q][l] = as; J[q][6 + l] = ak;
              }
            }
          } else {
            if (is_plane) functor_p2plane<PP>(xs, xk, ps, pd, nd, r); else functor_p2p<PP>(xs, xk, ps, pd, r);
          }
          double sq = 0; for (int q = 0; q < nr; ++q) sq += r[q] * r[q];
          if (robust) {
            double rho[3]; soft_l1(a, sq, rho);
            costs[(size_t)tid * CPAD] += 0.5 * rho[0];
            if (H) {   // Corrector with rho'' <= 0: scale residuals and Jacobian by sqrt(rho') [ext-knowledge corrector.cc]
              const double sr = std::sqrt(rho[1]);
              for (int q = 0; q < nr; ++q) { r[q] *= sr; for (int l = 0; l < 12; ++l) J[q][l] *= sr; }
            }
          } else costs[(size_t)tid * CPAD] += 0.5 * sq;
          if (H) {
            double* Ht = Hs[tid].data(); double* gt = gs[tid].data();
            for (int q = 0; q < nr; ++q) {
              for (int side_a = 0; side_a < 2; ++side_a) {
                const int ca = side_a ? ck : cs; if (ca < 0) continue;
                for (int la = 0; la < 6; ++la) {
                  const double ja = J[q][6 * side_a + la];
                  gt[ca + la] += ja * r[q];
                  for (int side_b = 0; side_b < 2; ++side_b) {
                    const int cb = side_b ? ck : cs; if (cb < 0) continue;
                    for (int lb = 0; lb < 6; ++lb) Ht[(size_t)(ca + la) * n + cb + lb] += ja * J[q][6 * side_b + lb];
                  }
                }
              }
            }
          }
        }
      }
    }
    double cst = 0; for (int t = 0; t < nth; ++t) cst += costs[(size_t)t * CPAD];
    *cost_out = cst;
    if (H) {
      std::fill(H, H + (size_t)n * n, 0.0); std::fill(g, g + n, 0.0);
      for (int t = 0; t < nth; ++t) {
        for (size_t i = 0; i < (size_t)n * n; ++i) H[i] += Hs[t][i];
        for (int i = 0; i < n; ++i) g[i] += gs[t][i];
      }
    }
    return std::isfinite(cst);
  }
  bool evaluate(const double* x, double* c, double* H, double* g) const {
    if (param == PARAM_AA) return evaluate_t<6, PARAM_AA>(x, c, H, g);
    if (param == PARAM_QUAT) return evaluate_t<7, PARAM_QUAT>(x, c, H, g);
    return evaluate_t<7, PARAM_SE3>(x, c, H, g);
  }
};

}  // namespace orc

// =================================================================================================
// C API (ctypes)
// =================================================================================================
using namespace orc;

extern "C" {

struct orc_lm_options {
  int32_t max_num_iterations;
  int32_t max_num_consecutive_invalid_steps;
  int32_t jacobi_scaling;
  int32_t _pad;
  double initial_trust_region_radius, max_trust_region_radius, min_trust_region_radius;
  double min_relative_decrease, min_lm_diagonal, max_lm_diagonal;
  double function_tolerance, gradient_tolerance, parameter_tolerance;
};
struct orc_lm_summary {
  int32_t termination, num_iterations, num_successful_steps, num_jacobian_evals, num_cost_evals, n_trace;
  double initial_cost, final_cost;
};

void orc_default_lm_options(orc_lm_options* o) {
  LmOptions d;
  o->max_num_iterations = d.max_num_iterations;
  o->max_num_consecutive_invalid_steps = d.max_num_consecutive_invalid_steps;
  o->jacobi_scaling = d.jacobi_scaling; o->_pad = 0;
  o->initial_trust_region_radius = d.initial_trust_region_radius;
  o->max_trust_region_radius = d.max_trust_region_radius;
  o->min_trust_region_radius = d.min_trust_region_radius;
  o->min_relative_decrease = d.min_relative_decrease;
  o->min_lm_diagonal = d.min_lm_diagonal; o->max_lm_diagonal = d.max_lm_diagonal;
  o->function_tolerance = d.function_tolerance; o->gradient_tolerance = d.gradient_tolerance;
  o->parameter_tolerance = d.parameter_tolerance;
}

static LmOptions to_opts(const orc_lm_options* o) {
  LmOptions d;
  if (!o) return d;
  d.max_num_iterations = o->max_num_iterations;
  d.max_num_consecutive_invalid_steps = o->max_num_consecutive_invalid_steps;
  d.jacobi_scaling = o->jacobi_scaling;
  d.initial_trust_region_radius = o->initial_trust_region_radius;
  d.max_trust_region_radius = o->max_trust_region_radius;
  d.min_trust_region_radius = o->min_trust_region_radius;
  d.min_relative_decrease = o->min_relative_decrease;
  d.min_lm_diagonal = o->min_lm_diagonal; d.max_lm_diagonal = o->max_lm_diagonal;
  d.function_tolerance = o->function_tolerance; d.gradient_tolerance = o->gradient_tolerance;
  d.parameter_tolerance = o->parameter_tolerance;
  return d;
}

// ---- correspondence ----------------------------------------------------------------------------
void* orc_kd_build(const double* pts, int64_t n) { return new KdTree(pts, n); }
void orc_kd_free(void* h) { delete (KdTree*)h; }
void orc_kd_query(void* h, const double* q, int32_t* idx, double* d2) { ((KdTree*)h)->query(q, idx, d2); }

// nearest neighbour of every transformed src point (frame.cpp:129-138). kd == NULL -> brute force.
void orc_closest_points(void* kd, const double* dst_pts, int64_t n_dst, const double* src_pts, int64_t n_src,
                        const double* pose_src16, const double* pose_dst16, int32_t* nn_idx, double* nn_d2,
                        double* query_out /*nullable n_src*3*/, int num_threads) {
  const EdgeXform X(pose_src16, pose_dst16);
  const KdTree* T = (const KdTree*)kd;
#pragma omp parallel for num_threads(num_threads > 0 ? num_threads : 1) schedule(static)
  for (int64_t k = 0; k < n_src; ++k) {
    double q[3]; X.apply(src_pts + 3 * k, q);
    if (query_out) { query_out[3 * k] = q[0]; query_out[3 * k + 1] = q[1]; query_out[3 * k + 2] = q[2]; }
    if (T) T->query(q, nn_idx + k, nn_d2 + k); else brute_query(dst_pts, n_dst, q, nn_idx + k, nn_d2 + k);
  }
}

// cutoff filter + median -> weight (frame.cpp:140-176). Returns inlier count; lists in ascending k.
int64_t orc_filter_edge(const int32_t* nn_idx, const double* nn_d2, int64_t n_src, float thresh, int32_t* first,
                        int32_t* second, double* dist, float* weight, double* median_out) {
  std::vector<double> dists;
  int64_t c = 0;
  for (int64_t k = 0; k < n_src; ++k) {
    const double pointDist = std::sqrt(nn_d2[k]);           // :142
    if (pointDist < thresh) {                               // :156 float promoted to double
      if (first) { first[c] = (int32_t)k; second[c] = nn_idx[k]; dist[c] = pointDist; }
      ++c; dists.push_back(pointDist);
    }
  }
  if (dists.empty()) {   // reference: UB (frame.cpp:166-168). Oracle choice: weight = 0, median = NaN.
    if (weight) *weight = 0.0f;
    if (median_out) *median_out = std::numeric_limits<double>::quiet_NaN();
    return 0;
  }
  auto middle = dists.begin() + (dists.size() / 2);         // :166
  std::nth_element(dists.begin(), middle, dists.end());     // :167
  const double nth = *middle;
  if (median_out) *median_out = nth;
  if (weight) *weight = (float)(nth * 1.5);                 // :176 (double product narrowed to float)
  return c;
}

// Frame::computePoseNeighboursKnn (frame.cpp:67-89): knn nearest frames by float |dt|.
// out_dst: M*knn (-1 padded). Order = partial_sort order by weight (stable tie rule: lowest j).
void orc_pose_graph_knn(int M, const double* poses16, int knn, int32_t* out_dst, float* out_w) {
  for (int i = 0; i < M; ++i) {
    std::vector<std::pair<float, int>> nb;
    for (int j = 0; j < M; ++j) {
      if (i == j) continue;
      const double* a = poses16 + 16 * i + 12; const double* b = poses16 + 16 * j + 12;
      const double d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
      nb.push_back({(float)std::sqrt(d0 * d0 + d1 * d1 + d2 * d2), j});
    }
    std::stable_sort(nb.begin(), nb.end(), [](const std::pair<float, int>& x, const std::pair<float, int>& y) { return x.first < y.first; });
    for (int q = 0; q < knn; ++q) {
      out_dst[i * knn + q] = q < (int)nb.size() ? nb[q].second : -1;
      if (out_w) out_w[i * knn + q] = q < (int)nb.size() ? nb[q].first : 0.f;
    }
  }
}

// k nearest neighbours of a query in the index's own frame
void orc_kd_knn(void* h, const double* q, int k, int32_t* idx, double* d2) {
  std::vector<std::pair<double, int32_t>> heap;
  const KdTree* T = (const KdTree*)h;
  if (T->n > 0) T->knn_rec(0, q, k, heap);
  for (int i = 0; i < k; ++i) { idx[i] = i < (int)heap.size() ? heap[i].second : -1; d2[i] = i < (int)heap.size() ? heap[i].first : 0.0; }
}

// symmetric 3x3 eigen-decomposition (cyclic Jacobi); eigenvalues ascending, eigenvectors in the columns of V
static void eig3(double A[3][3], double w[3], double V[3][3]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 50; ++sweep) {
    const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        const double th = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
        for (int k = 0; k < 3; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
        for (int k = 0; k < 3; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
      }
  }
  int o[3] = {0, 1, 2};
  std::sort(o, o + 3, [&](int a, int b) { return A[a][a] < A[b][b]; });
  double Vs[3][3];
  for (int j = 0; j < 3; ++j) { w[j] = A[o[j]][o[j]]; for (int i = 0; i < 3; ++i) Vs[i][j] = V[i][o[j]]; }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = Vs[i][j];
}

// Frame::recomputeNormals (frame.cpp:244-255) -> getNeighbours(i, k) (:208-242, the point itself included) ->
// pointSetPCA (common.h:331-346): centroid, cov = sum (p - c)(p - c)^T (not divided by n), normal = eigenvector of the
// smallest eigenvalue (SelfAdjointEigenSolver [ext-knowledge Eigen]: unit norm, sign arbitrary), flipped so that n.z <= 0.
// nn_out (nullable): the k neighbour indices per point, in knnSearch order.
void orc_recompute_normals(const double* pts, int64_t n, int k, double* nor_out, int32_t* nn_out, int num_threads) {
  KdTree T(pts, n);
#pragma omp parallel for num_threads(num_threads > 0 ? num_threads : 1) schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    std::vector<std::pair<double, int32_t>> heap;
    T.knn_rec(0, pts + 3 * i, k, heap);
    const int m = (int)heap.size();
    double c[3] = {0, 0, 0};
    for (int j = 0; j < m; ++j) for (int a = 0; a < 3; ++a) c[a] += pts[3 * (int64_t)heap[j].second + a];
    for (int a = 0; a < 3; ++a) c[a] /= m;
    double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int j = 0; j < m; ++j) {
      double d[3]; for (int a = 0; a < 3; ++a) d[a] = pts[3 * (int64_t)heap[j].second + a] - c[a];
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) C[a][b] += d[a] * d[b];
    }
    double w[3], V[3][3]; eig3(C, w, V);
    double nv[3] = {V[0][0], V[1][0], V[2][0]};
    if (nv[2] > 0) { nv[0] = -nv[0]; nv[1] = -nv[1]; nv[2] = -nv[2]; }   // "flip towards camera" common.h:343
    for (int a = 0; a < 3; ++a) nor_out[3 * i + a] = nv[a];
    if (nn_out) for (int j = 0; j < k; ++j) nn_out[(int64_t)k * i + j] = j < m ? heap[j].second : -1;
  }
}

// ---- LM ------------------------------------------------------------------------------------------
int orc_optimize(int M, const double* const* pts, const double* const* nor, const uint8_t* fixed, double* poses16, int E,
                 const int32_t* e_src, const int32_t* e_dst, const int64_t* e_off, const int32_t* first,
                 const int32_t* second, const float* e_weight, int param, int cost, int robust, int se3_autodiff,
                 int num_threads, const orc_lm_options* opt, orc_lm_summary* summary, double* trace, int max_trace) {
  MvProblem pr;
  pr.M = M; pr.param = param; pr.cost = cost; pr.robust = robust; pr.se3_autodiff = se3_autodiff;
  pr.threads = num_threads > 0 ? num_threads : 1;
  pr.pts.assign(pts, pts + M); pr.nor.assign(nor, nor + M);
  pr.fixed.assign(M, 0);
  for (int f = 0; f < M; ++f) pr.fixed[f] = (fixed && fixed[f]) ? 1 : 0;
  pr.fixed[0] = 1;   // frames[0]->fixed = true (icp-ceres.cpp:242-244,342-344,417-419)
  pr.E = E; pr.e_src = e_src; pr.e_dst = e_dst; pr.e_off = e_off; pr.first = first; pr.second = second; pr.e_weight = e_weight;
  const int G = pr.G();
  pr.params.resize((size_t)M * G);
  for (int f = 0; f < M; ++f) pr.pose_to_param(poses16 + 16 * f, &pr.params[(size_t)f * G]);
  pr.col.assign(M, -1); pr.xoff.assign(M, -1);
  for (int f = 0; f < M; ++f) if (!pr.fixed[f]) { pr.col[f] = pr.n_local; pr.n_local += 6; pr.xoff[f] = pr.n_ambient; pr.n_ambient += G; }
  std::vector<double> x(pr.n_ambient);
  for (int f = 0; f < M; ++f) if (pr.xoff[f] >= 0) std::memcpy(&x[pr.xoff[f]], &pr.params[(size_t)f * G], sizeof(double) * G);

  LmModel model; model.n_local = pr.n_local; model.n_ambient = pr.n_ambient;
  model.evaluate = [&](const double* xx, double* c, double* H, double* g) { return pr.evaluate(xx, c, H, g); };
  model.plus = [&](const double* xx, const double* d, double* out) {
    for (int f = 0; f < M; ++f) if (pr.xoff[f] >= 0) pr.frame_plus(xx + pr.xoff[f], d + pr.col[f], out + pr.xoff[f]);
  };
  LmSummary s;
  if (pr.n_local > 0) s = lm_minimize(model, to_opts(opt), x);
  std::vector<double> p; pr.reduced_to_params(x.data(), p);
  for (int f = 0; f < M; ++f) pr.param_to_pose(&p[(size_t)f * G], poses16 + 16 * f);   // ALL frames rewritten (:318-322,392-394,472-474)
  if (summary) {
    summary->termination = s.termination; summary->num_iterations = s.num_iterations;
    summary->num_successful_steps = s.num_successful_steps; summary->num_jacobian_evals = s.num_jacobian_evals;
    summary->num_cost_evals = s.num_cost_evals; summary->initial_cost = s.initial_cost; summary->final_cost = s.final_cost;
    summary->n_trace = (int32_t)s.trace.size();
  }
  if (trace)
    for (int i = 0; i < (int)s.trace.size() && i < max_trace; ++i) {
      const LmIterationRecord& r = s.trace[i]; double* t = trace + 10 * i;
      t[0] = r.iteration; t[1] = r.step_valid; t[2] = r.step_accepted; t[3] = r.cost; t[4] = r.candidate_cost;
      t[5] = r.model_cost_change; t[6] = r.relative_decrease; t[7] = r.radius; t[8] = r.step_norm; t[9] = r.gradient_max_norm;
    }
  return 0;
}

// cost (+ unscaled local H, g) at the given poses: lets tests compare a single evaluation.
int orc_evaluate(int M, const double* const* pts, const double* const* nor, const uint8_t* fixed, const double* poses16,
                 int E, const int32_t* e_src, const int32_t* e_dst, const int64_t* e_off, const int32_t* first,
                 const int32_t* second, const float* e_weight, int param, int cost, int robust, int se3_autodiff,
                 int num_threads, double* cost_out, double* H /*nullable (6*(#free))^2*/, double* g) {
  MvProblem pr;
  pr.M = M; pr.param = param; pr.cost = cost; pr.robust = robust; pr.se3_autodiff = se3_autodiff;
  pr.threads = num_threads > 0 ? num_threads : 1;
  pr.pts.assign(pts, pts + M); pr.nor.assign(nor, nor + M);
  pr.fixed.assign(M, 0);
  for (int f = 0; f < M; ++f) pr.fixed[f] = (fixed && fixed[f]) ? 1 : 0;
  pr.fixed[0] = 1;
  pr.E = E; pr.e_src = e_src; pr.e_dst = e_dst; pr.e_off = e_off; pr.first = first; pr.second = second; pr.e_weight = e_weight;
  const int G = pr.G();
  pr.params.resize((size_t)M * G);
  for (int f = 0; f < M; ++f) pr.pose_to_param(poses16 + 16 * f, &pr.params[(size_t)f * G]);
  pr.col.assign(M, -1); pr.xoff.assign(M, -1);
  for (int f = 0; f < M; ++f) if (!pr.fixed[f]) { pr.col[f] = pr.n_local; pr.n_local += 6; pr.xoff[f] = pr.n_ambient; pr.n_ambient += G; }
  std::vector<double> x(pr.n_ambient);
  for (int f = 0; f < M; ++f) if (pr.xoff[f] >= 0) std::memcpy(&x[pr.xoff[f]], &pr.params[(size_t)f * G], sizeof(double) * G);
  return pr.evaluate(x.data(), cost_out, H, g) ? 0 : 1;
}

// Pairwise solvers (icp-ceres.cpp:137-218,525-565): one pose from identity, no loss, same solve().
// Equivalent to the 2-frame multiview problem with the dst frame constant at identity (the Global
// functors reduce exactly to the pairwise ones when cam2 is the identity: q*v == v, R(0)v == v).
int orc_pairwise(const double* src, const double* dst, const double* nor, int64_t n, int param, int cost,
                 int se3_autodiff, int num_threads, const orc_lm_options* opt, double* pose16_out,
                 orc_lm_summary* summary) {
  std::vector<int32_t> first(n), second(n);
  for (int64_t i = 0; i < n; ++i) first[i] = second[i] = (int32_t)i;
  const double* pts[2] = {dst, src}; const double* nr[2] = {nor, nullptr};
  double poses[32]; for (int i = 0; i < 32; ++i) poses[i] = 0;
  for (int f = 0; f < 2; ++f) for (int i = 0; i < 4; ++i) poses[16 * f + 5 * i] = 1.0;
  const int32_t es[1] = {1}, ed[1] = {0}; const int64_t off[2] = {0, n}; const float w[1] = {0.f};
  const uint8_t fx[2] = {1, 0};
  int rc = orc_optimize(2, pts, nr, fx, poses, 1, es, ed, off, first.data(), second.data(), w, param, cost, 0, se3_autodiff,
                        num_threads, opt, summary, nullptr, 0);
  std::memcpy(pose16_out, poses + 16, sizeof(double) * 16);
  return rc;
}

// ---- closed-form pairwise solvers (src/internal/icp-closedform.cpp:9-54) -------
// pointToPoint (:9-26): centroids, K = sum (q - qbar)(p - pbar)^T, R = U V^T from the SVD of K, and -- as the reference does
// it -- `R.col(2) *= -1` when det R < 0 (the third column of R, not of U); t = qbar - R pbar.
// U V^T is the orthogonal polar factor of K, unique for non-singular K whatever the SVD's sign/order conventions are
// [ext-knowledge Eigen JacobiSVD], computed here as K V diag(1/sigma) V^T with V, sigma^2 from the symmetric
// eigen-decomposition of K^T K.
void orc_closed_p2p(const double* src, const double* dst, int64_t n, double* pose16_out) {
  double pb[3] = {0, 0, 0}, qb[3] = {0, 0, 0};
  for (int64_t i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) { pb[a] += src[3 * i + a]; qb[a] += dst[3 * i + a]; }
  for (int a = 0; a < 3; ++a) { pb[a] /= (double)n; qb[a] /= (double)n; }
  double K[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int64_t i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) K[a][b] += (dst[3 * i + a] - qb[a]) * (src[3 * i + b] - pb[b]);
  double S[3][3], w[3], V[3][3];
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { S[a][b] = 0; for (int c = 0; c < 3; ++c) S[a][b] += K[c][a] * K[c][b]; }
  eig3(S, w, V);
  // U column by column, largest singular value first; a (numerically) vanishing singular value leaves its column of U
  // free up to orthonormality, as in JacobiSVD's full U: complete the basis with a cross product (rank 2) or an arbitrary
  // perpendicular (rank 1) instead of dividing by ~0.
  int o[3] = {0, 1, 2};
  for (int i = 0; i < 3; ++i) for (int j = i + 1; j < 3; ++j) if (w[o[j]] > w[o[i]]) std::swap(o[i], o[j]);
  double Uc[3][3], Vc[3][3], sig[3];
  for (int j = 0; j < 3; ++j) {
    sig[j] = std::sqrt(std::max(w[o[j]], 0.0));
    for (int a = 0; a < 3; ++a) { Vc[j][a] = V[a][o[j]]; Uc[j][a] = 0; for (int c = 0; c < 3; ++c) Uc[j][a] += K[a][c] * V[c][o[j]]; }
  }
  auto unit = [](double* v) { const double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); if (n > 0) { v[0] /= n; v[1] /= n; v[2] /= n; } return n; };
  if (unit(Uc[0]) == 0) { Uc[0][0] = 1; Uc[0][1] = Uc[0][2] = 0; }
  {
    const double d = Uc[1][0] * Uc[0][0] + Uc[1][1] * Uc[0][1] + Uc[1][2] * Uc[0][2];
    for (int a = 0; a < 3; ++a) Uc[1][a] -= d * Uc[0][a];
    const double n1 = std::sqrt(Uc[1][0] * Uc[1][0] + Uc[1][1] * Uc[1][1] + Uc[1][2] * Uc[1][2]);
    if (n1 > 1e-13 * sig[0] && n1 > 0) unit(Uc[1]);
    else {
      int m = 0; for (int a = 1; a < 3; ++a) if (std::fabs(Uc[0][a]) < std::fabs(Uc[0][m])) m = a;
      for (int a = 0; a < 3; ++a) Uc[1][a] = (a == m ? 1.0 : 0.0) - Uc[0][m] * Uc[0][a];
      unit(Uc[1]);
    }
  }
  {
    const double c[3] = {Uc[0][1] * Uc[1][2] - Uc[0][2] * Uc[1][1], Uc[0][2] * Uc[1][0] - Uc[0][0] * Uc[1][2], Uc[0][0] * Uc[1][1] - Uc[0][1] * Uc[1][0]};
    // sigma_3 = c . K v_3 must be >= 0; if it vanishes both signs are valid SVDs and the proper rotation is taken
    const double s3 = Uc[2][0] * c[0] + Uc[2][1] * c[1] + Uc[2][2] * c[2];
    const double dV = Vc[0][0] * (Vc[1][1] * Vc[2][2] - Vc[1][2] * Vc[2][1]) - Vc[0][1] * (Vc[1][0] * Vc[2][2] - Vc[1][2] * Vc[2][0]) +
                      Vc[0][2] * (Vc[1][0] * Vc[2][1] - Vc[1][1] * Vc[2][0]);
    const double sgn = (std::fabs(s3) > 1e-13 * sig[0] ? s3 < 0 : dV < 0) ? -1.0 : 1.0;
    for (int a = 0; a < 3; ++a) Uc[2][a] = sgn * c[a];
  }
  double R[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int j = 0; j < 3; ++j) for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) R[a][b] += Uc[j][a] * Vc[j][b];
  const double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                     R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
  if (det < 0) for (int a = 0; a < 3; ++a) R[a][2] *= -1;      // icp-closedform.cpp:20-22
  for (int i = 0; i < 16; ++i) pose16_out[i] = 0; pose16_out[15] = 1;
  for (int a = 0; a < 3; ++a) {
    for (int b = 0; b < 3; ++b) pose16_out[4 * b + a] = R[a][b];
    pose16_out[12 + a] = qb[a] - (R[a][0] * pb[0] + R[a][1] * pb[1] + R[a][2] * pb[2]);
  }
}

// pointToPlane (:30-54): linearised rotation, 6x6 normal equations C x = d with x = (alpha, beta, gamma, t),
// rows [p x n ; n], d = -sum [p x n ; n] ((p - q).n), solved by LDL^T; R = Rx(alpha) Ry(beta) Rz(gamma), t = x[3..5].
void orc_closed_p2plane(const double* src, const double* dst, const double* nor, int64_t n, double* pose16_out) {
  double Cm[6][6] = {{0}}, d[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t i = 0; i < n; ++i) {
    const double* p = src + 3 * i; const double* q = dst + 3 * i; const double* nn = nor + 3 * i;
    const double a[6] = {p[1] * nn[2] - p[2] * nn[1], p[2] * nn[0] - p[0] * nn[2], p[0] * nn[1] - p[1] * nn[0], nn[0], nn[1], nn[2]};
    const double sum = (p[0] - q[0]) * nn[0] + (p[1] - q[1]) * nn[1] + (p[2] - q[2]) * nn[2];
    for (int r = 0; r < 6; ++r) { for (int c = 0; c < 6; ++c) Cm[r][c] += a[r] * a[c]; d[r] -= a[r] * sum; }
  }
  // LDL^T without pivoting (C is symmetric positive definite for a non-degenerate surface)
  double L[6][6] = {{0}}, D[6];
  for (int j = 0; j < 6; ++j) {
    double dj = Cm[j][j]; for (int k = 0; k < j; ++k) dj -= L[j][k] * L[j][k] * D[k];
    D[j] = dj; L[j][j] = 1;
    for (int i = j + 1; i < 6; ++i) { double v = Cm[i][j]; for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k] * D[k]; L[i][j] = v / dj; }
  }
  double y[6], x[6];
  for (int i = 0; i < 6; ++i) { y[i] = d[i]; for (int k = 0; k < i; ++k) y[i] -= L[i][k] * y[k]; }
  for (int i = 5; i >= 0; --i) { x[i] = y[i] / D[i]; for (int k = i + 1; k < 6; ++k) x[i] -= L[k][i] * x[k]; }
  const double ca = std::cos(x[0]), sa = std::sin(x[0]), cb = std::cos(x[1]), sb = std::sin(x[1]), cg = std::cos(x[2]), sg = std::sin(x[2]);
  const double R[3][3] = {{cb * cg, -cb * sg, sb},
                          {sa * sb * cg + ca * sg, -sa * sb * sg + ca * cg, -sa * cb},
                          {-ca * sb * cg + sa * sg, ca * sb * sg + sa * cg, ca * cb}};
  for (int i = 0; i < 16; ++i) pose16_out[i] = 0; pose16_out[15] = 1;
  for (int a = 0; a < 3; ++a) { for (int b = 0; b < 3; ++b) pose16_out[4 * b + a] = R[a][b]; pose16_out[12 + a] = x[3 + a]; }
}

// ---- small math exports for known-answer tests ---------------------------------------------------
void orc_se3_exp(const double* tangent6, double* out7) { se3_exp(tangent6, out7); }
void orc_se3_mul(const double* a7, const double* b7, double* out7) { se3_mul(a7, b7, out7); }
void orc_se3_plus(const double* x7, const double* d6, double* out7) { se3_plus(x7, d6, out7); }
void orc_se3_internal_jacobian(const double* x7, double* jac42) { se3_internal_jacobian(x7, jac42); }
void orc_se3_plus_jacobian_autodiff(const double* x7, double* jac42) {
  MvProblem pr; pr.param = PARAM_SE3; pr.se3_autodiff = 1; pr.frame_local_jac(x7, jac42);
}
void orc_quat_from_matrix(const double* R9, double* q4) { quat_from_matrix(R9, q4); }
void orc_quat_to_matrix(const double* q4, double* R9) { quat_to_matrix(q4, R9); }
void orc_quat_transform(const double* q4, const double* v, double* out) { quat_transform(q4, v, out); }
void orc_quat_plus(const double* x4, const double* d3, double* out4) { eigen_quat_plus(x4, d3, out4); }
void orc_quat_jacobian(const double* x4, double* jac12) { eigen_quat_jacobian(x4, jac12); }
// One restated multiview functor (functor_p2p / functor_p2plane) differentiated with Jets as evaluate_t does: residuals r[NR] and
// the ambient Jacobian jac[NR][2G] (G = 6 for aa, 7 for quat / se3).  Lets tests/test_oracle_functor_pin.py compare the
// restatement with the reference's own functor text compiled into oracle/_ref/libref_functors.so (oracle/ref_functors.cpp).
void orc_functor_eval(int param, int plane, const double* cam1, const double* cam2, const double* src, const double* dst,
                      const double* nor, double* r, double* jac) {
  auto run = [&](auto tag_g, auto tag_p) {
    constexpr int GG = decltype(tag_g)::value; constexpr int PP = decltype(tag_p)::value;
    Jet<2 * GG> c1[GG], c2[GG], rj[3];
    for (int i = 0; i < GG; ++i) { c1[i] = Jet<2 * GG>(cam1[i], i); c2[i] = Jet<2 * GG>(cam2[i], GG + i); }
    if (plane) functor_p2plane<PP>(c1, c2, src, dst, nor, rj); else functor_p2p<PP>(c1, c2, src, dst, rj);
    const int nr = plane ? 1 : 3;
    for (int q = 0; q < nr; ++q) { r[q] = rj[q].a; for (int k = 0; k < 2 * GG; ++k) jac[q * 2 * GG + k] = rj[q].v[k]; }
  };
  if (param == PARAM_AA) run(std::integral_constant<int, 6>(), std::integral_constant<int, PARAM_AA>());
  else if (param == PARAM_QUAT) run(std::integral_constant<int, 7>(), std::integral_constant<int, PARAM_QUAT>());
  else run(std::integral_constant<int, 7>(), std::integral_constant<int, PARAM_SE3>());
}
void orc_angle_axis_rotate(const double* aa, const double* p, double* out) { angle_axis_rotate_point(aa, p, out); }
void orc_rotmat_to_angle_axis(const double* R9, double* aa) { rotation_matrix_to_angle_axis(R9, aa); }
void orc_angle_axis_to_rotmat(const double* aa, double* R9) { angle_axis_to_rotation_matrix(aa, R9); }
void orc_mat3_inverse(const double* m, double* inv) { mat3_inverse_cofactor(m, inv); }
void orc_pose_to_param(const double* P16, int param, double* x) { MvProblem pr; pr.param = param; pr.pose_to_param(P16, x); }
void orc_param_to_pose(const double* x, int param, double* P16) { MvProblem pr; pr.param = param; pr.param_to_pose(x, P16); }
void orc_edge_transform(const double* pose_src16, const double* pose_dst16, const double* p, double* q) {
  EdgeXform X(pose_src16, pose_dst16); X.apply(p, q);
}
int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
