// tools/knn_host_check.cpp -- compiles the engine's own csrc/knn.cuh + tree_build.h with g++ (tools/hostshim) and runs the
// SEARCH TEXT (nn_query_init / nn_search, with and without the per-leaf neighbour lists of adjacency.h) on the host against a brute force in the
// reference's operation order, or (--file <bin>) against a golden answer.  Usage: knn_host_check <n_points> <n_queries> <seed> <mode>   mode 0: fp32-exact coordinates
// (fp32 storage), 1: arbitrary doubles (fp64 records + rounded fp32 screening copy).  Exit code 0 = all queries exact.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>
#include "../mv_lm_icp_b200/csrc/knn.cuh"
#include "../mv_lm_icp_b200/csrc/tree_build.h"

// frame from explicit points (fp64 records + rounded fp32 screening copy, as mvicp_set_frames stores non-fp32 data)
struct HostFrame {
  HostFrameBuild hb; std::vector<float4> sf; std::vector<double4a> sd; FrameDev fd{};
  HostFrame(const double* pts, int n, bool f32) {
    build_frame(pts, n, hb);
    const int64_t npad = ((n + LEAF - 1) / LEAF) * LEAF;
    sf.resize(npad); sd.resize(npad);
    for (int64_t i = 0; i < npad; ++i) {
      const int32_t w = i < n ? hb.order[i] : INT32_MAX; const int64_t j = i < n ? hb.order[i] : 0;
      float4 r; double4a q;
      if (i < n) { r.x = (float)pts[3 * j]; r.y = (float)pts[3 * j + 1]; r.z = (float)pts[3 * j + 2]; q.x = pts[3 * j]; q.y = pts[3 * j + 1]; q.z = pts[3 * j + 2]; }
      else { r.x = r.y = r.z = INFINITY; q.x = q.y = q.z = INFINITY; }
      std::memcpy(&r.w, &w, 4); const long long wl = w; std::memcpy(&q.w, &wl, 8);
      sf[i] = r; sd[i] = q;
    }
    fd.pts_s = f32 ? (const void*)sf.data() : (const void*)sd.data(); fd.pts_sf = sf.data(); fd.boxes = hb.boxes.data(); fd.faces = hb.faces.data();
    fd.pos_of = hb.pos_of.data(); fd.n = n; fd.n_leaf_pad = hb.n_leaf_pad; fd.depth = hb.depth; fd.absmax = hb.absmax;
  }
};

// file mode: int64 n, int64 nq, pts[3n], queries[3nq] (already in the frame's coordinates), want_idx[nq] (int32), want_d2[nq]:
// a golden answer (e.g. the reference's nanoflann on a real scan, tests/test_knn_host.py) instead of the local brute force
static int run_file(const char* path) {
  FILE* f = std::fopen(path, "rb"); if (!f) { std::printf("cannot open %s\n", path); return 1; }
  int64_t n = 0, nq = 0; if (std::fread(&n, 8, 1, f) != 1 || std::fread(&nq, 8, 1, f) != 1) return 1;
  std::vector<double> pts(3 * n), q(3 * nq), wd(nq); std::vector<int32_t> wi(nq);
  if (std::fread(pts.data(), 8, 3 * n, f) != (size_t)(3 * n) || std::fread(q.data(), 8, 3 * nq, f) != (size_t)(3 * nq) ||
      std::fread(wi.data(), 4, nq, f) != (size_t)nq || std::fread(wd.data(), 8, nq, f) != (size_t)nq) return 1;
  std::fclose(f);
  HostFrame F(pts.data(), (int)n, false);
  int bad = 0; std::vector<int> prev(nq, -1);
  for (int pass = 0; pass < 2; ++pass)        // pass 1 is seeded with pass 0's answers, as round r+1 is by round r
    for (int64_t i = 0; i < nq; ++i)
      for (int sched = 0; sched < 2; ++sched) {
        NNQuery s2; nn_query_init(s2, q[3 * i], q[3 * i + 1], q[3 * i + 2], F.fd.absmax);
        const int start_leaf = pass == 0 ? -1 : F.hb.pos_of[prev[i]] / LEAF;
        FrameDev fdx = F.fd; fdx.adj = sched == 0 ? F.hb.adj.data() : nullptr;   // with / without the neighbour lists
        nn_search<false, NNQuery>(fdx, s2, start_leaf);
        if (s2.bi != wi[i] || s2.best != wd[i]) { if (++bad < 10) std::printf("MISMATCH q %lld pass %d sched %d: got (%d, %.17g) want (%d, %.17g)\n", (long long)i, pass, sched, s2.bi, s2.best, wi[i], wd[i]); }
        prev[i] = s2.bi;
      }
  std::printf("file %s: n %lld queries %lld: %d mismatches\n", path, (long long)n, (long long)nq, bad);
  return bad;
}

template <bool F32> static int run(int n, int nq, unsigned seed) {
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> U(-1.0, 1.0); std::normal_distribution<double> G(0.0, 1.0);
  // a wavy surface patch (clouds of the workload are 2-manifolds), with a few exact duplicates (distance ties)
  std::vector<double> pts(3 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    double x = 0.1 * U(rng), y = 0.1 * U(rng), z = 0.45 + 0.02 * std::sin(40 * x) * std::cos(25 * y) + 1e-4 * G(rng);
    if (F32) { x = (float)x; y = (float)y; z = (float)z; }
    pts[3 * i] = x; pts[3 * i + 1] = y; pts[3 * i + 2] = z;
  }
  for (int d = 0; d < n / 50; ++d) { const int a = rng() % n, b = rng() % n; for (int k = 0; k < 3; ++k) pts[3 * a + k] = pts[3 * b + k]; }
  HostFrameBuild hb; build_frame(pts.data(), n, hb);
  const int64_t npad = ((n + LEAF - 1) / LEAF) * LEAF;
  std::vector<float4> sf(npad); std::vector<double4a> sd(npad);
  for (int64_t i = 0; i < npad; ++i) {
    const int32_t w = i < n ? hb.order[i] : INT32_MAX; const int64_t j = i < n ? hb.order[i] : 0;
    float4 r; double4a q;
    if (i < n) { r.x = (float)pts[3 * j]; r.y = (float)pts[3 * j + 1]; r.z = (float)pts[3 * j + 2]; q.x = pts[3 * j]; q.y = pts[3 * j + 1]; q.z = pts[3 * j + 2]; }
    else { r.x = r.y = r.z = INFINITY; q.x = q.y = q.z = INFINITY; }
    std::memcpy(&r.w, &w, 4); const long long wl = w; std::memcpy(&q.w, &wl, 8);
    sf[i] = r; sd[i] = q;
  }
  FrameDev fd{};
  fd.pts_s = F32 ? (const void*)sf.data() : (const void*)sd.data(); fd.pts_sf = sf.data(); fd.boxes = hb.boxes.data(); fd.faces = hb.faces.data();
  fd.pos_of = hb.pos_of.data(); fd.n = n; fd.n_leaf_pad = hb.n_leaf_pad; fd.depth = hb.depth; fd.absmax = hb.absmax;
  int bad = 0; long n_cert = 0, n_pos = 0, n_loose = 0;
  for (int qi = 0; qi < nq; ++qi) {
    // queries: near the surface, far from it, exactly on a point, and outside the bounding box
    const int base = rng() % n; const int kind = qi % 4;
    const double s = kind == 0 ? 1e-4 : (kind == 1 ? 2e-2 : (kind == 2 ? 0.0 : 0.5));
    const double q[3] = {pts[3 * base] + s * G(rng), pts[3 * base + 1] + s * G(rng), pts[3 * base + 2] + s * G(rng)};
    double best = INFINITY, second = INFINITY; int bi = INT32_MAX;   // brute force, frame.h:70-76 operation order, lowest index on ties
    for (int i = 0; i < n; ++i) {
      const double d0 = q[0] - pts[3 * i], d1 = q[1] - pts[3 * i + 1], d2 = q[2] - pts[3 * i + 2];
      const double d = (d0 * d0 + d1 * d1) + d2 * d2;
      if (d < best) { second = best; best = d; bi = i; } else if (d < second) second = d;
    }
    // seeds: none, the right leaf, a random (stale) leaf
    for (int sk = 0; sk < 3; ++sk) {
      const int start_leaf = sk == 0 ? -1 : (sk == 1 ? hb.pos_of[bi] / LEAF : (int)(rng() % ((n + LEAF - 1) / LEAF)));
      for (int sched = 0; sched < 2; ++sched) {
        NNQuery s2; nn_query_init(s2, q[0], q[1], q[2], fd.absmax);
        FrameDev fdx = fd; fdx.adj = sched == 0 ? hb.adj.data() : nullptr;
        nn_search<F32, NNQuery>(fdx, s2, start_leaf);
        if (s2.bi != bi || s2.best != best) {
          if (++bad < 10) std::printf("MISMATCH q %d kind %d seed-kind %d sched %d: got (%d, %.17g) want (%d, %.17g)\n", qi, kind, sk, sched, s2.bi, s2.best, bi, best);
        }
        // the certificate of the same search (knn.cuh, CERT): same answer, and its margin never exceeds the true gap to the runner-up
        NNQueryT st; nn_query_init(st, q[0], q[1], q[2], fd.absmax); nn_track_init(st);
        nn_search<F32, NNQueryT, true>(fdx, st, start_leaf);
        const float m = nn_margin(st);
        const double gap = std::sqrt(second) - std::sqrt(best);
        if (st.bi != bi || st.best != best || (double)m > gap) {
          if (++bad < 10) std::printf("CERTIFICATE q %d kind %d seed-kind %d sched %d: got (%d, %.17g, margin %.9g) want (%d, %.17g, gap %.9g)\n", qi, kind, sk, sched, st.bi, st.best, m, bi, best, gap);
        }
        if (sk == 1 && kind == 0) { ++n_cert; if (m > 0.f) ++n_pos;  if (gap > 4e-6 && (double)m < 0.5 * gap - 2e-6) ++n_loose; }
      }
    }
  }
  std::printf("n %d queries %d storage %s: %d mismatches; certificates of near-surface queries seeded with the right leaf: %ld, margin > 0: %ld, margin below half the true gap: %ld\n",
              n, nq, F32 ? "fp32" : "fp64", bad, n_cert, n_pos, n_loose);
  return bad;
}

int main(int argc, char** argv) {
  if (argc > 2 && std::string(argv[1]) == "--file") return run_file(argv[2]) ? 1 : 0;
  const int n = argc > 1 ? atoi(argv[1]) : 5000, nq = argc > 2 ? atoi(argv[2]) : 2000;
  const unsigned seed = argc > 3 ? (unsigned)atoi(argv[3]) : 1u; const int mode = argc > 4 ? atoi(argv[4]) : 0;
  return (mode == 0 ? run<true>(n, nq, seed) : run<false>(n, nq, seed)) ? 1 : 0;
}
