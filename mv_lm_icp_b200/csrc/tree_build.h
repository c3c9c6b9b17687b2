// tree_build.h -- host-side, one-time construction of the per-frame search tree (replaces the lazily built nanoflann
// index, src/internal/frame.cpp:188-193).  Included by mvicp.cu and by tools/sim_search.cpp (CPU step-count model).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>
#include "types.cuh"

using namespace mv;

static inline float f_down(double v) { float f = (float)v; if ((double)f > v) f = std::nextafterf(f, -INFINITY); return f; }
static inline float f_up(double v) { float f = (float)v; if ((double)f < v) f = std::nextafterf(f, INFINITY); return f; }

struct HostFrameBuild {
  std::vector<int32_t> order;      // tree order -> original index
  std::vector<int32_t> pos_of;     // original index -> tree position
  std::vector<Box> boxes;
  std::vector<float> faces;        // split-plane bound per node (axis in the low 2 mantissa bits)
  int n_leaf_pad = 1, depth = 0;
  float absmax = 0.f;
};

// Left-balanced KD ordering: the node that covers leaf slots [a, b) of the implicit tree holds the points at sorted
// positions [LEAF a, min(LEAF b, n)); each internal node splits its points at the capacity of its left half along
// the widest axis of their bounding box (nth_element), so sibling boxes never overlap and leaves are compact.
static void kd_order(const double* pts, int32_t* idx, int64_t begin, int64_t count, int64_t leaf_slots, int node, uint8_t* axis_of) {
  if (leaf_slots <= 1) return;
  axis_of[node] = 0;
  if (count <= LEAF) { kd_order(pts, idx, begin, count, leaf_slots / 2, 2 * node, axis_of); return; }
  const int64_t cap_left = (leaf_slots / 2) * LEAF;
  if (count > cap_left) {
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = begin; i < begin + count; ++i)
      for (int a = 0; a < 3; ++a) { const double v = pts[3 * (int64_t)idx[i] + a]; lo[a] = std::min(lo[a], v); hi[a] = std::max(hi[a], v); }
    int ax = 0; for (int a = 1; a < 3; ++a) if (hi[a] - lo[a] > hi[ax] - lo[ax]) ax = a;
    axis_of[node] = (uint8_t)ax;
    std::nth_element(idx + begin, idx + begin + cap_left, idx + begin + count, [&](int32_t x, int32_t y) {
      const double vx = pts[3 * (int64_t)x + ax], vy = pts[3 * (int64_t)y + ax];
      return vx < vy || (vx == vy && x < y);
    });
    kd_order(pts, idx, begin, cap_left, leaf_slots / 2, 2 * node, axis_of);
    kd_order(pts, idx, begin + cap_left, count - cap_left, leaf_slots / 2, 2 * node + 1, axis_of);
  } else {
    kd_order(pts, idx, begin, count, leaf_slots / 2, 2 * node, axis_of);
  }
}

static void build_frame(const double* pts, int64_t n, HostFrameBuild& out) {
  double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int64_t i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], pts[3 * i + a]); hi[a] = std::max(hi[a], pts[3 * i + a]); }
  const int64_t n_leaf = std::max<int64_t>(1, (n + LEAF - 1) / LEAF);
  int L = 1; while (L < n_leaf) L <<= 1;
  out.order.resize(n); out.pos_of.resize(n);
  std::iota(out.order.begin(), out.order.end(), 0);
  std::vector<uint8_t> axis_of((size_t)2 * L, 0);
  kd_order(pts, out.order.data(), 0, n, L, 1, axis_of.data());
  for (int64_t i = 0; i < n; ++i) out.pos_of[out.order[i]] = (int32_t)i;
  double am = 0; for (int a = 0; a < 3; ++a) am = std::max(am, std::max(std::fabs(lo[a]), std::fabs(hi[a])));
  out.absmax = f_up(am);
  out.n_leaf_pad = L;
  out.depth = 0; while ((1 << out.depth) < L) ++out.depth;
  // axis-aligned boxes, rounded outward, bottom-up; and per node a one-sided bound along its parent's split axis:
  // every point of a right (odd) node has coordinate >= face, of a left (even) node <= face (rounded conservatively,
  // the axis rides in the two low mantissa bits) -- the cheap split-plane test of the upward sweep.
  out.boxes.assign((size_t)2 * L, Box{{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}, {0, 0}});
  std::vector<double> dlo((size_t)2 * L * 3, INFINITY), dhi((size_t)2 * L * 3, -INFINITY);
  for (int64_t l = 0; l < n_leaf; ++l)
    for (int64_t i = l * LEAF; i < std::min<int64_t>(n, (l + 1) * LEAF); ++i)
      for (int a = 0; a < 3; ++a) {
        const double v = pts[3 * (int64_t)out.order[i] + a];
        dlo[(L + l) * 3 + a] = std::min(dlo[(L + l) * 3 + a], v); dhi[(L + l) * 3 + a] = std::max(dhi[(L + l) * 3 + a], v);
      }
  for (int i = L - 1; i >= 1; --i)
    for (int a = 0; a < 3; ++a) {
      dlo[(size_t)i * 3 + a] = std::min(dlo[(size_t)2 * i * 3 + a], dlo[(size_t)(2 * i + 1) * 3 + a]);
      dhi[(size_t)i * 3 + a] = std::max(dhi[(size_t)2 * i * 3 + a], dhi[(size_t)(2 * i + 1) * 3 + a]);
    }
  out.faces.assign((size_t)2 * L, 0.f);
  for (int i = 1; i < 2 * L; ++i) {
    for (int a = 0; a < 3; ++a) { out.boxes[i].lo[a] = f_down(dlo[(size_t)i * 3 + a]); out.boxes[i].hi[a] = f_up(dhi[(size_t)i * 3 + a]); }
    if (i >= 2) {
      const int axp = axis_of[i / 2]; const bool right = (i & 1) != 0;
      float ff = right ? f_down(dlo[(size_t)i * 3 + axp]) : f_up(dhi[(size_t)i * 3 + axp]);   // +inf / -inf for empty nodes
      for (int guard = 0; guard < 8 && std::isfinite(ff); ++guard) {
        uint32_t bits; std::memcpy(&bits, &ff, 4);
        if ((bits & 3u) == (uint32_t)axp) break;
        ff = std::nextafterf(ff, right ? -INFINITY : INFINITY);
      }
      if (!std::isfinite(ff)) { uint32_t bits; std::memcpy(&bits, &ff, 4); bits = (bits & ~3u); std::memcpy(&ff, &bits, 4); }   // inf: low bits 0 = axis 0, any axis prunes
      out.faces[i] = ff;
    }
  }
}

