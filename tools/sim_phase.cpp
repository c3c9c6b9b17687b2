// tools/sim_phase.cpp -- lock-step WARP model of a phase-separated search: all lanes expand internal nodes (phase A, bound
// frozen) until every node stack is empty or some lane's leaf list is full, then all lanes scan their listed leaves (phase B).
// Counts warp slots of each kind.  Development aid only.
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include "../mv_lm_icp_b200/csrc/tree_build.h"

struct Sim { HostFrameBuild b; std::vector<float> px, py, pz; std::vector<int> pi; int64_t n; };
static inline float lb32(const Box& b, float fx, float fy, float fz) {
  const float dx = std::fmax(std::fmax(b.lo[0] - fx, fx - b.hi[0]), 0.f), dy = std::fmax(std::fmax(b.lo[1] - fy, fy - b.hi[1]), 0.f),
              dz = std::fmax(std::fmax(b.lo[2] - fz, fz - b.hi[2]), 0.f);
  return dx * dx + dy * dy + dz * dz;
}
struct Lane {
  float fx, fy, fz; double q[3]; float bound; double best; int bi; bool active;
  std::vector<int> ns; std::vector<float> nl;    // node stack
  std::vector<int> ls; std::vector<float> ll;    // leaf list
};
extern "C" {
void* ph_build(const double* pts, int64_t n) {
  Sim* s = new Sim(); s->n = n; build_frame(pts, n, s->b);
  const int64_t npad = ((n + LEAF - 1) / LEAF) * LEAF;
  s->px.assign(npad, INFINITY); s->py.assign(npad, INFINITY); s->pz.assign(npad, INFINITY); s->pi.assign(npad, INT32_MAX);
  for (int64_t i = 0; i < n; ++i) { const int o = s->b.order[i]; s->px[i] = (float)pts[3 * o]; s->py[i] = (float)pts[3 * o + 1]; s->pz[i] = (float)pts[3 * o + 2]; s->pi[i] = o; }
  return s;
}
void ph_order(void* h, int* out) { Sim& s = *(Sim*)h; for (int64_t i = 0; i < s.n; ++i) out[i] = s.b.order[i]; }
int ph_leaf_of(void* h, int orig) { return ((Sim*)h)->b.pos_of[orig] / LEAF; }
// q: 32 x 3 doubles; start_leaf[32]; out_idx[32]; counts: [0] A slots, [1] B slots (one leaf of 8 points each), [2] descent slots,
// [3] lane box steps, [4] lane leaf scans, [5] phase switches
void ph_warp(void* h, const double* q, const int* start_leaf, int cap, int* out_idx, int64_t* counts) {
  Sim& s = *(Sim*)h; const HostFrameBuild& t = s.b; const int L = t.n_leaf_pad;
  Lane ln[32];
  auto scan_leaf = [&](Lane& a, int leaf) {
    for (int j = 0; j < LEAF; ++j) {
      const int64_t pos = (int64_t)leaf * LEAF + j;
      const float dx = a.fx - s.px[pos], dy = a.fy - s.py[pos], dz = a.fz - s.pz[pos];
      const float d32 = dx * dx + dy * dy + dz * dz;
      if (d32 <= a.bound) {
        const double ex = a.q[0] - (double)s.px[pos], ey = a.q[1] - (double)s.py[pos], ez = a.q[2] - (double)s.pz[pos];
        const double d = ex * ex + ey * ey + ez * ez;
        if (d < a.best || (d == a.best && s.pi[pos] < a.bi)) { a.best = d; a.bi = s.pi[pos]; const double r = std::sqrt(a.best) + 1e-6; a.bound = (float)(r * r * 1.000001); }
      }
    }
  };
  bool any_stale = false; int leaf_node[32];
  for (int i = 0; i < 32; ++i) {
    Lane& a = ln[i]; a.q[0] = q[3 * i]; a.q[1] = q[3 * i + 1]; a.q[2] = q[3 * i + 2]; a.fx = (float)a.q[0]; a.fy = (float)a.q[1]; a.fz = (float)a.q[2];
    a.bound = INFINITY; a.best = INFINITY; a.bi = INT32_MAX;
    int sl = start_leaf[i]; leaf_node[i] = -1;
    if (sl >= 0) {
      leaf_node[i] = L + sl; scan_leaf(a, sl);
      const Box& b = t.boxes[L + sl]; const float ex = b.hi[0] - b.lo[0], ey = b.hi[1] - b.lo[1], ez = b.hi[2] - b.lo[2];
      if (a.bound > 16.f * (ex * ex + ey * ey + ez * ez)) sl = -1;
    }
    if (sl < 0) {
      any_stale = true; int node = 1;
      while (node < L) { const float l0 = lb32(t.boxes[2 * node], a.fx, a.fy, a.fz), l1 = lb32(t.boxes[2 * node + 1], a.fx, a.fy, a.fz); node = (l1 < l0) ? 2 * node + 1 : 2 * node; }
      if (node != leaf_node[i]) scan_leaf(a, node - L);
      leaf_node[i] = node;
    }
    // sweep: siblings, pushed top-down
    for (int l = t.depth - 1; l >= 0; --l) {
      const int sib = (leaf_node[i] >> l) ^ 1;
      const float face = t.faces[sib]; uint32_t bits; std::memcpy(&bits, &face, 4); const int axis = bits & 3;
      const float qa = axis == 0 ? a.fx : (axis == 1 ? a.fy : a.fz);
      const float dpl = (sib & 1) ? face - qa : qa - face;
      if (dpl > 0.f && dpl * dpl > a.bound) continue;
      const float lb = lb32(t.boxes[sib], a.fx, a.fy, a.fz);
      if (lb <= a.bound) { if (sib >= L) { a.ls.push_back(sib); a.ll.push_back(lb); } else { a.ns.push_back(sib); a.nl.push_back(lb); } }
    }
  }
  if (any_stale) counts[2] += t.depth + 1;
  while (true) {
    // phase A
    bool anyA = false;
    while (true) {
      bool work = false, full = false;
      for (auto& a : ln) { if ((int)a.ls.size() >= cap) full = true; }
      if (full) break;
      for (auto& a : ln) {
        // pop until a live node
        int node = -1;
        while (!a.ns.empty()) { const int n2 = a.ns.back(); const float l2 = a.nl.back(); a.ns.pop_back(); a.nl.pop_back(); if (l2 <= a.bound) { node = n2; break; } }
        if (node < 0) continue;
        work = true; ++counts[3];
        const int c0 = 2 * node; const float l0 = lb32(t.boxes[c0], a.fx, a.fy, a.fz), l1 = lb32(t.boxes[c0 + 1], a.fx, a.fy, a.fz);
        const bool f0 = l0 <= l1; const int cn = f0 ? c0 : c0 + 1, cf = f0 ? c0 + 1 : c0; const float lnr = f0 ? l0 : l1, lfr = f0 ? l1 : l0;
        if (c0 >= L) {   // children are leaves: far first so that the near one is scanned first (list popped from the back)
          if (lfr <= a.bound) { a.ls.push_back(cf); a.ll.push_back(lfr); }
          if (lnr <= a.bound) { a.ls.push_back(cn); a.ll.push_back(lnr); }
        } else {
          if (lfr <= a.bound) { a.ns.push_back(cf); a.nl.push_back(lfr); }
          if (lnr <= a.bound) { a.ns.push_back(cn); a.nl.push_back(lnr); }
        }
      }
      if (!work) break;
      ++counts[0]; anyA = true;
    }
    // phase B
    bool anyB = false;
    while (true) {
      bool work = false;
      for (auto& a : ln) {
        int leaf = -1;
        while (!a.ls.empty()) { const int n2 = a.ls.back(); const float l2 = a.ll.back(); a.ls.pop_back(); a.ll.pop_back(); if (l2 <= a.bound) { leaf = n2; break; } }
        if (leaf < 0) continue;
        work = true; ++counts[4]; scan_leaf(a, leaf - L);
      }
      if (!work) break;
      ++counts[1]; anyB = true;
    }
    ++counts[5];
    if (!anyA && !anyB) break;
  }
  for (int i = 0; i < 32; ++i) out_idx[i] = ln[i].bi;
}
}
