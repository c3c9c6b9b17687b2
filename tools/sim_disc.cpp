// tools/sim_disc.cpp -- step-count model of the NN search with a different node bound: a thin DISC per node (centroid c,
// PCA normal n, half thickness along n, radius perpendicular to n: 8 floats, the size of the AABB record).  A depth scan is
// a tilted, locally flat sheet: its axis-aligned boxes are as thick as the tilt makes them, and a query that is still
// millimetres off the surface must open every box within sqrt(height * thickness) of its foot point; the disc's thickness is
// the sheet's own.  Modes: 0 = AABB (the engine today), 1 = disc, 2 = max(AABB, disc).  Same search order as csrc/knn.cuh
// (seed leaf / greedy descent, bottom-up sibling sweep with plane pre-filter, nearest-first DFS); exact arithmetic in
// double, so only the COUNTS are modelled.  Development aid.
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../mv_lm_icp_b200/csrc/tree_build.h"

struct Disc { double c[3], n[3], th, r; };
struct Sim { HostFrameBuild b; std::vector<double> p; std::vector<int> pi; std::vector<Disc> disc; int64_t n; };
static int g_mode = 0;

static void eig3(double A[3][3], double w[3], double V[3][3]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = (i == j);
  for (int sweep = 0; sweep < 60; ++sweep) {
    if (A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2] < 1e-300) break;
    for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
      if (A[p][q] == 0.0) continue;
      const double th = (A[q][q] - A[p][p]) / (2.0 * A[p][q]), t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0)), c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < 3; ++k) { const double a = A[k][p], b = A[k][q]; A[k][p] = c * a - s * b; A[k][q] = s * a + c * b; }
      for (int k = 0; k < 3; ++k) { const double a = A[p][k], b = A[q][k]; A[p][k] = c * a - s * b; A[q][k] = s * a + c * b; }
      for (int k = 0; k < 3; ++k) { const double a = V[k][p], b = V[k][q]; V[k][p] = c * a - s * b; V[k][q] = s * a + c * b; }
    }
  }
  for (int i = 0; i < 3; ++i) w[i] = A[i][i];
}

static double lb_aabb(const Box& b, const double* q) {
  double s = 0; for (int a = 0; a < 3; ++a) { const double d = std::fmax(std::fmax((double)b.lo[a] - q[a], q[a] - (double)b.hi[a]), 0.0); s += d * d; } return s;
}
static double lb_disc(const Disc& d, const double* q) {
  const double v[3] = {q[0] - d.c[0], q[1] - d.c[1], q[2] - d.c[2]};
  const double dn = v[0] * d.n[0] + v[1] * d.n[1] + v[2] * d.n[2];
  const double w[3] = {v[0] - dn * d.n[0], v[1] - dn * d.n[1], v[2] - dn * d.n[2]};
  const double gl = std::fmax(std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]) - d.r, 0.0), gn = std::fmax(std::fabs(dn) - d.th, 0.0);
  return gn * gn + gl * gl;
}
static double lb(const Sim& s, int node, const double* q) {
  if (g_mode == 0) return lb_aabb(s.b.boxes[node], q);
  if (g_mode == 1) return lb_disc(s.disc[node], q);
  return std::fmax(lb_aabb(s.b.boxes[node], q), lb_disc(s.disc[node], q));
}

extern "C" {
void sd_mode(int m) { g_mode = m; }
void* sd_build(const double* pts, int64_t n) {
  Sim* s = new Sim(); s->n = n; build_frame(pts, n, s->b);
  const int L = s->b.n_leaf_pad; const int64_t npad = (int64_t)L * LEAF;
  s->p.assign(3 * npad, INFINITY); s->pi.assign(npad, INT32_MAX);
  for (int64_t i = 0; i < n; ++i) { const int o = s->b.order[i]; for (int a = 0; a < 3; ++a) s->p[3 * i + a] = pts[3 * o + a]; s->pi[i] = o; }
  s->disc.assign((size_t)2 * L, Disc{{0, 0, 0}, {0, 0, 1}, INFINITY, INFINITY});
  for (int node = 1; node < 2 * L; ++node) {
    int lvl = 0; while ((1 << (lvl + 1)) <= node) ++lvl;
    const int64_t span = (int64_t)(L >> lvl) * LEAF, first = (int64_t)(node - (1 << lvl)) * span, last = std::min<int64_t>(n, first + span);
    if (first >= n) continue;
    Disc d; double c[3] = {0, 0, 0};
    for (int64_t i = first; i < last; ++i) for (int a = 0; a < 3; ++a) c[a] += s->p[3 * i + a];
    for (int a = 0; a < 3; ++a) c[a] /= (double)(last - first);
    double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int64_t i = first; i < last; ++i) { double v[3]; for (int a = 0; a < 3; ++a) v[a] = s->p[3 * i + a] - c[a]; for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) C[a][b] += v[a] * v[b]; }
    double w[3], V[3][3]; eig3(C, w, V); int m = 0; if (w[1] < w[m]) m = 1; if (w[2] < w[m]) m = 2;
    for (int a = 0; a < 3; ++a) { d.c[a] = c[a]; d.n[a] = V[a][m]; }
    // centre the slab: shift c along n to the middle of the extent
    double lo = INFINITY, hi = -INFINITY;
    for (int64_t i = first; i < last; ++i) { double dn = 0; for (int a = 0; a < 3; ++a) dn += (s->p[3 * i + a] - c[a]) * d.n[a]; lo = std::fmin(lo, dn); hi = std::fmax(hi, dn); }
    for (int a = 0; a < 3; ++a) d.c[a] += 0.5 * (lo + hi) * d.n[a];
    d.th = 0.5 * (hi - lo); d.r = 0;
    for (int64_t i = first; i < last; ++i) { double v[3], dn = 0; for (int a = 0; a < 3; ++a) { v[a] = s->p[3 * i + a] - d.c[a]; dn += v[a] * d.n[a]; } double rr = 0; for (int a = 0; a < 3; ++a) { const double u = v[a] - dn * d.n[a]; rr += u * u; } d.r = std::fmax(d.r, std::sqrt(rr)); }
    s->disc[node] = d;
  }
  return s;
}
int sd_leaf_of(void* h, int orig) { return ((Sim*)h)->b.pos_of[orig] / LEAF; }
// counts: [0] bound tests, [1] point tests, [2] plane tests
int sd_query(void* h, const double* q, int start_leaf, int64_t* counts) {
  Sim& s = *(Sim*)h; const HostFrameBuild& t = s.b; const int L = t.n_leaf_pad;
  double best = INFINITY; int bi = INT32_MAX;
  auto scan = [&](int leaf) { for (int j = 0; j < LEAF; ++j) { const int64_t pos = (int64_t)leaf * LEAF + j; if (pos >= s.n) break; ++counts[1];
      const double e0 = q[0] - s.p[3 * pos], e1 = q[1] - s.p[3 * pos + 1], e2 = q[2] - s.p[3 * pos + 2], d = e0 * e0 + e1 * e1 + e2 * e2;
      if (d < best || (d == best && s.pi[pos] < bi)) { best = d; bi = s.pi[pos]; } } };
  int leaf_node = -1;
  if (start_leaf >= 0) {
    leaf_node = L + start_leaf; scan(start_leaf);
    const Box& b = t.boxes[leaf_node]; const double ex = b.hi[0] - b.lo[0], ey = b.hi[1] - b.lo[1], ez = b.hi[2] - b.lo[2];
    if (best > 16.0 * (ex * ex + ey * ey + ez * ez)) start_leaf = -1;
  }
  if (start_leaf < 0) {
    int node = 1;
    while (node < L) { const double l0 = lb(s, 2 * node, q), l1 = lb(s, 2 * node + 1, q); counts[0] += 2; node = (l1 < l0) ? 2 * node + 1 : 2 * node; }
    if (node != leaf_node) scan(node - L);
    leaf_node = node;
  }
  std::vector<int> sn; std::vector<double> sl;
  for (int l = t.depth - 1; l >= 0; --l) {
    const int sib = (leaf_node >> l) ^ 1;
    const float face = t.faces[sib]; uint32_t bits; std::memcpy(&bits, &face, 4); const int axis = bits & 3;
    const double dpl = (sib & 1) ? (double)face - q[axis] : q[axis] - (double)face; ++counts[2];
    if (dpl > 0 && dpl * dpl > best) continue;
    const double v = lb(s, sib, q); ++counts[0]; if (v <= best) { sn.push_back(sib); sl.push_back(v); }
  }
  int node = -1;
  while (true) {
    if (node < 0) { if (sn.empty()) break; const int nn = sn.back(); const double ll = sl.back(); sn.pop_back(); sl.pop_back(); if (ll > best) continue; node = nn; }
    if (node >= L) { scan(node - L); node = -1; }
    else {
      const int c0 = 2 * node; const double l0 = lb(s, c0, q), l1 = lb(s, c0 + 1, q); counts[0] += 2;
      const bool f0 = l0 <= l1; const double ln = f0 ? l0 : l1, lf = f0 ? l1 : l0;
      if (ln <= best) { if (lf <= best) { sn.push_back(f0 ? c0 + 1 : c0); sl.push_back(lf); } node = f0 ? c0 : c0 + 1; } else node = -1;
    }
  }
  return bi;
}
}
