// walk.cuh -- EXPERIMENTAL (MVICP_FLAG_GRAPH_WALK): exact 1-NN for seeded rounds by a walk on the dst cloud's neighbour
// graph with a distance certificate, the tree search of knn.cuh as its fallback.  Same results as knn_kernel, bit for bit.
//
// Every dst point m keeps the tree positions of its WALK_K nearest other points and r2(m), the squared distance to the first
// point OUTSIDE that list.  For a query q seeded with the previous round's match:
//   walk    while some listed neighbour of the current point is closer to q (or as close, with a lower index), move there;
//   certify the point m where the walk stops is the best of itself and its list, and every other point p of the cloud has
//           |m - p| >= sqrt(r2(m)), hence |q - p| >= sqrt(r2(m)) - |q - m|: if 4 |q - m|^2 < r2(m) (with a 1e-9 margin for
//           the rounding of the computed squared distances) every such p is strictly farther than m, so m is the exact
//           arg-min under the engine's order (fp64 distance in the reference's operation order, then lowest index).
// With converged poses the certificate holds for ~93 % of the queries of the Bunny workload after 0-1 moves
// (tools/walk_rate.py); the others -- far matches, open boundaries, exhausted moves -- are compacted within the CTA and run
// the tree search, densely packed into its first warps, starting at the leaf where the walk stopped.  The kernel also produces
// the first histogram of the median select (bits 63:53 of the inlier distances), which saves that pass over the results.
#pragma once
#include "knn.cuh"
#include "normals.cuh"
#include "select.cuh"

namespace mv {

constexpr int WALK_K = 8;       // listed neighbours per point (two int4 loads)
constexpr int WALK_HOPS = 4;    // moves before giving up
struct WalkDev { const int32_t* nbr; const double* r2; };   // per frame: [n][WALK_K] tree positions, [n] by tree position

// Median bracket (converged rounds): the poses of a round that follows a one-iteration LM solve equal the previous round's
// up to conversion rounding, so the median of an edge's inlier distances sits where it was.  The NN kernel then counts, per
// edge, the inliers below a +-1/64-binade bracket around the previous median's bit pattern and collects the keys inside
// it; one small kernel per edge finishes the exact order statistic among them (select_guess_finish_kernel) -- no pass over
// the results at all.  If the wanted rank falls outside the bracket (or the bracket overflows), that kernel scans the edge
// itself: slow, exact, and not expected when the host's condition (previous LM solve: one iteration) holds.
constexpr int GUESS_CAP = 8192;
constexpr unsigned long long GUESS_HALF = 1ull << 46;        // 1/64 of a binade in key space (52 mantissa bits)
struct SelGuess { unsigned long long lo; unsigned long long below; unsigned long long n_in; unsigned int n_cand; int armed; };

// nn: output of normals_kernel with k = WALK_K + 2 (indices by (distance, index), the point itself among them), per
// original index.  A point with fewer than WALK_K + 1 other points gets r2 = 0: its certificate never holds.
template <bool F32>
__global__ void walk_build_kernel(const FrameDev* __restrict__ frames, int frame, const int32_t* __restrict__ nn,
                                  int32_t* __restrict__ nbr, double* __restrict__ r2) {
  const FrameDev fd = frames[frame];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= fd.n) return;
  double px, py, pz; int orig;
  Rec<F32>::load(fd.pts_s, t, px, py, pz, orig);
  int others[WALK_K + 1]; int m = 0;
  for (int j = 0; j < WALK_K + 2 && m < WALK_K + 1; ++j) {
    const int v = nn[(size_t)(WALK_K + 2) * orig + j];
    if (v >= 0 && v != orig) others[m++] = v;
  }
  for (int j = 0; j < WALK_K; ++j) nbr[(size_t)WALK_K * t + j] = (m == WALK_K + 1) ? __ldg(fd.pos_of + others[j]) : t;
  double rr = 0.0;
  if (m == WALK_K + 1) {
    double ox, oy, oz; int dummy; Rec<F32>::load(fd.pts_o, others[WALK_K], ox, oy, oz, dummy);
    rr = d2_rn(px, py, pz, ox, oy, oz);
  }
  r2[t] = rr;
}

template <bool F32>
__global__ void __launch_bounds__(KNN_TILE)
knn_walk_kernel(const FrameDev* __restrict__ frames, const EdgeDev* __restrict__ edges, const EdgeXf* __restrict__ xfs,
                const Tile* __restrict__ tiles, int32_t* __restrict__ corr, double* __restrict__ d2out,
                const int32_t* __restrict__ seed, double thresh, const WalkDev* __restrict__ walk,
                unsigned int* __restrict__ hist /* [E][SEL_BINS]: first pass of the median select, fused (select.cuh); null with guess */,
                SelGuess* __restrict__ guess /* nullable: [E] median brackets */, unsigned long long* __restrict__ gcand /* [E][GUESS_CAP] */) {
  const Tile t = tiles[blockIdx.x];
  const EdgeDev e = edges[t.edge];
  __shared__ EdgeXf sx;
  __shared__ int s_fail[KNN_TILE], s_start[KNN_TILE];
  __shared__ int s_nfail;
  __shared__ unsigned int s_hist[SEL_BINS];
  __shared__ unsigned int s_in, s_below;
  if (hist) for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) s_hist[i] = 0u;
  const unsigned long long g_lo = guess ? guess[t.edge].lo : 0ull;
  {
    if (threadIdx.x == 0) { s_in = 0u; s_below = 0u; }
    const double* g = reinterpret_cast<const double*>(xfs + t.edge);
    double* s = reinterpret_cast<double*>(&sx);
    for (int i = threadIdx.x; i < (int)(sizeof(EdgeXf) / sizeof(double)); i += blockDim.x) s[i] = g[i];
    if (threadIdx.x == 0) s_nfail = 0;
  }
  __syncthreads();
  const FrameDev fs = frames[e.src];
  const FrameDev fd = frames[e.dst];
  const WalkDev wd = walk[e.dst];
  // phase A: walk + certificate, one query per thread; phase B (pass 1): tree search of the queries that failed
  int n_work = KNN_TILE;
  for (int pass = 0; pass < 2; ++pass) {
    for (int w = threadIdx.x; w < n_work; w += KNN_TILE) {
      const int slot = pass == 0 ? w : s_fail[w];
      const int ks = t.start + slot;
      if (ks >= e.n_src) continue;
      double px, py, pz; int orig;
      Rec<F32>::load(fs.pts_s, ks, px, py, pz, orig);
      // query transform, the operation sequence of knn_kernel (frame.cpp:117-118,131,136)
      const double gx = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(sx.Rs[0], px), __dmul_rn(sx.Rs[1], py)), __dmul_rn(sx.Rs[2], pz)), sx.ts[0]);
      const double gy = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(sx.Rs[3], px), __dmul_rn(sx.Rs[4], py)), __dmul_rn(sx.Rs[5], pz)), sx.ts[1]);
      const double gz = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(sx.Rs[6], px), __dmul_rn(sx.Rs[7], py)), __dmul_rn(sx.Rs[8], pz)), sx.ts[2]);
      const double ex = __dsub_rn(gx, sx.td[0]), ey = __dsub_rn(gy, sx.td[1]), ez = __dsub_rn(gz, sx.td[2]);
      const double qx = __dadd_rn(__dadd_rn(__dmul_rn(sx.Rinv[0], ex), __dmul_rn(sx.Rinv[1], ey)), __dmul_rn(sx.Rinv[2], ez));
      const double qy = __dadd_rn(__dadd_rn(__dmul_rn(sx.Rinv[3], ex), __dmul_rn(sx.Rinv[4], ey)), __dmul_rn(sx.Rinv[5], ez));
      const double qz = __dadd_rn(__dadd_rn(__dmul_rn(sx.Rinv[6], ex), __dmul_rn(sx.Rinv[7], ey)), __dmul_rn(sx.Rinv[8], ez));
      NNQuery nq; nn_query_init(nq, qx, qy, qz, fd.absmax);
      bool finished = false;
      if (pass == 0) {
        int start_leaf = -1;
        const int sd = seed[e.off + orig];
        const int si = sd >= 0 ? sd : ~sd;
        if (si >= 0 && si < fd.n) {
          int cur = __ldg(fd.pos_of + si);
          { const float4 r = __ldg(fd.pts_sf + cur); nn_exact<F32>(fd, cur, r, nq); }
          for (int hop = 0; hop <= WALK_HOPS; ++hop) {
            const int4 la = __ldg(reinterpret_cast<const int4*>(wd.nbr + (size_t)WALK_K * cur));
            const int4 lb = __ldg(reinterpret_cast<const int4*>(wd.nbr + (size_t)WALK_K * cur) + 1);
            const int nb[WALK_K] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
            int next = cur;
#pragma unroll
            for (int j = 0; j < WALK_K; ++j) {
              const float4 r = __ldg(fd.pts_sf + nb[j]);
              if (pt_d32(r, nq) <= nq.bound32) {
                const int before = nq.bi;
                nn_exact<F32>(fd, nb[j], r, nq);
                if (nq.bi != before) next = nb[j];
              }
            }
            if (next != cur) { cur = next; continue; }
            finished = __dmul_rn(__dmul_rn(4.0, nq.best), 1.000000001) < __ldg(wd.r2 + cur);
            break;
          }
          start_leaf = cur / LEAF;
        }
        if (!finished) { const int k = atomicAdd(&s_nfail, 1); s_fail[k] = slot; s_start[k] = start_leaf; continue; }
      } else {
        nn_search<F32, NNQuery>(fd, nq, s_start[w]);
      }
      const double best = nq.best; const int bi = nq.bi;
      const bool inlier = __dsqrt_rn(best) < thresh;
      corr[e.off + orig] = inlier ? bi : ~bi;
      d2out[e.off + orig] = best;
      if (inlier) {
        const unsigned long long key = (unsigned long long)__double_as_longlong(best);
        if (hist) atomicAdd(&s_hist[(unsigned int)(key >> 53)], 1u);   // select_hist_kernel, shift 53
        if (guess) {
          atomicAdd(&s_in, 1u);
          if (key < g_lo) atomicAdd(&s_below, 1u);
          else if (key - g_lo < 2 * GUESS_HALF) {
            const unsigned int k = atomicAdd(&guess[t.edge].n_cand, 1u);
            if (k < (unsigned int)GUESS_CAP) gcand[(size_t)t.edge * GUESS_CAP + k] = key;
          }
        }
      }
    }
    __syncthreads();
    n_work = s_nfail;
  }
  if (hist) {
    unsigned int* h = hist + (size_t)t.edge * SEL_BINS;
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) if (s_hist[i]) atomicAdd(&h[i], s_hist[i]);
  }
  if (guess && threadIdx.x == 0) {
    if (s_in) atomicAdd(&guess[t.edge].n_in, (unsigned long long)s_in);
    if (s_below) atomicAdd(&guess[t.edge].below, (unsigned long long)s_below);
  }
}

// After a full select: bracket around the exact median bit pattern it left in SelState::prefix; counters cleared.
__global__ void select_guess_arm_kernel(const SelState* __restrict__ st, SelGuess* __restrict__ guess, int n_edges) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  SelGuess g; g.below = 0ull; g.n_in = 0ull; g.n_cand = 0u;
  g.armed = st[e].count != 0 && st[e].prefix >= GUESS_HALF ? 1 : 0;
  g.lo = g.armed ? st[e].prefix - GUESS_HALF : 0ull;
  guess[e] = g;
}

// One CTA per edge: the exact order statistic (position count / 2 of the sorted inlier distances, frame.cpp:166-168) from the
// bracket's counts and keys; if the rank is not inside the bracket, a radix select over the edge itself.  Leaves SelState,
// weight, median and count as the full select does, and re-arms the bracket around the new median.
__global__ void __launch_bounds__(SEL_THREADS)
select_guess_finish_kernel(const EdgeDev* __restrict__ edges, const int32_t* __restrict__ corr, const double* __restrict__ d2,
                           SelState* __restrict__ st, SelGuess* __restrict__ guess, const unsigned long long* __restrict__ gcand,
                           float* __restrict__ weight, double* __restrict__ median, unsigned long long* __restrict__ edge_count) {
  const int e = blockIdx.x;
  __shared__ unsigned int sh[SEL_BINS];
  __shared__ unsigned int part[SEL_THREADS];
  __shared__ int s_bin;
  __shared__ unsigned long long s_prefix, s_rank, s_before;
  const SelGuess g = guess[e];
  const EdgeDev ed = edges[e];
  const unsigned long long n_in = g.n_in;
  const bool inside = g.armed && g.n_cand <= (unsigned int)GUESS_CAP && n_in / 2 >= g.below && n_in / 2 < g.below + g.n_cand;
  if (threadIdx.x == 0) { s_prefix = 0ull; s_rank = inside ? n_in / 2 - g.below : n_in / 2; }
  __syncthreads();
  const int shifts[6] = {53, 42, 31, 20, 9, 0}, nbits[6] = {11, 11, 11, 11, 11, 9};
  if (n_in != 0) {
    for (int p = 0; p < 6; ++p) {
      for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) sh[i] = 0u;
      __syncthreads();
      const int shift = shifts[p], hi = shift + nbits[p];
      const unsigned int mask = (1u << nbits[p]) - 1u;
      const unsigned long long prefix = s_prefix;
      if (inside) {
        for (unsigned int i = threadIdx.x; i < g.n_cand; i += blockDim.x) {
          const unsigned long long key = gcand[(size_t)e * GUESS_CAP + i];
          if (hi >= 64 || (key >> hi) == (prefix >> hi)) atomicAdd(&sh[(unsigned int)(key >> shift) & mask], 1u);
        }
      } else {   // not expected after a one-iteration LM solve: the bracket missed -- scan the edge itself
        for (int k = threadIdx.x; k < ed.n_src; k += blockDim.x) {
          if (corr[ed.off + k] < 0) continue;
          const unsigned long long key = (unsigned long long)__double_as_longlong(d2[ed.off + k]);
          if (hi >= 64 || (key >> hi) == (prefix >> hi)) atomicAdd(&sh[(unsigned int)(key >> shift) & mask], 1u);
        }
      }
      __syncthreads();
      {
        constexpr int PER = SEL_BINS / SEL_THREADS;
        unsigned int sum = 0;
        for (int i = 0; i < PER; ++i) sum += sh[threadIdx.x * PER + i];
        part[threadIdx.x] = sum;
        if (threadIdx.x == 0) s_bin = -1;
        __syncthreads();
        if (threadIdx.x == 0) {
          unsigned long long acc = 0; const unsigned long long rank = s_rank;
          for (int q = 0; q < SEL_THREADS; ++q) { if (rank < acc + part[q]) { s_bin = q; s_before = acc; break; } acc += part[q]; }
        }
        __syncthreads();
        if (s_bin == (int)threadIdx.x) {
          unsigned long long acc = s_before; const unsigned long long rank = s_rank;
          for (int i = 0; i < PER; ++i) {
            const unsigned int hcount = sh[threadIdx.x * PER + i];
            if (rank < acc + hcount) { s_prefix = prefix | ((unsigned long long)(threadIdx.x * PER + i) << shift); s_rank = rank - acc; break; }
            acc += hcount;
          }
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    st[e].prefix = s_prefix; st[e].rank = s_rank; st[e].count = n_in;
    edge_count[e] = n_in;
    if (n_in == 0) { weight[e] = 0.0f; median[e] = __longlong_as_double(0x7ff8000000000000LL); }
    else {
      const double nth = __dsqrt_rn(__longlong_as_double((long long)s_prefix));
      median[e] = nth;
      weight[e] = __double2float_rn(__dmul_rn(nth, 1.5));
    }
    SelGuess ng; ng.below = 0ull; ng.n_in = 0ull; ng.n_cand = 0u;
    ng.armed = n_in != 0 && s_prefix >= GUESS_HALF ? 1 : 0;
    ng.lo = ng.armed ? s_prefix - GUESS_HALF : 0ull;
    guess[e] = ng;
  }
}

}  // namespace mv
