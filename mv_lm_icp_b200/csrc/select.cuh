// select.cuh -- exact per-edge median of the inlier distances -> OutgoingEdge::weight.
//
// Reference: std::nth_element(dists.begin(), dists.begin() + size/2, dists.end()); weight = nth * 1.5 narrowed
// to float (src/internal/frame.cpp:166-176).  dists = sqrt(d2) and sqrt is monotone, so the element at sorted
// position size/2 of the distances is the sqrt of the element at that position of the squared distances.
// Non-negative doubles order like their bit patterns => an MSB-first radix select over the 64-bit patterns
// gives the exact order statistic.  Two full histogram passes (bits 63:53, 52:42; the first also counts the inliers)
// narrow each edge to the values sharing a 22-bit prefix -- a few hundred out of 200 k -- which one more pass collects
// into a small buffer; a final one-block-per-edge kernel finishes the remaining 42 bits on that buffer in shared memory
// (or, if an edge's bucket overflows the buffer -- masses of near-equal distances -- by scanning the edge itself).
//
// Converged rounds (the previous LM solve took one iteration) skip all three passes: a WINDOW of keys around the previous round's
// median is a guess, the NN kernel's epilogue (knn.cuh, SEL) counts the inliers and those below the window and collects the keys
// inside it, and select_guess_finish_kernel checks that the wanted rank falls inside the window -- then the collected keys hold the
// answer -- or else redoes the edge from scratch.  Exact either way; the window's half-width adapts per edge to hold ~10^3 keys.
#pragma once
#include <cuda_runtime.h>
#include "types.cuh"

namespace mv {

constexpr int SEL_BINS = 2048;
constexpr int SEL_THREADS = 256;

// next round's window around this round's median; `nc` = keys the current window held (0: there was none), `miss` = it failed
__device__ __forceinline__ void select_set_window(unsigned long long* __restrict__ win, int E, int e, unsigned long long med,
                                                  unsigned long long count, bool had_window, bool miss, unsigned int nc) {
  // first half-width: about 1000 of the edge's `count` keys on either side, for a density of order count / median
  int sh = had_window ? (int)win[2 * (size_t)E + e] : 53 - (64 - __clzll((long long)(count / 1024ull + 1ull)));
  if (had_window) {
    if (miss) sh += 2;
    else if (nc < 256u) sh += 1;
    else if (nc > (unsigned int)SEL_CAP / 2) sh -= 1;
  }
  sh = sh < 20 ? 20 : (sh > 51 ? 51 : sh);
  const unsigned long long d = 1ull << sh;
  win[e] = med > d ? med - d : 0ull;
  win[(size_t)E + e] = med + d;            // keys of finite doubles stay far below 2^63: no overflow
  win[2 * (size_t)E + e] = (unsigned long long)sh;
}

__global__ void select_init_kernel(SelState* __restrict__ st, int n_edges) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  st[e].prefix = 0ull; st[e].count = 0ull; st[e].rank = 0ull;   // count / rank are set by the first pick (histogram total)
}

// hist[e][bin] += #inliers of the tile whose high bits equal the edge's prefix and whose digit is `bin`
__global__ void __launch_bounds__(SEL_THREADS)
select_hist_kernel(const EdgeDev* __restrict__ edges, const Tile* __restrict__ tiles, int tile_len,
                   const int32_t* __restrict__ corr, const double* __restrict__ d2, const SelState* __restrict__ st,
                   int shift, int bits, unsigned int* __restrict__ hist) {
  __shared__ unsigned int sh[SEL_BINS];
  for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const Tile t = tiles[blockIdx.x];
  const EdgeDev e = edges[t.edge];
  const unsigned long long prefix = st[t.edge].prefix;
  const int hi = shift + bits;
  const unsigned int mask = (1u << bits) - 1u;
  const int end = min(t.start + tile_len, e.n_src);
  for (int k = t.start + threadIdx.x; k < end; k += blockDim.x) {
    if (corr[e.off + k] < 0) continue;
    const unsigned long long key = (unsigned long long)__double_as_longlong(d2[e.off + k]);
    if (hi < 64 && (key >> hi) != (prefix >> hi)) continue;
    atomicAdd(&sh[(unsigned int)(key >> shift) & mask], 1u);
  }
  __syncthreads();
  unsigned int* h = hist + (size_t)t.edge * SEL_BINS;
  for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) if (sh[i]) atomicAdd(&h[i], sh[i]);
}

// one block per edge: find the bin holding the wanted rank, extend the prefix, clear the histogram.
// On the last pass the prefix is the exact bit pattern of the median squared distance -> weight.
__global__ void __launch_bounds__(SEL_THREADS)
select_pick_kernel(SelState* __restrict__ st, unsigned int* __restrict__ hist, int shift, int first, int last,
                   float* __restrict__ weight, double* __restrict__ median, unsigned long long* __restrict__ edge_count) {
  const int e = blockIdx.x;
  unsigned int* h = hist + (size_t)e * SEL_BINS;
  __shared__ unsigned int part[SEL_THREADS];
  __shared__ int s_bin; __shared__ unsigned long long s_before;
  constexpr int PER = SEL_BINS / SEL_THREADS;
  unsigned int loc[PER]; unsigned int sum = 0;
  for (int i = 0; i < PER; ++i) { loc[i] = h[threadIdx.x * PER + i]; sum += loc[i]; }
  part[threadIdx.x] = sum;
  if (threadIdx.x == 0) { s_bin = -1; s_before = 0; }
  __syncthreads();
  if (threadIdx.x == 0) {   // 256 partial sums: serial scan is fine (once per edge per pass)
    if (first) {   // the first histogram covers every inlier: its total is dists.size(), the wanted rank size()/2
      unsigned long long tot = 0; for (int t = 0; t < SEL_THREADS; ++t) tot += part[t];
      st[e].count = tot; st[e].rank = tot / 2; edge_count[e] = tot;
    }
    unsigned long long acc = 0; const unsigned long long rank = st[e].rank;
    for (int t = 0; t < SEL_THREADS; ++t) {
      if (rank < acc + part[t]) { s_bin = t; s_before = acc; break; }
      acc += part[t];
    }
  }
  __syncthreads();
  if (s_bin == (int)threadIdx.x) {
    unsigned long long acc = s_before; const unsigned long long rank = st[e].rank;
    for (int i = 0; i < PER; ++i) {
      if (rank < acc + loc[i]) {
        st[e].prefix |= ((unsigned long long)(threadIdx.x * PER + i)) << shift;
        st[e].rank = rank - acc;
        break;
      }
      acc += loc[i];
    }
  }
  for (int i = 0; i < PER; ++i) h[threadIdx.x * PER + i] = 0u;
  __syncthreads();
  if (last && threadIdx.x == 0) {
    if (st[e].count == 0) { weight[e] = 0.0f; median[e] = __longlong_as_double(0x7ff8000000000000LL); }
    else {
      const double nth = __dsqrt_rn(__longlong_as_double((long long)st[e].prefix));
      median[e] = nth;
      weight[e] = __double2float_rn(__dmul_rn(nth, 1.5));
    }
  }
}


// append the keys that match the 22-bit prefix to the edge's candidate buffer
__global__ void __launch_bounds__(SEL_THREADS)
select_collect_kernel(const EdgeDev* __restrict__ edges, const Tile* __restrict__ tiles, int tile_len,
                      const int32_t* __restrict__ corr, const double* __restrict__ d2, const SelState* __restrict__ st,
                      unsigned long long* __restrict__ cand, unsigned int* __restrict__ cand_n) {
  const Tile t = tiles[blockIdx.x];
  const EdgeDev e = edges[t.edge];
  const unsigned long long prefix = st[t.edge].prefix >> 42;
  const int end = min(t.start + tile_len, e.n_src);
  for (int k = t.start + threadIdx.x; k < end; k += blockDim.x) {
    if (corr[e.off + k] < 0) continue;
    const unsigned long long key = (unsigned long long)__double_as_longlong(d2[e.off + k]);
    if ((key >> 42) != prefix) continue;
    const unsigned int slot = atomicAdd(&cand_n[t.edge], 1u);
    if (slot < SEL_CAP) cand[(size_t)t.edge * SEL_CAP + slot] = key;
  }
}

// One block: digits p0..5 of {63:53, 52:42, 41:31, 30:20, 19:9, 8:0} by histogram + pick, over the nc candidate keys or
// (from_edge) over the edge's own inliers; s_prefix / s_rank (shared, set by the caller before a barrier) are refined in place.
__device__ __forceinline__ void select_digits(const EdgeDev& ed, const int32_t* __restrict__ corr, const double* __restrict__ d2,
                                              const unsigned long long* __restrict__ cand_e, unsigned int nc, bool from_edge, int p0,
                                              unsigned int* sh, unsigned long long* wtot /* 8 */,
                                              unsigned long long* s_prefix, unsigned long long* s_rank) {
  const int shifts[6] = {53, 42, 31, 20, 9, 0}, nbits[6] = {11, 11, 11, 11, 11, 9};
  for (int p = p0; p < 6; ++p) {
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const int shift = shifts[p], hi = shift + nbits[p];
    const unsigned int mask = (1u << nbits[p]) - 1u;
    const unsigned long long prefix = *s_prefix;
    if (!from_edge) {
      for (unsigned int i = threadIdx.x; i < nc; i += blockDim.x) {
        const unsigned long long key = cand_e[i];
        if (hi >= 64 || (key >> hi) == (prefix >> hi)) atomicAdd(&sh[(unsigned int)(key >> shift) & mask], 1u);
      }
    } else {   // rare: more equal-prefix values than the buffer holds, or a guess that missed -- scan the edge itself
      for (int k = threadIdx.x; k < ed.n_src; k += blockDim.x) {
        if (corr[ed.off + k] < 0) continue;
        const unsigned long long key = (unsigned long long)__double_as_longlong(d2[ed.off + k]);
        if (hi >= 64 || (key >> hi) == (prefix >> hi)) atomicAdd(&sh[(unsigned int)(key >> shift) & mask], 1u);
      }
    }
    __syncthreads();
    {   // 8 bins per thread; prefix sums over the 256 threads by warp shuffles + 8 warp totals; the thread whose range holds the rank picks
      constexpr int PER = SEL_BINS / SEL_THREADS;
      unsigned long long sum = 0;
      for (int i = 0; i < PER; ++i) sum += sh[threadIdx.x * PER + i];
      const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
      unsigned long long incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
      if (lane == 31) wtot[wid] = incl;
      __syncthreads();
      unsigned long long base = 0;
      for (int w = 0; w < wid; ++w) base += wtot[w];
      incl += base;
      const unsigned long long rank = *s_rank;
      __syncthreads();                       // everyone has read s_rank and wtot before the owner rewrites s_rank
      if (incl - sum <= rank && rank < incl) {
        unsigned long long acc = incl - sum;
        for (int i = 0; i < PER; ++i) {
          const unsigned int hcount = sh[threadIdx.x * PER + i];
          if (rank < acc + hcount) { *s_prefix = prefix | ((unsigned long long)(threadIdx.x * PER + i) << shift); *s_rank = rank - acc; break; }
          acc += hcount;
        }
      }
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void select_publish(int e, unsigned long long count, unsigned long long prefix, float* __restrict__ weight, double* __restrict__ median) {
  if (count == 0) { weight[e] = 0.0f; median[e] = __longlong_as_double(0x7ff8000000000000LL); }
  else {
    const double nth = __dsqrt_rn(__longlong_as_double((long long)prefix));
    median[e] = nth;
    weight[e] = __double2float_rn(__dmul_rn(nth, 1.5));
  }
}

// one block per edge: the remaining digits (41:31, 30:20, 19:9, 8:0) over the candidates -> exact median -> weight
__global__ void __launch_bounds__(SEL_THREADS)
select_finish_kernel(const EdgeDev* __restrict__ edges, const int32_t* __restrict__ corr, const double* __restrict__ d2,
                     SelState* __restrict__ st, const unsigned long long* __restrict__ cand, unsigned int* __restrict__ cand_n,
                     float* __restrict__ weight, double* __restrict__ median, unsigned long long* __restrict__ win) {
  const int e = blockIdx.x;
  __shared__ unsigned int sh[SEL_BINS];
  __shared__ unsigned long long wtot[SEL_THREADS / 32];
  __shared__ unsigned long long s_prefix, s_rank;
  const unsigned int nc = cand_n[e];
  const EdgeDev ed = edges[e];
  const unsigned long long count = st[e].count;
  if (threadIdx.x == 0) { s_prefix = st[e].prefix; s_rank = st[e].rank; }
  __syncthreads();
  if (count != 0) select_digits(ed, corr, d2, cand + (size_t)e * SEL_CAP, nc, nc > SEL_CAP, 2, sh, wtot, &s_prefix, &s_rank);
  if (threadIdx.x == 0) {
    cand_n[e] = 0u;   // ready for the next round
    st[e].prefix = s_prefix; st[e].rank = s_rank;
    select_publish(e, count, s_prefix, weight, median);
    select_set_window(win, gridDim.x, e, s_prefix, count, false, false, 0u);
  }
}

// one block per edge, after an NN kernel that ran with a guess (knn.cuh, SEL): the wanted rank inside the window -> all six
// digits over the collected keys; anything else (the median left the window, the window overflowed) -> over the edge itself.
__global__ void __launch_bounds__(SEL_THREADS)
select_guess_finish_kernel(const EdgeDev* __restrict__ edges, const int32_t* __restrict__ corr, const double* __restrict__ d2,
                           SelState* __restrict__ st, unsigned int* __restrict__ total, unsigned int* __restrict__ below,
                           const unsigned long long* __restrict__ cand, unsigned int* __restrict__ cand_n,
                           float* __restrict__ weight, double* __restrict__ median, unsigned long long* __restrict__ edge_count,
                           unsigned long long* __restrict__ win, unsigned int* __restrict__ misses, unsigned int* __restrict__ todo_n) {
  const int e = blockIdx.x;
  if (e == 0 && threadIdx.x == 0) *todo_n = 0u;   // the certified round's list of searched queries (knn.cuh) has been consumed
  __shared__ unsigned int sh[SEL_BINS];
  __shared__ unsigned long long wtot[SEL_THREADS / 32];
  __shared__ unsigned long long s_prefix, s_rank;
  const EdgeDev ed = edges[e];
  const unsigned long long count = total[e], bel = below[e], rank = count / 2;
  const unsigned int nc = cand_n[e];
  const bool hit = bel <= rank && rank < bel + nc && nc <= (unsigned int)SEL_CAP;
  if (threadIdx.x == 0) { s_prefix = 0ull; s_rank = hit ? rank - bel : rank; }
  __syncthreads();
  if (count != 0) select_digits(ed, corr, d2, cand + (size_t)e * SEL_CAP, nc, !hit, 0, sh, wtot, &s_prefix, &s_rank);
  if (threadIdx.x == 0) {
    cand_n[e] = 0u; total[e] = 0u; below[e] = 0u;   // ready for the next round
    st[e].prefix = s_prefix; st[e].rank = s_rank; st[e].count = count;
    edge_count[e] = count;
    if (count != 0 && !hit) atomicAdd(misses, 1u);
    select_publish(e, count, s_prefix, weight, median);
    select_set_window(win, gridDim.x, e, s_prefix, count, true, count != 0 && !hit, nc);
  }
}

}  // namespace mv
