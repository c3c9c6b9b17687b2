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
#pragma once
#include <cuda_runtime.h>
#include "types.cuh"

namespace mv {

constexpr int SEL_BINS = 2048;
constexpr int SEL_THREADS = 256;

struct SelState {            // one per edge
  unsigned long long prefix; // bits decided so far (high part)
  unsigned long long rank;   // remaining rank inside the current prefix bucket
  unsigned long long count;  // inliers of the edge
};

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


constexpr int SEL_CAP = 4096;   // collected candidates per edge

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

// one block per edge: the remaining digits (41:31, 30:20, 19:9, 8:0) over the candidates -> exact median -> weight
__global__ void __launch_bounds__(SEL_THREADS)
select_finish_kernel(const EdgeDev* __restrict__ edges, const int32_t* __restrict__ corr, const double* __restrict__ d2,
                     SelState* __restrict__ st, const unsigned long long* __restrict__ cand, unsigned int* __restrict__ cand_n,
                     float* __restrict__ weight, double* __restrict__ median) {
  const int e = blockIdx.x;
  __shared__ unsigned int sh[SEL_BINS];
  __shared__ unsigned int part[SEL_THREADS];
  __shared__ int s_bin;
  __shared__ unsigned long long s_prefix, s_rank, s_before;
  const unsigned int nc = cand_n[e];
  const bool overflow = nc > SEL_CAP;
  const EdgeDev ed = edges[e];
  if (threadIdx.x == 0) { s_prefix = st[e].prefix; s_rank = st[e].rank; }
  __syncthreads();
  const int shifts[4] = {31, 20, 9, 0}, nbits[4] = {11, 11, 11, 9};
  if (st[e].count != 0) {
    for (int p = 0; p < 4; ++p) {
      for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) sh[i] = 0u;
      __syncthreads();
      const int shift = shifts[p], hi = shift + nbits[p];
      const unsigned int mask = (1u << nbits[p]) - 1u;
      const unsigned long long prefix = s_prefix;
      if (!overflow) {
        for (unsigned int i = threadIdx.x; i < nc; i += blockDim.x) {
          const unsigned long long key = cand[(size_t)e * SEL_CAP + i];
          if ((key >> hi) == (prefix >> hi)) atomicAdd(&sh[(unsigned int)(key >> shift) & mask], 1u);
        }
      } else {   // rare: more equal-prefix values than the buffer holds -- scan the edge itself
        for (int k = threadIdx.x; k < ed.n_src; k += blockDim.x) {
          if (corr[ed.off + k] < 0) continue;
          const unsigned long long key = (unsigned long long)__double_as_longlong(d2[ed.off + k]);
          if ((key >> hi) == (prefix >> hi)) atomicAdd(&sh[(unsigned int)(key >> shift) & mask], 1u);
        }
      }
      __syncthreads();
      {   // two-level scan: 8 bins per thread, 256 partial sums scanned by one thread
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
    cand_n[e] = 0u;   // ready for the next round
    st[e].prefix = s_prefix; st[e].rank = s_rank;
    if (st[e].count == 0) { weight[e] = 0.0f; median[e] = __longlong_as_double(0x7ff8000000000000LL); }
    else {
      const double nth = __dsqrt_rn(__longlong_as_double((long long)s_prefix));
      median[e] = nth;
      weight[e] = __double2float_rn(__dmul_rn(nth, 1.5));
    }
  }
}

}  // namespace mv
