// compact.cuh -- the inlier lists of every edge as the reference's own records, built on the device.
//
// Reference: Frame::computeClosestPointsToNeighbours pushes {k, idxMin, dist} for every src point whose nearest neighbour lies
// within the cutoff, in ascending src index (src/internal/frame.cpp:156-160) into OutgoingEdge::correspondances
// (include/frame.h:18-29).  The engine keeps one (match, squared distance) slot per src point; a caller that wants the lists
// on the host (the viewer draws them, Visualize.cpp:470-479) gets them from three small kernels -- inliers per 256-slot tile,
// exclusive scan over the tiles, ordered scatter of 16-byte records -- and ONE device-to-host copy, instead of copying every
// slot and filtering on the CPU.
#pragma once
#include <cuda_runtime.h>
#include "types.cuh"

namespace mv {

struct CorrRec { int32_t first, second; double dist; };   // == struct Correspondance (include/frame.h:18-22)
static_assert(sizeof(CorrRec) == 16, "Correspondance is 16 bytes");

__global__ void __launch_bounds__(KNN_TILE)
compact_count_kernel(const EdgeDev* __restrict__ edges, const Tile* __restrict__ tiles, const int32_t* __restrict__ corr,
                     unsigned int* __restrict__ tile_count) {
  const Tile t = tiles[blockIdx.x];
  const EdgeDev e = edges[t.edge];
  const int k = t.start + threadIdx.x;
  const bool in = k < e.n_src && corr[e.off + k] >= 0;
  const int c = __syncthreads_count(in);
  if (threadIdx.x == 0) tile_count[blockIdx.x] = (unsigned int)c;
}

// exclusive scan of the tile counts (one block; tiles are in edge order, so an edge's list is contiguous) and the
// per-edge offsets: edge_off[e] = first record of edge e, edge_off[E] = total
__global__ void __launch_bounds__(1024)
compact_scan_kernel(const unsigned int* __restrict__ tile_count, int n_tiles, const Tile* __restrict__ tiles, int n_edges,
                    unsigned long long* __restrict__ tile_off, unsigned long long* __restrict__ edge_off) {
  __shared__ unsigned long long s_part[1024];
  __shared__ unsigned long long s_carry;
  const int T = blockDim.x, tid = threadIdx.x;
  if (tid == 0) s_carry = 0ull;
  for (int e = tid; e <= n_edges; e += T) edge_off[e] = ~0ull;
  __syncthreads();
  for (int base = 0; base < n_tiles; base += T) {
    const int i = base + tid;
    const unsigned long long v = i < n_tiles ? tile_count[i] : 0ull;
    s_part[tid] = v;
    __syncthreads();
    for (int o = 1; o < T; o <<= 1) {   // Hillis-Steele inclusive scan
      const unsigned long long a = tid >= o ? s_part[tid - o] : 0ull;
      __syncthreads();
      s_part[tid] += a;
      __syncthreads();
    }
    const unsigned long long excl = s_carry + s_part[tid] - v;
    if (i < n_tiles) {
      tile_off[i] = excl;
      if (tiles[i].start == 0) edge_off[tiles[i].edge] = excl;   // first tile of its edge
    }
    __syncthreads();
    if (tid == T - 1) s_carry += s_part[T - 1];
    __syncthreads();
  }
  if (tid == 0) {   // edges without tiles (fixed src / other rank): empty lists at the position of the next edge that has one
    unsigned long long next = s_carry;
    edge_off[n_edges] = next;
    for (int e = n_edges - 1; e >= 0; --e) { if (edge_off[e] == ~0ull) edge_off[e] = next; else next = edge_off[e]; }
  }
}

__global__ void __launch_bounds__(KNN_TILE)
compact_scatter_kernel(const EdgeDev* __restrict__ edges, const Tile* __restrict__ tiles, const int32_t* __restrict__ corr,
                       const double* __restrict__ d2, const unsigned long long* __restrict__ tile_off, CorrRec* __restrict__ out) {
  const Tile t = tiles[blockIdx.x];
  const EdgeDev e = edges[t.edge];
  const int k = t.start + threadIdx.x;
  int c = -1;
  if (k < e.n_src) c = corr[e.off + k];
  const bool in = c >= 0;
  __shared__ int s_warp[KNN_TILE / 32];
  const unsigned m = __ballot_sync(0xffffffffu, in);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) s_warp[w] = __popc(m);
  __syncthreads();
  int before = 0;
  for (int i = 0; i < w; ++i) before += s_warp[i];
  if (in) {
    CorrRec r; r.first = k; r.second = c; r.dist = __dsqrt_rn(d2[e.off + k]);   // dist = sqrt(d2), frame.cpp:142
    out[tile_off[blockIdx.x] + before + __popc(m & ((1u << lane) - 1u))] = r;
  }
}

}  // namespace mv
