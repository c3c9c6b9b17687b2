// closed.cuh -- closed-form pairwise solvers: ICP_Closedform::pointToPoint / pointToPlane
// (src/internal/icp-closedform.cpp:9-54), the comparison baseline of the reference's `pairwise` benchmark and a one-shot
// initialiser.  The O(N) part -- centroids, the 3x3 cross-covariance of the centred clouds, the 6x6 normal equations of
// the linearised point-to-plane problem -- is a streaming reduction on the device; the O(1) part (polar factor of a 3x3
// matrix / 6x6 LDL^T) runs on the host in mvicp_pairwise_closed.  Partials are written per CTA and summed by the host in
// CTA order: the result does not depend on scheduling.
#pragma once
#include <cuda_runtime.h>

namespace mv {

constexpr int CLOSED_THREADS = 256;
constexpr int CLOSED_MAXV = 27;   // values per partial: sums 6 | K 9 | C upper 21 + d 6

// MODE 0: sum p, sum q (6).  MODE 1: K = sum (q - qbar)(p - pbar)^T (9; centroids in aux[0..5]).
// MODE 2: C = sum a a^T (upper triangle, 21), d = -sum a ((p - q).n) (6), a = [p x n ; n].
template <int MODE>
__global__ void __launch_bounds__(CLOSED_THREADS)
closed_reduce_kernel(const double* __restrict__ src, const double* __restrict__ dst, const double* __restrict__ nor, long long n,
                     const double* __restrict__ aux, double* __restrict__ partial /*[gridDim.x][CLOSED_MAXV]*/) {
  constexpr int NV = MODE == 0 ? 6 : (MODE == 1 ? 9 : 27);
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
  double pb[3] = {0, 0, 0}, qb[3] = {0, 0, 0};
  if (MODE == 1) { for (int i = 0; i < 3; ++i) { pb[i] = aux[i]; qb[i] = aux[3 + i]; } }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double p[3] = {src[3 * i], src[3 * i + 1], src[3 * i + 2]};
    const double q[3] = {dst[3 * i], dst[3 * i + 1], dst[3 * i + 2]};
    if (MODE == 0) {
#pragma unroll
      for (int a = 0; a < 3; ++a) { acc[a] += p[a]; acc[3 + a] += q[a]; }
    } else if (MODE == 1) {
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) acc[3 * a + b] += (q[a] - qb[a]) * (p[b] - pb[b]);
    } else {
      const double m[3] = {nor[3 * i], nor[3 * i + 1], nor[3 * i + 2]};
      const double a6[6] = {p[1] * m[2] - p[2] * m[1], p[2] * m[0] - p[0] * m[2], p[0] * m[1] - p[1] * m[0], m[0], m[1], m[2]};
      const double e = (p[0] - q[0]) * m[0] + (p[1] - q[1]) * m[1] + (p[2] - q[2]) * m[2];
      int k = 0;
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = r; c < 6; ++c) acc[k++] += a6[r] * a6[c];
#pragma unroll
      for (int r = 0; r < 6; ++r) acc[21 + r] -= a6[r] * e;
    }
  }
  __shared__ double red[CLOSED_THREADS / 32][CLOSED_MAXV];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int w = 0; w < CLOSED_THREADS / 32; ++w) s += red[w][threadIdx.x];
    partial[(size_t)blockIdx.x * CLOSED_MAXV + threadIdx.x] = s;
  }
}

}  // namespace mv
