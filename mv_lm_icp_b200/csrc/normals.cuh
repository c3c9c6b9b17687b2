// normals.cuh -- normal estimation (the step the reference runs on every point before round 0).
//
// Reference: Frame::recomputeNormals (src/internal/frame.cpp:244-255) -> Frame::getNeighbours(i, 10) (:208-242; nanoflann
// knnSearch, the point itself included) -> pointSetPCA (include/common.h:331-346): centroid, cov = sum (p - c)(p - c)^T,
// normal = unit eigenvector of the smallest eigenvalue (SelfAdjointEigenSolver), flipped so that n.z <= 0.
//
// One thread per point, walked in tree order; the k-NN search is the search of knn.cuh (same tree, same fp32 screen /
// fp64 re-rank, start leaf = the point's own leaf) with a sorted k-best list instead of a single best: the k smallest by
// (squared distance, index).  nanoflann orders equal distances by traversal instead (nanoflann.hpp:107-128) -- on real scans,
// whose coordinates sit on a quantised grid, exact ties are common; only ties AT the k-th distance change the neighbour
// set, and the tests compare only points without such a tie.
#pragma once
#include "knn.cuh"

namespace mv {

constexpr int KNN_MAXK = 16;

struct KnnQuery {
  double qx, qy, qz;
  float fx, fy, fz;
  float eaf;
  float bound32;
  int k, count;
  double bd[KNN_MAXK]; int bi[KNN_MAXK];
};

__device__ __forceinline__ void knn_query_init(KnnQuery& s, double qx, double qy, double qz, float absmax, int k) {
  s.qx = qx; s.qy = qy; s.qz = qz;
  s.fx = (float)qx; s.fy = (float)qy; s.fz = (float)qz;
  const double m = fmax(fmax(fabs(qx), fabs(qy)), fabs(qz)) + (double)absmax;
  s.eaf = __double2float_ru(6.0 * 1.7320508075688774 * 1.1920928955078125e-7 * m);
  s.bound32 = __int_as_float(0x7f800000);
  s.k = k; s.count = 0;
}

// candidate that passed the screen: exact distance, sorted insertion by (d, index); the screen bound follows the k-th best
template <bool F32>
__device__ __forceinline__ void nn_exact(const FrameDev& fd, int64_t pos, const float4& r, KnnQuery& s) {
  double px, py, pz; int pi;
  if (F32) { px = (double)r.x; py = (double)r.y; pz = (double)r.z; pi = __float_as_int(r.w); }
  else Rec<false>::load(fd.pts_s, pos, px, py, pz, pi);
  if (pi == INT_MAX) return;   // padding
  const double d = d2_rn(s.qx, s.qy, s.qz, px, py, pz);
  if (s.count == s.k && !(d < s.bd[s.k - 1] || (d == s.bd[s.k - 1] && pi < s.bi[s.k - 1]))) return;
  for (int j = 0; j < s.count; ++j) if (s.bi[j] == pi) return;   // a leaf can be reached twice (start leaf, then through the tree)
  int i = s.count < s.k ? s.count : s.k - 1;
  while (i > 0 && (s.bd[i - 1] > d || (s.bd[i - 1] == d && s.bi[i - 1] > pi))) { s.bd[i] = s.bd[i - 1]; s.bi[i] = s.bi[i - 1]; --i; }
  s.bd[i] = d; s.bi[i] = pi;
  if (s.count < s.k) ++s.count;
  if (s.count == s.k) {
    const float b = __double2float_ru(s.bd[s.k - 1]);
    const float rr = __fsqrt_ru(b);
    s.bound32 = __fmul_ru(__fmaf_ru(s.eaf, __fmaf_ru(2.0f, rr, s.eaf), b), 1.000001f);
  }
}

// symmetric 3x3 eigen-decomposition by cyclic Jacobi rotations; returns the unit eigenvector of the smallest eigenvalue
__device__ __forceinline__ void smallest_eigvec3(double A[3][3], double* v) {
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 50; ++sweep) {
    const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        const double th = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
        for (int k = 0; k < 3; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
        for (int k = 0; k < 3; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
      }
  }
  int m = 0;
  if (A[1][1] < A[m][m]) m = 1;
  if (A[2][2] < A[m][m]) m = 2;
  v[0] = V[0][m]; v[1] = V[1][m]; v[2] = V[2][m];
}

template <bool F32>
__global__ void __launch_bounds__(128)
normals_kernel(const FrameDev* __restrict__ frames, int frame, int k, double* __restrict__ nor_out /*[n][3], caller's order*/,
               int32_t* __restrict__ nn_out /*nullable [n][k]*/) {
  const FrameDev fd = frames[frame];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= fd.n) return;
  double px, py, pz; int orig;
  Rec<F32>::load(fd.pts_s, t, px, py, pz, orig);
  KnnQuery s; knn_query_init(s, px, py, pz, fd.absmax, k);
  nn_search<F32, KnnQuery>(fd, s, t / LEAF);
  // pointSetPCA over the neighbours in knnSearch order
  const int m = s.count;
  double c[3] = {0, 0, 0};
  double nx[KNN_MAXK], ny[KNN_MAXK], nz[KNN_MAXK];
  for (int j = 0; j < m; ++j) {
    int dummy; Rec<F32>::load(fd.pts_o, s.bi[j], nx[j], ny[j], nz[j], dummy);
    c[0] += nx[j]; c[1] += ny[j]; c[2] += nz[j];
  }
  c[0] /= m; c[1] /= m; c[2] /= m;
  double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int j = 0; j < m; ++j) {
    const double d[3] = {nx[j] - c[0], ny[j] - c[1], nz[j] - c[2]};
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) C[a][b] += d[a] * d[b];
  }
  double v[3]; smallest_eigvec3(C, v);
  if (v[2] > 0) { v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2]; }   // "flip towards camera", common.h:343
  nor_out[3 * (size_t)orig] = v[0]; nor_out[3 * (size_t)orig + 1] = v[1]; nor_out[3 * (size_t)orig + 2] = v[2];
  if (nn_out) for (int j = 0; j < k; ++j) nn_out[(size_t)k * orig + j] = j < m ? s.bi[j] : -1;
}

// normals (fp64, caller's order) -> the 32-byte records the LM kernels gather
__global__ void pack_normals_kernel(const double* __restrict__ nor, int n, double4a* __restrict__ rec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4a r; r.x = nor[3 * (size_t)i]; r.y = nor[3 * (size_t)i + 1]; r.z = nor[3 * (size_t)i + 2]; r.w = 0.0;
  rec[i] = r;
}

// fp32 storage: point + normal of every point in one 32-byte record (the LM streaming kernel gathers both per match)
__global__ void pack_pn_kernel(const float4* __restrict__ pts, const float4* __restrict__ nor, int n, float4* __restrict__ pn) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pn[2 * (size_t)i] = pts[i]; pn[2 * (size_t)i + 1] = nor[i];
}

}  // namespace mv
