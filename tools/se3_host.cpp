// tools/se3_host.cpp -- csrc/se3_math.cuh (host + device header of the engine) compiled by g++ and exported with C linkage
// so that tests/test_se3_math_host.py can compare the PRODUCT's rigid-motion arithmetic with the oracle's, on the CPU.
#include "../mv_lm_icp_b200/csrc/se3_math.cuh"
using namespace mv;
extern "C" {
void h_param_of_pose(int param, const double* P16, double* x) { param_of_pose(param, P16, x); }
void h_pose_of_param(int param, const double* x, double* P16) { pose_of_param(param, x, P16); }
void h_param_plus(int param, const double* x, const double* d, double* o) { param_plus(param, x, d, o); }
void h_Rt_of_param(int param, const double* x, double* R9, double* t3) { Rt a; Rt_of_param(param, x, &a); for (int i = 0; i < 9; ++i) R9[i] = a.R[i]; for (int i = 0; i < 3; ++i) t3[i] = a.t[i]; }
void h_tangent_map(int param, const double* x, double* K36) { Rt a; Rt_of_param(param, x, &a); tangent_map(param, x, &a, K36); }
void h_frame_general(int param, const double* x, double* F9, double* t3, double* D54, double* c18) {
  FrameGen g; frame_general(param, x, &g);
  for (int i = 0; i < 9; ++i) F9[i] = g.F[i];
  for (int i = 0; i < 3; ++i) t3[i] = g.t[i];
  for (int j = 0; j < 6; ++j) { for (int i = 0; i < 9; ++i) D54[9 * j + i] = g.D[j][i]; for (int i = 0; i < 3; ++i) c18[3 * j + i] = g.c[j][i]; }
}
void h_se3_exp(const double* tg, double* out7) { se3_exp(tg, out7); }
void h_se3_compose(const double* a, const double* b, double* o) { se3_compose(a, b, o); }
void h_quat_of_matrix(const double* m, double* q) { quat_of_matrix(m, q); }
void h_matrix_of_quat(const double* q, double* m) { matrix_of_quat(q, m); }
void h_quat_rotate(const double* q, const double* v, double* o) { quat_rotate(q, v, o); }
void h_aa_of_matrix(const double* m, double* aa) { aa_of_matrix(m, aa); }
void h_matrix_of_aa(const double* aa, double* m) { matrix_of_aa(aa, m); }
void h_rotation_of_aa_functor(const double* aa, double* m) { rotation_of_aa_functor(aa, m); }
}
