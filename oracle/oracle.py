"""ctypes front-end of the CPU oracle (oracle/liboracle.so, oracle/_ref/libref_nanoflann.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (mv_lm_icp_b200) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PARAM_AA, PARAM_QUAT, PARAM_SE3 = 0, 1, 2
COST_P2P, COST_P2PLANE, COST_MIXED = 0, 1, 2
TERMINATION = ["CONVERGENCE_FUNCTION", "CONVERGENCE_GRADIENT", "CONVERGENCE_PARAMETER", "NO_CONVERGENCE_MAX_ITER",
               "CONVERGENCE_MIN_RADIUS", "FAILURE_INVALID_STEPS", "FAILURE_EVAL"]


class LmOptions(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int32), ("max_num_consecutive_invalid_steps", C.c_int32),
                ("jacobi_scaling", C.c_int32), ("_pad", C.c_int32),
                ("initial_trust_region_radius", C.c_double), ("max_trust_region_radius", C.c_double),
                ("min_trust_region_radius", C.c_double), ("min_relative_decrease", C.c_double),
                ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double)]


class LmSummary(C.Structure):
    _fields_ = [("termination", C.c_int32), ("num_iterations", C.c_int32), ("num_successful_steps", C.c_int32),
                ("num_jacobian_evals", C.c_int32), ("num_cost_evals", C.c_int32), ("n_trace", C.c_int32),
                ("initial_cost", C.c_double), ("final_cost", C.c_double)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build(force=False):
    """Compile liboracle.so (+ _ref when /root/reference is present). Idempotent."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_icp.cpp", "geom.h", "jet.h", "lm.h", "ref_nanoflann.cpp", "ref_functors.cpp", "Makefile")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    ref_so = os.path.join(_HERE, "_ref", "libref_nanoflann.so")
    fun_so = os.path.join(_HERE, "_ref", "libref_functors.so")
    if (not os.path.exists(ref_so) or not os.path.exists(fun_so)) and os.path.exists("/root/reference/include/nanoflann.hpp"):
        stale = True
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s", "all"], check=True, stdout=subprocess.DEVNULL)
    return so


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_kd_build.restype = C.c_void_p
        _lib.orc_filter_edge.restype = C.c_int64
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def ref_lib():
    """The reference's own nanoflann behind a shim; None if it was never built (no /root/reference)."""
    global _ref
    if _ref is None:
        build()
        p = os.path.join(_HERE, "_ref", "libref_nanoflann.so")
        if not os.path.exists(p):
            return None
        _ref = C.CDLL(p)
        _ref.ref_kd_build.restype = C.c_void_p
    return _ref


_fun = None


def ref_functors():
    """The reference's own cost-functor text (include/icp-ceres.h, include/eigen_quaternion.h) compiled against oracle/stubs
    (oracle/ref_functors.cpp); None if it was never built (no /root/reference at build time)."""
    global _fun
    if _fun is None:
        build()
        p = os.path.join(_HERE, "_ref", "libref_functors.so")
        if not os.path.exists(p):
            return None
        _fun = C.CDLL(p)
    return _fun


def functor_eval(param, plane, cam1, cam2, src, dst, nor, which="oracle"):
    """(residuals[NR], ambient Jacobian[NR, 2G]) of one multiview functor: which = 'oracle' (restatement) or 'ref' (reference text)."""
    G = 6 if param == PARAM_AA else 7
    nr = 1 if plane else 3
    r = np.zeros(3); jac = np.zeros(3 * 2 * G)
    fn = lib().orc_functor_eval if which == "oracle" else ref_functors().ref_functor_global
    fn(C.c_int(param), C.c_int(int(plane)), _p(_f64(cam1)), _p(_f64(cam2)), _p(_f64(src)), _p(_f64(dst)), _p(_f64(nor)), _p(r), _p(jac))
    return r[:nr].copy(), jac[:nr * 2 * G].reshape(nr, 2 * G).copy()


def functor_eval_pairwise_ref(param, plane, cam, src, dst, nor):
    G = 6 if param == PARAM_AA else 7
    nr = 1 if plane else 3
    r = np.zeros(3); jac = np.zeros(3 * G)
    ref_functors().ref_functor_pairwise(C.c_int(param), C.c_int(int(plane)), _p(_f64(cam)), _p(_f64(src)), _p(_f64(dst)), _p(_f64(nor)), _p(r), _p(jac))
    return r[:nr].copy(), jac[:nr * G].reshape(nr, G).copy()


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def max_threads():
    return lib().orc_max_threads()


class KdIndex:
    """Exact 1-NN index over a cloud. kind: 'kd' (oracle restatement), 'brute', 'ref' (reference nanoflann)."""

    def __init__(self, pts, kind="kd"):
        self.pts = _f64(pts).reshape(-1, 3)
        self.kind = kind
        self.h = None
        if kind == "kd":
            self.h = C.c_void_p(lib().orc_kd_build(_p(self.pts), C.c_int64(len(self.pts))))
        elif kind == "ref":
            r = ref_lib()
            if r is None:
                raise RuntimeError("oracle/_ref/libref_nanoflann.so not built")
            self.h = C.c_void_p(r.ref_kd_build(_p(self.pts), C.c_int64(len(self.pts))))

    def __del__(self):
        try:
            if self.h is not None:
                (lib().orc_kd_free if self.kind == "kd" else ref_lib().ref_kd_free)(self.h)
        except Exception:
            pass

    def closest_points(self, src_pts, pose_src, pose_dst, threads=1):
        """Restated frame.cpp:129-138 for one edge. Poses are 4x4 (row/col indexed numpy) matrices."""
        src = _f64(src_pts).reshape(-1, 3)
        n = len(src)
        ps = _f64(np.asarray(pose_src).T).reshape(16)   # column-major double[16] like Isometry3d
        pd = _f64(np.asarray(pose_dst).T).reshape(16)
        idx = np.empty(n, np.int32)
        d2 = np.empty(n, np.float64)
        if self.kind == "ref":
            ref_lib().ref_closest_points(self.h, _p(src), C.c_int64(n), _p(ps), _p(pd), _p(idx, C.c_int32), _p(d2),
                                         C.c_int(threads))
        else:
            lib().orc_closest_points(self.h if self.kind == "kd" else None, _p(self.pts), C.c_int64(len(self.pts)),
                                     _p(src), C.c_int64(n), _p(ps), _p(pd), _p(idx, C.c_int32), _p(d2), None,
                                     C.c_int(threads))
        return idx, d2


def recompute_normals(pts, k=10, threads=1, want_nn=False):
    """Restated Frame::recomputeNormals (frame.cpp:244-255, common.h:331-346). Returns normals [N,3] (and the k-NN indices)."""
    pts = _f64(pts).reshape(-1, 3)
    n = len(pts)
    nor = np.empty((n, 3)); nn = np.empty((n, k), np.int32)
    lib().orc_recompute_normals(_p(pts), C.c_int64(n), C.c_int(k), _p(nor), _p(nn, C.c_int32) if want_nn else None, C.c_int(threads))
    return (nor, nn) if want_nn else nor


def knn(index, q, k):
    """k nearest neighbours of q in a KdIndex ('kd' oracle restatement or 'ref' = the reference's nanoflann knnSearch)."""
    q = _f64(q)
    if index.kind == "ref":
        idx = np.empty(k, np.int64); d2 = np.empty(k)
        ref_lib().ref_kd_knn(index.h, _p(q), C.c_int64(k), _p(idx, C.c_int64), _p(d2))
        return idx.astype(np.int32), d2
    idx = np.empty(k, np.int32); d2 = np.empty(k)
    lib().orc_kd_knn(index.h, _p(q), C.c_int(k), _p(idx, C.c_int32), _p(d2))
    return idx, d2


def edge_queries(src_pts, pose_src, pose_dst):
    src = _f64(src_pts).reshape(-1, 3)
    ps = _f64(np.asarray(pose_src).T).reshape(16)
    pd = _f64(np.asarray(pose_dst).T).reshape(16)
    out = np.empty_like(src)
    for k in range(len(src)):
        lib().orc_edge_transform(_p(ps), _p(pd), _p(src[k]), _p(out[k]))
    return out


def filter_edge(nn_idx, nn_d2, thresh):
    """frame.cpp:140-176 -> (first, second, dist, weight(float32), median)."""
    n = len(nn_idx)
    nn_idx = np.ascontiguousarray(nn_idx, np.int32)
    nn_d2 = _f64(nn_d2)
    first = np.empty(n, np.int32); second = np.empty(n, np.int32); dist = np.empty(n, np.float64)
    w = C.c_float(0); med = C.c_double(0)
    c = lib().orc_filter_edge(_p(nn_idx, C.c_int32), _p(nn_d2), C.c_int64(n), C.c_float(thresh), _p(first, C.c_int32),
                              _p(second, C.c_int32), _p(dist), C.byref(w), C.byref(med))
    return first[:c].copy(), second[:c].copy(), dist[:c].copy(), np.float32(w.value), med.value


def pose_graph_knn(poses, knn):
    """frame.cpp:67-89 -> int32 [M, knn] neighbour indices."""
    M = len(poses)
    P = _f64(np.stack([np.asarray(p).T.reshape(16) for p in poses]))
    out = np.empty((M, knn), np.int32)
    lib().orc_pose_graph_knn(C.c_int(M), _p(P), C.c_int(knn), _p(out, C.c_int32), None)
    return out


def _problem_args(pts, nor, poses, edges, corr, weights, fixed):
    M = len(pts)
    pts_c = [_f64(p).reshape(-1, 3) for p in pts]
    nor_c = [None if n is None else _f64(n).reshape(-1, 3) for n in nor]
    PP = (C.POINTER(C.c_double) * M)(*[_p(p) for p in pts_c])
    NN = (C.POINTER(C.c_double) * M)(*[(_p(n) if n is not None else None) for n in nor_c])
    P = _f64(np.stack([np.asarray(p).T.reshape(16) for p in poses]))
    E = len(edges)
    es = np.ascontiguousarray([e[0] for e in edges], np.int32)
    ed = np.ascontiguousarray([e[1] for e in edges], np.int32)
    off = np.zeros(E + 1, np.int64)
    for i, (f, s) in enumerate(corr):
        off[i + 1] = off[i] + len(f)
    first = np.ascontiguousarray(np.concatenate([np.asarray(f, np.int32) for f, _ in corr]) if E else np.zeros(0), np.int32)
    second = np.ascontiguousarray(np.concatenate([np.asarray(s, np.int32) for _, s in corr]) if E else np.zeros(0), np.int32)
    w = np.ascontiguousarray(weights if weights is not None else np.zeros(E), np.float32)
    fx = np.zeros(M, np.uint8)
    if fixed is not None:
        fx[:] = np.asarray(fixed, np.uint8)
    keep = (pts_c, nor_c, es, ed, off, first, second, w, fx)
    return M, PP, NN, P, E, es, ed, off, first, second, w, fx, keep


def optimize(pts, nor, poses, edges, corr, weights, param=PARAM_SE3, cost=COST_P2PLANE, robust=True,
             se3_autodiff=True, threads=1, options=None, fixed=None, max_trace=64):
    """Restated ICP_Ceres::ceresOptimizer* (icp-ceres.cpp:220-475). Returns (poses[M,4,4], summary dict, trace)."""
    M, PP, NN, P, E, es, ed, off, first, second, w, fx, _keep = _problem_args(pts, nor, poses, edges, corr, weights, fixed)
    summ = LmSummary()
    trace = np.zeros((max_trace, 10), np.float64)
    lib().orc_optimize(C.c_int(M), PP, NN, _p(fx, C.c_uint8), _p(P), C.c_int(E), _p(es, C.c_int32), _p(ed, C.c_int32),
                       _p(off, C.c_int64), _p(first, C.c_int32), _p(second, C.c_int32), _p(w, C.c_float),
                       C.c_int(param), C.c_int(cost), C.c_int(int(robust)), C.c_int(int(se3_autodiff)),
                       C.c_int(threads), C.byref(options) if options is not None else None, C.byref(summ),
                       _p(trace), C.c_int(max_trace))
    out = np.stack([P[i].reshape(4, 4).T for i in range(M)])
    d = summ.asdict()
    return out, d, trace[:min(d["n_trace"], max_trace)].copy()


def evaluate(pts, nor, poses, edges, corr, weights, param=PARAM_SE3, cost=COST_P2PLANE, robust=True,
             se3_autodiff=True, threads=1, fixed=None, want_jac=True):
    M, PP, NN, P, E, es, ed, off, first, second, w, fx, _keep = _problem_args(pts, nor, poses, edges, corr, weights, fixed)
    fx[0] = 1
    n = 6 * int((fx == 0).sum())
    H = np.zeros((n, n)); g = np.zeros(n); c = C.c_double(0)
    lib().orc_evaluate(C.c_int(M), PP, NN, _p(fx, C.c_uint8), _p(P), C.c_int(E), _p(es, C.c_int32), _p(ed, C.c_int32),
                       _p(off, C.c_int64), _p(first, C.c_int32), _p(second, C.c_int32), _p(w, C.c_float),
                       C.c_int(param), C.c_int(cost), C.c_int(int(robust)), C.c_int(int(se3_autodiff)),
                       C.c_int(threads), C.byref(c), _p(H) if want_jac else None, _p(g) if want_jac else None)
    return c.value, H, g


def pairwise(src, dst, nor=None, param=PARAM_SE3, cost=COST_P2P, se3_autodiff=False, threads=1, options=None):
    """Restated ICP_Ceres::pointToPoint_* / pointToPlane_* (icp-ceres.cpp:137-218,525-565)."""
    src = _f64(src).reshape(-1, 3); dst = _f64(dst).reshape(-1, 3)
    nr = None if nor is None else _f64(nor).reshape(-1, 3)
    out = np.zeros(16); summ = LmSummary()
    lib().orc_pairwise(_p(src), _p(dst), _p(nr) if nr is not None else None, C.c_int64(len(src)), C.c_int(param),
                       C.c_int(cost), C.c_int(int(se3_autodiff)), C.c_int(threads),
                       C.byref(options) if options is not None else None, _p(out), C.byref(summ))
    return out.reshape(4, 4).T.copy(), summ.asdict()


def closed_form(src, dst, nor=None):
    """Restated ICP_Closedform::pointToPoint (nor is None) / pointToPlane (icp-closedform.cpp:9-54): 4x4 src -> dst."""
    src = _f64(src).reshape(-1, 3); dst = _f64(dst).reshape(-1, 3)
    out = np.zeros(16)
    if nor is None:
        lib().orc_closed_p2p(_p(src), _p(dst), C.c_int64(len(src)), _p(out))
    else:
        nr = _f64(nor).reshape(-1, 3)
        lib().orc_closed_p2plane(_p(src), _p(dst), _p(nr), C.c_int64(len(src)), _p(out))
    return out.reshape(4, 4).T.copy()


def default_options():
    o = LmOptions()
    lib().orc_default_lm_options(C.byref(o))
    return o


# ---- small math exports (known-answer tests) ----------------------------------------------------
def _call(name, outn, *ins):
    arrs = [_f64(a).ravel() for a in ins]
    out = np.zeros(outn)
    getattr(lib(), name)(*[_p(a) for a in arrs], _p(out))
    return out


def se3_exp(t6): return _call("orc_se3_exp", 7, t6)
def se3_mul(a7, b7): return _call("orc_se3_mul", 7, a7, b7)
def se3_plus(x7, d6): return _call("orc_se3_plus", 7, x7, d6)
def se3_internal_jacobian(x7): return _call("orc_se3_internal_jacobian", 42, x7).reshape(7, 6)
def se3_plus_jacobian_autodiff(x7): return _call("orc_se3_plus_jacobian_autodiff", 42, x7).reshape(7, 6)
def quat_from_matrix(R): return _call("orc_quat_from_matrix", 4, np.asarray(R).reshape(9))
def quat_to_matrix(q): return _call("orc_quat_to_matrix", 9, q).reshape(3, 3)
def quat_transform(q, v): return _call("orc_quat_transform", 3, q, v)
def quat_plus(x, d): return _call("orc_quat_plus", 4, x, d)
def angle_axis_rotate(aa, p): return _call("orc_angle_axis_rotate", 3, aa, p)
def rotmat_to_angle_axis(R): return _call("orc_rotmat_to_angle_axis", 3, np.asarray(R).reshape(9))
def angle_axis_to_rotmat(aa): return _call("orc_angle_axis_to_rotmat", 9, aa).reshape(3, 3)
def mat3_inverse(m): return _call("orc_mat3_inverse", 9, np.asarray(m).reshape(9)).reshape(3, 3)


def pose_to_param(P, param):
    out = np.zeros(7)
    lib().orc_pose_to_param(_p(_f64(np.asarray(P).T).reshape(16)), C.c_int(param), _p(out))
    return out[:6] if param == PARAM_AA else out


def param_to_pose(x, param):
    out = np.zeros(16)
    lib().orc_param_to_pose(_p(_f64(x)), C.c_int(param), _p(out))
    return out.reshape(4, 4).T.copy()
