"""Python host mirror of the reference's frame / optimiser interface, on top of the C ABI.

Reference names kept: Frame.{pts,nor,pose,fixed,neighbours}, Frame.computePoseNeighboursKnn,
Frame.computeClosestPointsToNeighbours (include/frame.h:38-55), ICP_Ceres.ceresOptimizer{,_ceresAngleAxis,_sophusSE3}
and the pairwise pointToPoint_* / pointToPlane_* (include/icp-ceres.h:30-42)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Config, LmOptions, LmSummary, Stats, check

PARAM_AA, PARAM_QUAT, PARAM_SE3 = 0, 1, 2
COST_P2P, COST_P2PLANE, COST_MIXED = 0, 1, 2
TERMINATION = ["FUNCTION_TOLERANCE", "GRADIENT_TOLERANCE", "PARAMETER_TOLERANCE", "MAX_ITERATIONS", "MIN_RADIUS",
               "INVALID_STEPS", "EVAL_FAILURE"]
FLAG_NO_SEED = 1
FLAG_NCCL_ONLY = 2
FLAG_HOST_BUILD = 4
FLAG_NO_ADJ = 8
FLAG_NO_OBB = 16
FLAG_STEP_LOOP = 32
FLAG_NO_SELECT_GUESS = 64
FLAG_NO_CERT = 128


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _pose16(P):
    return _f64(np.asarray(P, dtype=np.float64).T).reshape(16)   # column-major, as Isometry3d::data()


def nccl_unique_id():
    buf = (C.c_char * 128)()
    check(_lib.lib().mvicp_nccl_unique_id(buf))
    return bytes(buf)


class Engine:
    """One mvicp_ctx (one GPU, one host thread at a time)."""

    def __init__(self, device=0, flags=0, stream=None):
        self._l = _lib.lib()
        self._ctx = C.c_void_p()
        cfg = Config(device, flags, stream)
        check(self._l.mvicp_create(C.byref(cfg), C.byref(self._ctx)))
        self.M = 0
        self.n_pts = []
        self._keep = None

    def _free_pinned(self):
        if getattr(self, "_rec_ptr", None) is not None and self._rec_ptr.value:
            self.host_edges = None; self._rec_buf = None
            self._l.mvicp_host_free(self._rec_ptr)
        self._rec_ptr = None

    def close(self):
        if self._ctx:
            self._free_pinned()
            self._l.mvicp_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- data ---------------------------------------------------------------------------------------
    def set_frames(self, pts, nor=None):
        M = len(pts)
        P = [_f64(p).reshape(-1, 3) for p in pts]
        N = None if nor is None else [None if n is None else _f64(n).reshape(-1, 3) for n in nor]
        PP = (C.POINTER(C.c_double) * M)(*[_p(p) for p in P])
        NN = None
        if N is not None:
            NN = (C.POINTER(C.c_double) * M)(*[(_p(n) if n is not None else None) for n in N])
        n = np.ascontiguousarray([len(p) for p in P], np.int64)
        check(self._l.mvicp_set_frames(self._ctx, C.c_int32(M), PP, NN, _p(n, C.c_int64)))
        self.M = M
        self.n_pts = [len(p) for p in P]

    def set_poses(self, poses, fixed=None):
        P = _f64(np.stack([_pose16(p) for p in poses]))
        fx = None if fixed is None else np.ascontiguousarray(fixed, np.uint8)
        check(self._l.mvicp_set_poses(self._ctx, _p(P), _p(fx, C.c_uint8) if fx is not None else None))

    def get_poses(self):
        P = np.zeros((self.M, 16))
        check(self._l.mvicp_get_poses(self._ctx, _p(P)))
        return np.stack([P[i].reshape(4, 4).T.copy() for i in range(self.M)])

    def set_graph(self, edges):
        E = len(edges)
        s = np.ascontiguousarray([e[0] for e in edges], np.int32)
        d = np.ascontiguousarray([e[1] for e in edges], np.int32)
        check(self._l.mvicp_set_graph(self._ctx, C.c_int32(E), _p(s, C.c_int32), _p(d, C.c_int32)))
        self.edges = [(int(a), int(b)) for a, b in edges]

    def pose_graph_knn(self, knn):
        check(self._l.mvicp_pose_graph_knn(self._ctx, C.c_int32(knn)))
        E = C.c_int32(0)
        check(self._l.mvicp_get_graph(self._ctx, C.byref(E), None, None))
        s = np.zeros(E.value, np.int32); d = np.zeros(E.value, np.int32)
        check(self._l.mvicp_get_graph(self._ctx, C.byref(E), _p(s, C.c_int32), _p(d, C.c_int32)))
        self.edges = list(zip(s.tolist(), d.tolist()))
        return self.edges

    # ---- hot path -----------------------------------------------------------------------------------
    def correspond(self, thresh=0.05):
        check(self._l.mvicp_correspond(self._ctx, C.c_float(thresh)))

    def get_edge(self, e, arrays=True):
        n = self.n_pts[self.edges[e][0]]
        cnt = C.c_int64(0); w = C.c_float(0)
        if not arrays:
            check(self._l.mvicp_get_edge(self._ctx, C.c_int32(e), None, None, None, C.byref(cnt), C.byref(w)))
            return cnt.value, np.float32(w.value)
        first = np.empty(n, np.int32); second = np.empty(n, np.int32); dist = np.empty(n, np.float64)
        check(self._l.mvicp_get_edge(self._ctx, C.c_int32(e), _p(first, C.c_int32), _p(second, C.c_int32), _p(dist),
                                     C.byref(cnt), C.byref(w)))
        c = cnt.value
        return first[:c].copy(), second[:c].copy(), dist[:c].copy(), np.float32(w.value)

    CORR_DTYPE = np.dtype([("first", np.int32), ("second", np.int32), ("dist", np.float64)])   # struct Correspondance (frame.h:18-22)

    def pull_all_edges(self, records=True):
        """Every edge's (first, second, dist) list and weight into host arrays, as Frame::computeClosestPointsToNeighbours
        leaves them in OutgoingEdge::correspondances (frame.cpp:158,176): one structured array + offsets (mvicp_get_all_edges).
        Returns the number of bytes that reached the host; self.host_edges[e] = (records view, weight)."""
        E = len(self.edges)
        off = np.zeros(E + 1, np.int64); w = np.zeros(E, np.float32)
        cap = sum(self.n_pts[s] for s, _ in self.edges)
        if records:
            if getattr(self, "_rec_buf", None) is None or len(self._rec_buf) < cap:
                self._free_pinned()
                ptr = C.c_void_p()
                check(self._l.mvicp_host_alloc(C.c_size_t(cap * self.CORR_DTYPE.itemsize), C.byref(ptr)))   # page-locked: the copy runs at link speed
                self._rec_ptr = ptr
                self._rec_buf = np.frombuffer((C.c_char * (cap * self.CORR_DTYPE.itemsize)).from_address(ptr.value), dtype=self.CORR_DTYPE) if cap else np.empty(0, self.CORR_DTYPE)
            check(self._l.mvicp_get_all_edges(self._ctx, self._rec_buf.ctypes.data_as(C.c_void_p), C.c_int64(cap), _p(off, C.c_int64), _p(w, C.c_float)))
            self.host_edges = [(self._rec_buf[off[e]:off[e + 1]], w[e]) for e in range(E)]
        else:
            check(self._l.mvicp_get_all_edges(self._ctx, None, C.c_int64(0), _p(off, C.c_int64), _p(w, C.c_float)))
            self.host_edges = [(None, w[e]) for e in range(E)]
        self.edge_offsets = off
        return int(off[E]) * 16 * int(records) + 4 * E + 8 * (E + 1)

    def get_nn(self, e):
        n = self.n_pts[self.edges[e][0]]
        idx = np.empty(n, np.int32); d2 = np.empty(n, np.float64)
        check(self._l.mvicp_get_nn(self._ctx, C.c_int32(e), _p(idx, C.c_int32), _p(d2)))
        return idx, d2

    def set_edge(self, e, first, second, weight):
        f = np.ascontiguousarray(first, np.int32); s = np.ascontiguousarray(second, np.int32)
        check(self._l.mvicp_set_edge(self._ctx, C.c_int32(e), _p(f, C.c_int32), _p(s, C.c_int32), C.c_int64(len(f)),
                                     C.c_float(weight)))

    def closest_point(self, frame, q):
        q = _f64(q); idx = C.c_int64(0); d2 = C.c_double(0)
        check(self._l.mvicp_closest_point(self._ctx, C.c_int32(frame), _p(q), C.byref(idx), C.byref(d2)))
        return idx.value, d2.value

    def optimize(self, param=PARAM_SE3, cost=COST_P2PLANE, robust=True, options=None):
        s = LmSummary()
        check(self._l.mvicp_optimize(self._ctx, C.c_int32(param), C.c_int32(cost), C.c_int32(int(robust)),
                                     C.byref(options) if options is not None else None, C.byref(s)))
        return s.asdict()

    def icp_round(self, thresh=0.05, param=PARAM_SE3, cost=COST_P2PLANE, robust=True, options=None):
        s = LmSummary()
        check(self._l.mvicp_icp_round(self._ctx, C.c_float(thresh), C.c_int32(param), C.c_int32(cost),
                                      C.c_int32(int(robust)), C.byref(options) if options is not None else None, C.byref(s)))
        return s.asdict()

    def recompute_normals(self, k=10, fetch=True):
        """Frame::recomputeNormals for every frame (frame.cpp:244-255). Returns (list of [N,3] normals, device ms);
        fetch=False leaves the normals on the device (list is None)."""
        check(self._l.mvicp_recompute_normals(self._ctx, C.c_int32(k)))
        out = []; ms = C.c_float(0)
        if not fetch:
            nor = np.empty((self.n_pts[0], 3))
            check(self._l.mvicp_get_normals(self._ctx, C.c_int32(0), _p(nor), C.byref(ms)))
            return None, ms.value
        for f in range(self.M):
            nor = np.empty((self.n_pts[f], 3))
            check(self._l.mvicp_get_normals(self._ctx, C.c_int32(f), _p(nor), C.byref(ms)))
            out.append(nor)
        return out, ms.value

    def knn_self(self, frame, k=10):
        """Frame::getNeighbours for every point of a frame (frame.cpp:208-242): int32 [N, k]."""
        nn = np.empty((self.n_pts[frame], k), np.int32)
        check(self._l.mvicp_knn_self(self._ctx, C.c_int32(frame), C.c_int32(k), _p(nn, C.c_int32)))
        return nn

    # ---- multi-GPU / introspection ---------------------------------------------------------------------
    def comm_init(self, unique_id, rank, world):
        check(self._l.mvicp_comm_init(self._ctx, C.c_char_p(unique_id), C.c_int32(rank), C.c_int32(world)))

    def stats(self):
        s = Stats()
        check(self._l.mvicp_get_stats(self._ctx, C.byref(s)))
        return s.asdict()

    def stream(self):
        p = C.c_void_p()
        check(self._l.mvicp_get_stream(self._ctx, C.byref(p)))
        return p.value

    def sync(self):
        check(self._l.mvicp_sync(self._ctx))


def default_options():
    o = LmOptions()
    _lib.lib().mvicp_default_lm_options(C.byref(o))
    return o


class OutgoingEdge:
    """include/frame.h:24-29"""

    def __init__(self, neighbourIdx, weight=0.0):
        self.neighbourIdx = neighbourIdx
        self.weight = np.float32(weight)
        self.correspondances = []   # list of (first, second, dist)


class Frame:
    """include/frame.h:31-102 (hot-path members only)."""

    def __init__(self, pts, nor=None, pose=None, fixed=False):
        self.pts = _f64(pts).reshape(-1, 3)
        self.nor = None if nor is None else _f64(nor).reshape(-1, 3)
        self.pose = np.eye(4) if pose is None else np.array(pose, dtype=np.float64)
        self.fixed = fixed
        self.neighbours = []

    # bound by ICP_Ceres / FrameSet below
    _engine = None
    _index = -1


class ICP_Ceres:
    """Drop-in for namespace ICP_Ceres (include/icp-ceres.h:22-45) over a list of Frames sharing one Engine."""

    def __init__(self, frames, device=0, flags=0):
        self.frames = frames
        self.engine = Engine(device, flags)
        self.engine.set_frames([f.pts for f in frames], None if any(f.nor is None for f in frames) else [f.nor for f in frames])
        for i, f in enumerate(frames):
            f._engine, f._index = self.engine, i

    def _push_poses(self):
        self.engine.set_poses([f.pose for f in self.frames], [1 if f.fixed else 0 for f in self.frames])

    def _pull_poses(self):
        P = self.engine.get_poses()
        for i, f in enumerate(self.frames):
            f.pose = P[i]

    def recomputeNormals(self, k=10):   # Frame::recomputeNormals, main_multiview.cpp:68
        nor, _ = self.engine.recompute_normals(k)
        for f, n in zip(self.frames, nor):
            f.nor = n

    def computePoseNeighbours(self, knn):   # main_multiview.cpp:104-117
        self._push_poses()
        edges = self.engine.pose_graph_knn(knn)
        for f in self.frames:
            f.neighbours = []
        for s, d in edges:
            self.frames[s].neighbours.append(OutgoingEdge(d))
        return edges

    def computeClosestPoints(self, cutoff, materialize=False):   # main_multiview.cpp:119-127
        self._push_poses()
        self.engine.correspond(cutoff)
        if materialize:
            self.pull_correspondances()

    def pull_correspondances(self):
        e = 0
        for f in self.frames:
            for ne in f.neighbours:
                if f.fixed:
                    ne.correspondances = []
                else:
                    a, b, d, w = self.engine.get_edge(e)
                    ne.correspondances = list(zip(a.tolist(), b.tolist(), d.tolist())); ne.weight = w
                e += 1

    def _opt(self, param, pointToPlane, robust, options=None):
        self.frames[0].fixed = True   # icp-ceres.cpp:242-244,342-344,417-419
        s = self.engine.optimize(param, COST_P2PLANE if pointToPlane else COST_P2P, robust, options)
        self._pull_poses()
        return s

    def ceresOptimizer(self, pointToPlane, robust, options=None):
        return self._opt(PARAM_QUAT, pointToPlane, robust, options)

    def ceresOptimizer_ceresAngleAxis(self, pointToPlane, robust, options=None):
        return self._opt(PARAM_AA, pointToPlane, robust, options)

    def ceresOptimizer_sophusSE3(self, pointToPlane, robust, automaticDiffLocalParam=True, options=None):
        return self._opt(PARAM_SE3, pointToPlane, robust, options)

    # ---- pairwise (include/icp-ceres.h:30-36) ------------------------------------------------------------
    @staticmethod
    def _pairwise(param, cost, src, dst, nor=None, device=0, options=None):
        src = _f64(src).reshape(-1, 3); dst = _f64(dst).reshape(-1, 3)
        nr = None if nor is None else _f64(nor).reshape(-1, 3)
        out = np.zeros(16); s = LmSummary(); cfg = Config(device, 0, None)
        check(_lib.lib().mvicp_pairwise(C.byref(cfg), C.c_int32(param), C.c_int32(cost), _p(src), _p(dst),
                                        _p(nr) if nr is not None else None, C.c_int64(len(src)),
                                        C.byref(options) if options is not None else None, _p(out), C.byref(s)))
        return out.reshape(4, 4).T.copy(), s.asdict()

    @staticmethod
    def closed_form(src, dst, nor=None, device=0):
        """ICP_Closedform::pointToPoint (nor is None) / pointToPlane (icp-closedform.cpp:9-54): 4x4 src -> dst."""
        src = _f64(src).reshape(-1, 3); dst = _f64(dst).reshape(-1, 3)
        nr = None if nor is None else _f64(nor).reshape(-1, 3)
        out = np.zeros(16); cfg = Config(device, 0, None)
        check(_lib.lib().mvicp_pairwise_closed(C.byref(cfg), C.c_int32(COST_P2P if nr is None else COST_P2PLANE), _p(src), _p(dst),
                                               _p(nr) if nr is not None else None, C.c_int64(len(src)), _p(out)))
        return out.reshape(4, 4).T.copy()

    @staticmethod
    def pointToPoint_EigenQuaternion(src, dst, **kw): return ICP_Ceres._pairwise(PARAM_QUAT, COST_P2P, src, dst, **kw)
    @staticmethod
    def pointToPoint_CeresAngleAxis(src, dst, **kw): return ICP_Ceres._pairwise(PARAM_AA, COST_P2P, src, dst, **kw)
    @staticmethod
    def pointToPoint_SophusSE3(src, dst, **kw): return ICP_Ceres._pairwise(PARAM_SE3, COST_P2P, src, dst, **kw)
    @staticmethod
    def pointToPlane_EigenQuaternion(src, dst, nor, **kw): return ICP_Ceres._pairwise(PARAM_QUAT, COST_P2PLANE, src, dst, nor, **kw)
    @staticmethod
    def pointToPlane_CeresAngleAxis(src, dst, nor, **kw): return ICP_Ceres._pairwise(PARAM_AA, COST_P2PLANE, src, dst, nor, **kw)
    @staticmethod
    def pointToPlane_SophusSE3(src, dst, nor, **kw): return ICP_Ceres._pairwise(PARAM_SE3, COST_P2PLANE, src, dst, nor, **kw)
