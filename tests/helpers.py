"""Shared scene / oracle plumbing for the parity tests."""
import functools

import numpy as np

from mv_lm_icp_b200 import synth


@functools.lru_cache(maxsize=8)
def scene(n_views, n_points, config_id):
    return synth.make_scene(n_views, n_points, config_id=config_id)


def oracle_correspond(O, pts, poses, edges, thresh=0.05, kind="kd", threads=8, fixed0=True):
    """Restated ApproachComponents::computeClosestPoints: per edge (first, second, dist, weight, nn_idx, nn_d2)."""
    idxs = {}
    out = []
    for s, d in edges:
        if fixed0 and s == 0:
            out.append(None)
            continue
        if d not in idxs:
            idxs[d] = O.KdIndex(pts[d], kind)
        i, d2 = idxs[d].closest_points(pts[s], poses[s], poses[d], threads=threads)
        f, sec, dist, w, med = O.filter_edge(i, d2, np.float32(thresh))
        out.append(dict(first=f, second=sec, dist=dist, weight=w, nn_idx=i, nn_d2=d2, median=med))
    return out


def pose_rel_err(A, B):
    """max over frames of |A - B|_max / max(1, |B|_max): relative difference of the 3x4 pose parameters."""
    A = np.asarray(A)[:, :3, :]; B = np.asarray(B)[:, :3, :]
    return float(np.max(np.abs(A - B)) / max(1.0, np.max(np.abs(B))))


def rot_err_deg(A, B):
    R = A[:3, :3] @ B[:3, :3].T
    return float(np.degrees(np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1))))


def host_threads():
    """Threads for the oracle in the full-size tests: affinity, cgroup quota and physical cores (bench.py:cpu_threads)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("_bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m.cpu_threads()


def oracle_round(O, pts, nor, poses, edges, param, cost, robust=True, threads=8, kind="kd", index_cache=None):
    """One outer round on the CPU oracle from `poses`: (new poses, summary, per-edge reference dicts)."""
    idxs = index_cache if index_cache is not None else {}
    ref = []
    for s, d in edges:
        if s == 0:
            ref.append(None); continue
        if d not in idxs:
            idxs[d] = O.KdIndex(pts[d], kind)
        i, d2 = idxs[d].closest_points(pts[s], poses[s], poses[d], threads=threads)
        f, sec, dist, w, med = O.filter_edge(i, d2, np.float32(0.05))
        ref.append(dict(first=f, second=sec, dist=dist, weight=w, nn_idx=i, nn_d2=d2, median=med))
    corr = [((r["first"], r["second"]) if r else (np.zeros(0, np.int32), np.zeros(0, np.int32))) for r in ref]
    w = [np.float32(r["weight"]) if r else np.float32(0) for r in ref]
    P, summ, _ = O.optimize(pts, nor, poses, edges, corr, w, param=param, cost=cost, robust=robust, se3_autodiff=True, threads=threads)
    return P, summ, ref
