"""Normal estimation on the GPU vs the oracle restatement of Frame::recomputeNormals."""
import numpy as np
import pytest

from helpers import oracle_correspond, pose_rel_err, scene
from mv_lm_icp_b200 import COST_P2PLANE, PARAM_SE3, Engine, synth

pytestmark = pytest.mark.gpu


def test_normals_match_oracle_synthetic(oracle):
    sc = scene(4, 5000, 21)
    eng = Engine(); eng.set_frames(sc["pts"], sc["nor"])
    nor, ms = eng.recompute_normals(10)
    for f in range(4):
        ref = oracle.recompute_normals(sc["pts"][f], 10, threads=8)
        assert np.max(np.abs(nor[f] - ref)) < 1e-9          # jittered data: no distance ties, identical neighbour sets
        assert np.all(nor[f][:, 2] <= 0) and np.max(np.abs(np.linalg.norm(nor[f], axis=1) - 1)) < 1e-12
        # they are the surface normals up to sign and noise
        assert np.median(np.abs(np.sum(nor[f] * sc["nor"][f], axis=1))) > 0.99
    eng.close()


def test_normals_real_scan_and_lm_uses_them(oracle, golden_dir):
    """Real scan (fp64 records, quantised coordinates -> exact distance ties): compare where the 10th neighbour is unique;
    then the LM step must run on the recomputed (fp64) normals: same result as the oracle given those normals."""
    g = np.load(f"{golden_dir}/bunny_pair.npz")
    pts = [g["pts0"], g["pts1"]]
    eng = Engine(); eng.set_frames(pts, [g["nor0"], g["nor1"]])
    nor, _ = eng.recompute_normals(10)
    for f in range(2):
        ref, nn = oracle.recompute_normals(pts[f], 10, threads=8, want_nn=True)
        kd = oracle.KdIndex(pts[f], "kd")
        ok = np.ones(len(pts[f]), bool)
        for i in range(0, len(pts[f]), 7):      # sample: 11-NN to detect a tie at the 10th distance
            idx, d2 = oracle.knn(kd, pts[f][i], 11)
            ok[i] = d2[9] != d2[10]
        sel = np.arange(0, len(pts[f]), 7); sel = sel[ok[sel]]
        # ... and where the smallest eigenvalue is separated (collinear neighbourhoods at scan borders have two zero
        # eigenvalues: any vector of that plane is "the" normal, and fused vs unfused rounding picks different ones)
        P = pts[f][nn[sel]]; Cn = P - P.mean(1, keepdims=True)
        wv = np.linalg.eigvalsh(np.einsum("nki,nkj->nij", Cn, Cn))
        sel = sel[(wv[:, 1] - wv[:, 0]) > 1e-6 * wv[:, 2]]
        assert len(sel) > 1000
        assert np.max(np.abs(nor[f][sel] - ref[sel])) < 1e-7
    # LM with the new normals (rigidified poses so that only the normals differ from the other tests)
    poses = np.stack([np.eye(4), np.eye(4)])
    for i, P in enumerate([g["pose0"], g["pose1"]]):
        U, _, Vt = np.linalg.svd(P[:3, :3]); poses[i][:3, :3] = U @ Vt; poses[i][:3, 3] = P[:3, 3]
    eng.set_graph([(1, 0)]); eng.set_poses(poses); eng.correspond(0.05)
    f_, s_, d_, w_ = eng.get_edge(0)
    summ = eng.optimize(PARAM_SE3, COST_P2PLANE, True)
    P = eng.get_poses()
    Pref, sref, _ = oracle.optimize(pts, nor, poses, [(1, 0)], [(f_, s_)], [w_], param=PARAM_SE3, cost=COST_P2PLANE, robust=True, threads=8)
    assert summ["num_iterations"] == sref["num_iterations"]
    assert pose_rel_err(P, Pref) <= 1e-8
    eng.close()


def test_normals_full_size_timing():
    M, N = 20, 200_000
    sc = scene(M, N, 3)
    eng = Engine(); eng.set_frames(sc["pts"], sc["nor"])
    nor, ms = eng.recompute_normals(10)
    assert len(nor) == M and nor[0].shape == (N, 3) and np.isfinite(nor[5]).all()
    assert np.median(np.abs(np.sum(nor[3] * sc["nor"][3], axis=1))) > 0.99
    print("normals of 20 x 200k points: %.2f ms on device" % ms)
    eng.close()
