"""Closed-form pairwise solvers on the device (mvicp_pairwise_closed; icp-closedform.cpp:9-54) against the oracle's
restatement (itself pinned against numpy, tests/test_oracle_closed_form.py).  Written after the round's GPU budget was spent:
its logic is covered on the host model (tests/test_hostemu_engine.py); the file is named to run last."""
import numpy as np
import pytest

from helpers import scene
from mv_lm_icp_b200 import ICP_Ceres

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _rot(w):
    th = np.linalg.norm(w); k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


@pytest.mark.parametrize("n", [1, 3, 300, 20011, 200_000])
def test_point_to_point_matches_oracle(oracle, n):
    sc = scene(2, max(n, 4), 13)
    src = sc["pts"][0][:n]
    R = _rot(np.array([0.7, -1.1, 0.4])); t = np.array([0.01, -0.01, -0.005])
    dst = src @ R.T + t + np.random.default_rng(n).normal(0, 1e-4, src.shape)
    if n <= 3:
        T = ICP_Ceres.closed_form(src, dst)      # rank-deficient K (<= 3 points): the null direction of U V^T is a convention of the
        R3 = T[:3, :3]                           # SVD in the reference too -- but the result is a finite proper rotation, as JacobiSVD's
        assert np.all(np.isfinite(T)) and np.max(np.abs(R3 @ R3.T - np.eye(3))) < 1e-12 and np.linalg.det(R3) > 0
        return
    T, To = ICP_Ceres.closed_form(src, dst), oracle.closed_form(src, dst)
    assert np.max(np.abs(T - To)) < TOL
    if n >= 300:
        T = ICP_Ceres.closed_form(src, src @ R.T + t)             # exact correspondences: the transform itself
        assert np.max(np.abs(T[:3, :3] - R)) < 1e-11 and np.max(np.abs(T[:3, 3] - t)) < 1e-11


def test_rank_deficient_clouds(oracle):
    """Coplanar / collinear centred clouds: the cross-covariance has a vanishing singular value; the engine must
    return the finite rotation an SVD yields (for a rigidly moved coplanar cloud: the motion itself), equal to the oracle's."""
    rng = np.random.default_rng(11)
    R = _rot(np.array([0.3, -0.5, 0.8])); t = np.array([0.2, -0.1, 0.05])
    plane = rng.normal(size=(3000, 3)) * np.array([1.0, 0.6, 0.0])
    T = ICP_Ceres.closed_form(plane, plane @ R.T + t)
    assert np.max(np.abs(T[:3, :3] - R)) < 1e-11 and np.max(np.abs(T[:3, 3] - t)) < 1e-11
    assert np.max(np.abs(T - oracle.closed_form(plane, plane @ R.T + t))) < 1e-11
    line = np.outer(rng.normal(size=2000), np.array([0.3, -0.2, 0.9]))
    T = ICP_Ceres.closed_form(line, line @ R.T + t)
    assert np.all(np.isfinite(T)) and np.max(np.abs(T[:3, :3] @ T[:3, :3].T - np.eye(3))) < 1e-12 and np.linalg.det(T[:3, :3]) > 0
    assert np.max(np.abs(line @ T[:3, :3].T + T[:3, 3] - (line @ R.T + t))) < 1e-11


def test_reflection_branch_follows_the_reference(oracle):
    rng = np.random.default_rng(3)
    src = rng.normal(size=(5000, 3)) * np.array([1.0, 0.7, 0.4])
    dst = src * np.array([1.0, 1.0, -1.0])
    T, To = ICP_Ceres.closed_form(src, dst), oracle.closed_form(src, dst)
    assert np.max(np.abs(T - To)) < TOL and np.linalg.det(T[:3, :3]) > 0


@pytest.mark.parametrize("n", [300, 20011, 200_000])
def test_point_to_plane_matches_oracle(oracle, n):
    sc = scene(2, n, 13)
    src, nor0 = sc["pts"][0], sc["nor"][0]
    R = _rot(np.array([0.004, -0.003, 0.002])); t = np.array([0.001, -0.002, 0.0015])
    dst = src @ R.T + t; nor = nor0 @ R.T
    T, To = ICP_Ceres.closed_form(src, dst, nor), oracle.closed_form(src, dst, nor)
    assert np.max(np.abs(T - To)) < 1e-10          # 6x6 system with condition ~1e6: summation order shows at 1e-12
    assert np.max(np.abs(T[:3, :3] - R)) < 1e-4 and np.max(np.abs(T[:3, 3] - t)) < 1e-4


def test_arguments():
    from mv_lm_icp_b200._lib import MvicpError
    import ctypes as C
    from mv_lm_icp_b200 import _lib
    a = np.zeros((4, 3)); out = np.zeros(16); cfg = _lib.Config(0, 0, None)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    assert _lib.lib().mvicp_pairwise_closed(C.byref(cfg), C.c_int32(1), p(a), p(a), None, C.c_int64(4), p(out)) != 0   # p2plane without normals
    assert _lib.lib().mvicp_pairwise_closed(C.byref(cfg), C.c_int32(2), p(a), p(a), p(a), C.c_int64(4), p(out)) != 0   # MIXED has no closed form
    assert _lib.lib().mvicp_pairwise_closed(C.byref(cfg), C.c_int32(0), p(a), p(a), None, C.c_int64(0), p(out)) != 0


def test_pairwise_driver_prints_the_closed_form_row(tmp_path):
    """apps/pairwise_b200: the "closed form" row of the reference binary's accuracy table (main_pairwise.cpp:121)."""
    import re
    import subprocess
    import test_app_multiview as A
    A._build()
    sc = scene(2, 3000, 9)
    A._write_cloud(tmp_path / "c.xyz", sc["pts"][0], sc["nor"][0])
    r = subprocess.run([A.BIN_PAIR, f"--cloud={tmp_path}/c.xyz"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    m = re.search(r"closed form\s+diff_tra:([0-9.e+-]+)\t diff_rot_degrees:([0-9.e+-]+)", r.stdout)
    assert m and float(m.group(1)) < 1e-12 and float(m.group(2)) < 1e-5 and "TIMING[closed]" in r.stdout
