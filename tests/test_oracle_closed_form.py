"""Closed-form pairwise solvers restated in the oracle (icp-closedform.cpp:9-54) against numpy."""
import numpy as np

from helpers import scene


def _rot(w):
    th = np.linalg.norm(w); k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def test_closed_form_p2p_is_the_svd_solution(oracle):
    sc = scene(2, 2000, 13)
    src = sc["pts"][0]
    R = _rot(np.array([0.7, -1.1, 0.4])); t = np.array([0.01, -0.01, -0.005])
    rng = np.random.default_rng(0)
    dst = src @ R.T + t + rng.normal(0, 1e-4, src.shape)          # noisy: the answer is the least-squares fit, not (R, t)
    T = oracle.closed_form(src, dst)
    pb, qb = src.mean(0), dst.mean(0)
    U, S, Vt = np.linalg.svd((dst - qb).T @ (src - pb))
    Rn = U @ Vt
    assert np.linalg.det(Rn) > 0
    assert np.max(np.abs(T[:3, :3] - Rn)) < 1e-12 and np.max(np.abs(T[:3, 3] - (qb - Rn @ pb))) < 1e-12
    # exact correspondences: recovers the transform (README.md:141-150 "closed form" row)
    T = oracle.closed_form(src, src @ R.T + t)
    assert np.max(np.abs(T[:3, :3] - R)) < 1e-12 and np.max(np.abs(T[:3, 3] - t)) < 1e-12


def test_closed_form_p2p_reflection_branch_follows_the_reference(oracle):
    """det(U V^T) < 0 (planar, mirrored data): the reference negates the third COLUMN OF R (icp-closedform.cpp:20-22)."""
    rng = np.random.default_rng(3)
    src = rng.normal(size=(500, 3)) * np.array([1.0, 0.7, 0.4])
    dst = src * np.array([1.0, 1.0, -1.0])                        # mirror
    T = oracle.closed_form(src, dst)
    pb, qb = src.mean(0), dst.mean(0)
    U, S, Vt = np.linalg.svd((dst - qb).T @ (src - pb))
    Rn = U @ Vt
    assert np.linalg.det(Rn) < 0
    Rn[:, 2] *= -1
    assert np.max(np.abs(T[:3, :3] - Rn)) < 1e-12


def test_closed_form_p2plane_is_the_linearised_solution(oracle):
    sc = scene(2, 2000, 13)
    src, nor0 = sc["pts"][0], sc["nor"][0]
    R = _rot(np.array([0.004, -0.003, 0.002])); t = np.array([0.001, -0.002, 0.0015])
    dst = src @ R.T + t; nor = nor0 @ R.T
    T = oracle.closed_form(src, dst, nor)
    A = np.hstack([np.cross(src, nor), nor]); b = -np.sum((src - dst) * nor, axis=1)
    x = np.linalg.solve(A.T @ A, A.T @ b)
    cx, sx, cy, sy, cz, sz = np.cos(x[0]), np.sin(x[0]), np.cos(x[1]), np.sin(x[1]), np.cos(x[2]), np.sin(x[2])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    assert np.max(np.abs(T[:3, :3] - Rx @ Ry @ Rz)) < 1e-10 and np.max(np.abs(T[:3, 3] - x[3:])) < 1e-10
    # small motion: the linearisation is accurate to second order
    assert np.max(np.abs(T[:3, :3] - R)) < 1e-4 and np.max(np.abs(T[:3, 3] - t)) < 1e-4


def test_closed_form_p2p_rank_deficient_clouds(oracle):
    """Coplanar and collinear centred clouds (rank-2 / rank-1 cross-covariance): JacobiSVD still returns orthonormal U, V, so the
    reference yields a finite rotation (icp-closedform.cpp:18-22); the restatement must too -- and for a
    coplanar cloud moved rigidly it is the motion itself."""
    rng = np.random.default_rng(11)
    R = _rot(np.array([0.3, -0.5, 0.8])); t = np.array([0.2, -0.1, 0.05])
    plane = rng.normal(size=(300, 3)) * np.array([1.0, 0.6, 0.0])
    T = oracle.closed_form(plane, plane @ R.T + t)
    assert np.max(np.abs(T[:3, :3] - R)) < 1e-12 and np.max(np.abs(T[:3, 3] - t)) < 1e-12
    line = np.outer(rng.normal(size=200), np.array([0.3, -0.2, 0.9]))
    T = oracle.closed_form(line, line @ R.T + t)
    assert np.all(np.isfinite(T)) and np.max(np.abs(T[:3, :3] @ T[:3, :3].T - np.eye(3))) < 1e-12 and np.linalg.det(T[:3, :3]) > 0
    assert np.max(np.abs(line @ T[:3, :3].T + T[:3, 3] - (line @ R.T + t))) < 1e-12      # maps the line onto its image
    T = oracle.closed_form(plane[:2], plane[:2] @ R.T + t)                                   # n < 3
    assert np.all(np.isfinite(T))
