"""Pins the oracle's SE(3) / quaternion / angle-axis arithmetic (oracle/geom.h) against independent references:
scipy (matrix exponential, Rotation), numpy, finite differences, and the group elements / tangents of the vendored
Sophus tests (ext/sophus-ceres/test/core/test_se3.cpp:40-82, tolerance 1e-10 as in tests.hpp)."""
import numpy as np
import pytest
from scipy.linalg import expm
from scipy.spatial.transform import Rotation

SMALL = 1e-10


def _hat6(t):
    u, w = t[:3], t[3:]
    H = np.zeros((4, 4))
    H[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
    H[:3, 3] = u
    return H


def _mat(x7, O):
    T = np.eye(4); T[:3, :3] = O.quat_to_matrix(x7[:4]); T[:3, 3] = x7[4:]
    return T


def test_se3_exp_matches_matrix_exponential(oracle, golden_dir):
    g = np.load(f"{golden_dir}/sophus_vectors.npz")
    for t in g["tangents"]:
        x = oracle.se3_exp(t)
        assert abs(np.linalg.norm(x[:4]) - 1) < 1e-15
        assert np.max(np.abs(_mat(x, oracle) - expm(_hat6(t)))) < SMALL * max(1, np.abs(t).max())
    # tiny-angle branch (theta < 1e-10, se3.hpp:478-480)
    t = np.array([1.0, -2.0, 0.5, 1e-12, -2e-12, 3e-12])
    assert np.max(np.abs(_mat(oracle.se3_exp(t), oracle) - expm(_hat6(t)))) < 1e-11


def test_se3_group_product_and_action(oracle, golden_dir):
    g = np.load(f"{golden_dir}/sophus_vectors.npz")
    els = []
    for w, t in zip(g["so3_omega"], g["trans"]):
        x = oracle.se3_exp(np.concatenate([np.zeros(3), w])); x[4:] = t
        els.append(x)
    pts = [np.array([1.0, 2, 4]), np.array([1.0, -3, 0.5])]   # test_se3.cpp:84-86
    for a in els:
        Ta = _mat(a, oracle)
        R = Rotation.from_quat(a[:4]).as_matrix()
        assert np.max(np.abs(Ta[:3, :3] - R)) < SMALL
        for p in pts:
            assert np.max(np.abs(oracle.quat_transform(a[:4], p) + a[4:] - (Ta @ np.append(p, 1))[:3])) < SMALL * 100
        for b in els:
            c = oracle.se3_mul(a, b)
            assert abs(np.linalg.norm(c[:4]) - 1) < 1e-15            # operator*= renormalises (se3.hpp:317-321)
            assert np.max(np.abs(_mat(c, oracle) - Ta @ _mat(b, oracle))) < SMALL * 1e3


def test_se3_plus_jacobians(oracle):
    rng = np.random.default_rng(0)
    for _ in range(5):
        x = oracle.se3_exp(rng.normal(size=6))
        Ja = oracle.se3_internal_jacobian(x)            # analytic LocalParameterizationSE3 (sophus_se3.h:45-51)
        Jd = oracle.se3_plus_jacobian_autodiff(x)       # AutoDiffLocalParameterization<SophusSE3Plus> (sophus_se3.h:64-68)
        assert np.max(np.abs(Ja - Jd)) < 1e-12          # equal on unit quaternions
        h = 1e-6
        for c in range(6):
            d = np.zeros(6); d[c] = h
            fd = (oracle.se3_plus(x, d) - oracle.se3_plus(x, -d)) / (2 * h)
            assert np.max(np.abs(fd - Jd[:, c])) < 1e-8
    # non-unit quaternion: autodiff differentiates through the normalisation, the analytic one does not
    x = oracle.se3_exp(rng.normal(size=6)); x[:4] *= 1.01
    assert np.max(np.abs(oracle.se3_internal_jacobian(x) - oracle.se3_plus_jacobian_autodiff(x))) > 1e-4


def test_quaternion_conversions_and_plus(oracle):
    rng = np.random.default_rng(1)
    mats = [Rotation.from_rotvec(rng.normal(size=3) * s).as_matrix() for s in (0.01, 1.0, 3.0) for _ in range(4)]
    mats += [Rotation.from_rotvec([np.pi - 1e-3, 0, 0]).as_matrix(), Rotation.from_rotvec([0, np.pi - 1e-3, 0]).as_matrix(),
             Rotation.from_rotvec([0, 0, np.pi - 1e-3]).as_matrix()]      # negative-trace branches of Eigen's formula
    for R in mats:
        q = oracle.quat_from_matrix(R)
        qs = Rotation.from_matrix(R).as_quat()
        assert min(np.max(np.abs(q - qs)), np.max(np.abs(q + qs))) < 1e-12
        assert np.max(np.abs(oracle.quat_to_matrix(q) - R)) < 1e-12
        v = rng.normal(size=3)
        assert np.max(np.abs(oracle.quat_transform(q, v) - R @ v)) < 1e-12
        aa = oracle.rotmat_to_angle_axis(R)
        assert np.max(np.abs(aa - Rotation.from_matrix(R).as_rotvec())) < 1e-9
        assert np.max(np.abs(oracle.angle_axis_to_rotmat(aa) - R)) < 1e-12
        assert np.max(np.abs(oracle.angle_axis_rotate(aa, v) - R @ v)) < 1e-12
    # EigenQuaternionParameterization::Plus rotates by 2|delta| about delta, in world axes (eigen_quaternion.h:89-106)
    q = oracle.quat_from_matrix(mats[5]); d = np.array([0.01, -0.02, 0.03])
    qp = oracle.quat_plus(q, d)
    Rexp = Rotation.from_rotvec(2 * d).as_matrix() @ mats[5]
    assert np.max(np.abs(oracle.quat_to_matrix(qp) - Rexp)) < 1e-12
    assert np.array_equal(oracle.quat_plus(q, np.zeros(3)), q)
    # tiny-angle branches of Ceres' angle-axis routines
    tiny = np.array([1e-9, -2e-9, 5e-10])
    assert np.max(np.abs(oracle.angle_axis_to_rotmat(tiny) - (np.eye(3) + _hat6(np.concatenate([np.zeros(3), tiny]))[:3, :3]))) == 0.0


def test_general_inverse_and_edge_transform(oracle, golden_dir):
    g = np.load(f"{golden_dir}/bunny_pair.npz")
    for P in (g["pose0"], g["pose1"]):
        A = P[:3, :3]
        assert np.max(np.abs(A @ A.T - np.eye(3))) > 1e-3          # the sample poses are not rigid
        assert np.max(np.abs(oracle.mat3_inverse(A) - np.linalg.inv(A))) < 1e-12
    q = oracle.edge_queries(g["pts1"][:50], g["pose1"], g["pose0"])
    want = (np.linalg.inv(g["pose0"][:3, :3]) @ ((g["pts1"][:50] @ g["pose1"][:3, :3].T + g["pose1"][:3, 3]) - g["pose0"][:3, 3]).T).T
    assert np.max(np.abs(q - want)) < 1e-13


@pytest.mark.parametrize("param", [0, 1, 2])
def test_pose_param_round_trip(oracle, param):
    rng = np.random.default_rng(2)
    for _ in range(5):
        P = np.eye(4); P[:3, :3] = Rotation.from_rotvec(rng.normal(size=3)).as_matrix(); P[:3, 3] = rng.normal(size=3)
        x = oracle.pose_to_param(P, param)
        assert np.max(np.abs(oracle.param_to_pose(x, param) - P)) < 1e-12
