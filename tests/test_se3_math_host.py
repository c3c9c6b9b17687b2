"""The PRODUCT's rigid-motion arithmetic (csrc/se3_math.cuh: the header lm_init_kernel / lm_step_kernel / lm_edge_kernel use
on the device) compiled for the host (tools/se3_host.cpp) against the oracle's independent restatement and finite
differences: pose <-> parameter conversions, the three Plus operators, the tangent maps and the general frame model."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AA, QUAT, SE3 = 0, 1, 2


@pytest.fixture(scope="module")
def H(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("se3host") / "libse3host.so")
    r = subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                        os.path.join(ROOT, "tools", "se3_host.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(so)

    def call(name, outs, *ins):
        args = []
        for a in ins:
            if isinstance(a, int):
                args.append(C.c_int(a))
            else:
                a = np.ascontiguousarray(a, dtype=np.float64).ravel(); args.append(a.ctypes.data_as(C.POINTER(C.c_double))); ins_keep.append(a)
        res = [np.zeros(n) for n in outs]
        getattr(lib, name)(*args, *[r_.ctypes.data_as(C.POINTER(C.c_double)) for r_ in res])
        return res if len(res) > 1 else res[0]
    ins_keep = []
    return call


def _poses(rng, n, nonrigid=False):
    out = []
    for _ in range(n):
        w = rng.normal(0, 1.0, 3); th = np.linalg.norm(w); k = w / th
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
        if nonrigid:
            R = R @ np.diag(1 + rng.normal(0, 2e-3, 3))
        P = np.eye(4); P[:3, :3] = R; P[:3, 3] = rng.normal(0, 0.3, 3)
        out.append(P)
    return out


def _close(a, b, tol=4e-15):
    """A few ulps: the product and the oracle implement the same formulas with different operation orders."""
    a = np.asarray(a, float); b = np.asarray(b, float)
    return a.shape == b.shape and np.max(np.abs(a - b)) <= tol * max(1.0, np.max(np.abs(b)))


def _col(P):          # 4x4 -> 16 doubles column-major (Isometry3d::data())
    return np.asarray(P).T.reshape(16)


@pytest.mark.parametrize("param", [AA, QUAT, SE3])
def test_pose_parameter_conversions_match_oracle(H, oracle, param):
    rng = np.random.default_rng(param)
    for P in _poses(rng, 20) + _poses(rng, 10, nonrigid=True):
        x = H("h_param_of_pose", [7], param, _col(P))
        xo = oracle.pose_to_param(P, param)
        n = 6 if param == AA else 7
        assert np.array_equal(x[:n], np.asarray(xo)[:n])                 # same formulas, same operation order: bit-identical
        P2 = H("h_pose_of_param", [16], param, x).reshape(4, 4).T
        assert np.array_equal(P2, oracle.param_to_pose(xo, param))


def test_plus_operators_match_oracle(H, oracle):
    rng = np.random.default_rng(5)
    for P in _poses(rng, 20):
        for scale in (1e-9, 1e-3, 0.5):
            d = rng.normal(0, scale, 6)
            x = np.asarray(oracle.pose_to_param(P, SE3))
            assert _close(H("h_param_plus", [7], SE3, x, d), oracle.se3_plus(x, d))
            xq = np.asarray(oracle.pose_to_param(P, QUAT))
            got = H("h_param_plus", [7], QUAT, xq, d)
            assert _close(got[:4], oracle.quat_plus(xq[:4], d[:3])) and _close(got[4:], xq[4:7] + d[3:])
            xa = np.asarray(oracle.pose_to_param(P, AA))
            assert _close(H("h_param_plus", [7], AA, xa, d)[:6], xa[:6] + d)
    z = np.zeros(6); xq = np.asarray(oracle.pose_to_param(_poses(rng, 1)[0], QUAT))
    assert _close(H("h_param_plus", [7], QUAT, xq, z)[:7], xq[:7])      # |delta| = 0 branch (eigen_quaternion.h:95-100)


def test_group_and_rotation_primitives_match_oracle(H, oracle):
    rng = np.random.default_rng(9)
    for _ in range(30):
        tg = rng.normal(0, 1, 6) * rng.choice([1e-10, 1e-3, 1.0])
        assert _close(H("h_se3_exp", [7], tg), oracle.se3_exp(tg))
        a, b = oracle.se3_exp(rng.normal(0, 1, 6)), oracle.se3_exp(rng.normal(0, 1, 6))
        assert _close(H("h_se3_compose", [7], a, b), oracle.se3_mul(a, b))
        R = _poses(rng, 1)[0][:3, :3]
        assert _close(H("h_quat_of_matrix", [4], R.reshape(9)), oracle.quat_from_matrix(R))
        q = oracle.quat_from_matrix(R) * (1 + rng.normal(0, 1e-3))                 # not normalised on purpose
        assert _close(H("h_matrix_of_quat", [9], q).reshape(3, 3), oracle.quat_to_matrix(q))
        v = rng.normal(size=3)
        assert _close(H("h_quat_rotate", [3], q, v), oracle.quat_transform(q, v))
        assert _close(H("h_aa_of_matrix", [3], R.reshape(9)), oracle.rotmat_to_angle_axis(R))
        aa = rng.normal(0, 1, 3) * rng.choice([1e-9, 1e-2, 1.0])
        assert _close(H("h_matrix_of_aa", [9], aa).reshape(3, 3), oracle.angle_axis_to_rotmat(aa))
        Rf = H("h_rotation_of_aa_functor", [9], aa).reshape(3, 3)                  # what AngleAxisRotatePoint applies
        assert np.max(np.abs(Rf @ v - oracle.angle_axis_rotate(aa, v))) < 1e-15


@pytest.mark.parametrize("param", [AA, QUAT, SE3])
def test_tangent_map_is_the_derivative_of_plus(H, oracle, param):
    """T(x (+) delta) = T(x) exp(K delta) to first order: K maps the parameterisation's tangent to the body tangent."""
    rng = np.random.default_rng(20 + param)
    for P in _poses(rng, 8):
        x = H("h_param_of_pose", [7], param, _col(P))
        K = H("h_tangent_map", [36], param, x).reshape(6, 6)
        R0, t0 = H("h_Rt_of_param", [9, 3], param, x); R0 = R0.reshape(3, 3)
        h = 1e-6
        for j in range(6):
            d = np.zeros(6); d[j] = h
            R1, t1 = H("h_Rt_of_param", [9, 3], param, H("h_param_plus", [7], param, x, d)); R1 = R1.reshape(3, 3)
            d[j] = -h
            R2, t2 = H("h_Rt_of_param", [9, 3], param, H("h_param_plus", [7], param, x, d)); R2 = R2.reshape(3, 3)
            W = R0.T @ (R1 - R2) / (2 * h)                                         # [omega]x
            omega = np.array([W[2, 1], W[0, 2], W[1, 0]]); ups = R0.T @ (t1 - t2) / (2 * h)
            assert np.max(np.abs(np.concatenate([ups, omega]) - K[:, j])) < 1e-8, (param, j)


@pytest.mark.parametrize("param", [QUAT, SE3])
def test_general_frame_model_is_ceres_chain_rule(H, oracle, param):
    """Non-unit quaternion: y(v) = F v + t and dy/d delta_j = D_j v + c_j must equal what Ceres computes: the functor's
    global Jacobian dy/dx AT x (not renormalised) times the local parameterisation's Jacobian dPlus/d delta at 0 --
    for SE3 the autodiff of x * exp(delta) including its renormalisation (sophus_se3.h:64-68), for the quaternion
    eigen_quaternion.h:108-114."""
    rng = np.random.default_rng(40 + param)
    tested = 0
    for P in _poses(rng, 10, nonrigid=True):
        x = H("h_param_of_pose", [7], param, _col(P))
        if abs(np.linalg.norm(x[:4]) - 1) < 1e-4:                                   # want the really non-unit case
            continue
        tested += 1
        F, t, D, c = H("h_frame_general", [9, 3, 54, 18], param, x)
        F = F.reshape(3, 3); D = D.reshape(6, 3, 3); c = c.reshape(6, 3)
        v = rng.normal(size=3)
        y = lambda z: oracle.quat_transform(z[:4], v) + z[4:7]                       # what the functors apply
        assert np.max(np.abs(F @ v + t - y(x))) < 1e-14
        h = 1e-6
        Jg = np.zeros((3, 7))
        for k in range(7):
            e = np.zeros(7); e[k] = h
            Jg[:, k] = (y(x + e) - y(x - e)) / (2 * h)                               # polynomial in x: central differences are exact to rounding
        if param == SE3:
            Pl = oracle.se3_plus_jacobian_autodiff(x)
        else:
            Pl = np.zeros((7, 6))
            for j in range(3):
                d = np.zeros(3); d[j] = h
                Pl[:4, j] = (oracle.quat_plus(x[:4], d) - oracle.quat_plus(x[:4], -d)) / (2 * h)
            Pl[4:, 3:] = np.eye(3)
        want = Jg @ Pl
        for j in range(6):
            assert np.max(np.abs(want[:, j] - (D[j] @ v + c[j]))) < 2e-9, (param, j)
    assert tested >= 3
