"""Pins the oracle's LM restatement (oracle/lm.h + functors): finite differences, the pairwise known-answer of the
reference's README, scipy as an independent optimiser, and the committed golden trace.  (Never against Ceres itself:
it is not installable here -- "parity unpinned", see oracle_icp.cpp header.)"""
import numpy as np
import pytest
from scipy.optimize import least_squares
from scipy.spatial.transform import Rotation

from helpers import oracle_correspond, scene
from mv_lm_icp_b200 import synth


def _problem(O, n=1500, cfg=41):
    sc = scene(3, n, cfg)
    edges = synth.ring_edges(3, 2)
    ref = oracle_correspond(O, sc["pts"], sc["poses_init"], edges, threads=4)
    corr = [((r["first"], r["second"]) if r else (np.zeros(0, np.int32), np.zeros(0, np.int32))) for r in ref]
    w = [np.float32(r["weight"]) if r else np.float32(0) for r in ref]
    return sc, edges, corr, w


@pytest.mark.parametrize("param", [0, 1, 2])
@pytest.mark.parametrize("cost", [0, 1, 2])
@pytest.mark.parametrize("robust", [False, True])
def test_gradient_matches_finite_differences(oracle, param, cost, robust):
    """g = J^T r (after the robust correction and the local parameterisation) must be the derivative of the cost along
    Plus(x, delta): checks functors, Jets, SoftL1 and the three Plus/Jacobian pairs together."""
    sc, edges, corr, w = _problem(oracle)
    poses = sc["poses_init"]
    c0, H, g = oracle.evaluate(sc["pts"], sc["nor"], poses, edges, corr, w, param=param, cost=cost, robust=robust, threads=4)
    assert np.allclose(H, H.T, rtol=0, atol=1e-9 * np.abs(H).max())
    x = [oracle.pose_to_param(P, param) for P in poses]
    h = 1e-6
    for col in range(12):
        f, l = 1 + col // 6, col % 6
        d = np.zeros(6); d[l] = h
        cs = []
        for sgn in (+1, -1):
            xf = x[f].copy()
            if param == 0:
                xf = xf + sgn * d
            elif param == 1:
                xf = np.concatenate([oracle.quat_plus(xf[:4], sgn * d[:3]), xf[4:] + sgn * d[3:]])
            else:
                xf = oracle.se3_plus(xf, sgn * d)
            pp = [p.copy() for p in poses]; pp[f] = oracle.param_to_pose(xf, param)
            cs.append(oracle.evaluate(sc["pts"], sc["nor"], pp, edges, corr, w, param=param, cost=cost, robust=robust, threads=4, want_jac=False)[0])
        fd = (cs[0] - cs[1]) / (2 * h)
        assert abs(fd - g[col]) <= 2e-6 * max(1.0, np.abs(g).max()), (col, fd, g[col])


@pytest.mark.parametrize("param", [0, 1, 2])
@pytest.mark.parametrize("cost", [0, 1])
def test_pairwise_known_answer(oracle, golden_dir, param, cost):
    """main_pairwise.cpp:29-134 / README.md:141-150: with exact correspondences every Ceres solver recovers P to
    diff_tra ~1e-10."""
    g = np.load(f"{golden_dir}/bunny_pair.npz")
    src, nor = g["pts0"], g["nor0"]
    P = np.eye(4)
    P[:3, :3] = (Rotation.from_euler("x", np.pi / 4) * Rotation.from_euler("y", 1.0) * Rotation.from_euler("z", -0.2)).as_matrix()
    P[:3, 3] = [0.01, -0.01, -0.005]
    dst = src @ P[:3, :3].T + P[:3, 3]; ndst = nor @ P[:3, :3].T
    for autodiff in ((False, True) if param == 2 else (False,)):
        Pe, s = oracle.pairwise(src, dst, ndst, param=param, cost=cost, se3_autodiff=autodiff, threads=4)
        assert np.linalg.norm(Pe[:3, 3] - P[:3, 3]) < 2e-9
        assert np.degrees(np.arccos(np.clip((np.trace(Pe[:3, :3] @ P[:3, :3].T) - 1) / 2, -1, 1))) < 1e-5
        assert s["termination"] in (0, 1, 2) and s["num_iterations"] <= 10


def test_lm_reaches_the_minimum_scipy_finds(oracle):
    """Independent optimiser on the same robust point-to-plane objective (own numpy residuals, SE3 left out of it:
    scipy optimises rotation-vector + translation per free frame)."""
    sc, edges, corr, w = _problem(oracle, 800, 42)
    o = oracle.default_options(); o.function_tolerance = 1e-15; o.parameter_tolerance = 1e-14; o.max_num_iterations = 200
    Pfin, s, _ = oracle.optimize(sc["pts"], sc["nor"], sc["poses_init"], edges, corr, w, param=2, cost=1, robust=True, options=o, threads=4)

    def unpack(z):
        Ps = [sc["poses_init"][0]]
        for f in range(2):
            P = np.eye(4); P[:3, :3] = Rotation.from_rotvec(z[6 * f:6 * f + 3]).as_matrix(); P[:3, 3] = z[6 * f + 3:6 * f + 6]
            Ps.append(P)
        return Ps

    def resid(z):
        Ps = unpack(z); out = []
        for e, (s_, k_) in enumerate(edges):
            if s_ == 0:
                continue
            f, sec = corr[e]
            ys = sc["pts"][s_][f] @ Ps[s_][:3, :3].T + Ps[s_][:3, 3]
            yk = sc["pts"][k_][sec] @ Ps[k_][:3, :3].T + Ps[k_][:3, 3]
            n2 = sc["nor"][k_][sec] @ Ps[k_][:3, :3].T
            r = np.einsum("ij,ij->i", ys - yk, n2)
            b = float(w[e]) ** 2
            out.append(np.sqrt(2 * b * (np.sqrt(1 + r * r / b) - 1)))     # sqrt(rho): 1/2 sum of squares = the cost
        return np.concatenate(out)

    z0 = np.concatenate([np.concatenate([Rotation.from_matrix(P[:3, :3]).as_rotvec(), P[:3, 3]]) for P in sc["poses_init"][1:]])
    sol = least_squares(resid, z0, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=2000)
    assert abs(0.5 * np.sum(sol.fun ** 2) - s["final_cost"]) <= 1e-9 * s["final_cost"]
    assert np.max(np.abs(np.stack(unpack(sol.x)) - Pfin)) < 1e-6


def test_golden_trace_is_reproduced(oracle, golden_dir):
    g = np.load(f"{golden_dir}/lm_golden.npz")
    sc = scene(4, 3000, 7)
    edges = synth.ring_edges(4, 2)
    ref = oracle_correspond(oracle, sc["pts"], sc["poses_init"], edges, threads=4)
    corr = [((r["first"], r["second"]) if r else (np.zeros(0, np.int32), np.zeros(0, np.int32))) for r in ref]
    w = [np.float32(r["weight"]) if r else np.float32(0) for r in ref]
    for param, cost, robust in [(0, 0, 0), (1, 1, 1), (2, 1, 1), (2, 2, 1), (1, 0, 1)]:
        P, s, tr = oracle.optimize(sc["pts"], sc["nor"], sc["poses_init"], edges, corr, w, param=param, cost=cost, robust=bool(robust), threads=1)
        k = f"p{param}_c{cost}_r{robust}"
        assert np.max(np.abs(P - g[k + "_poses"])) < 1e-12
        assert [s["termination"], s["num_iterations"], s["num_successful_steps"]] == g[k + "_summary"].tolist()
        assert np.allclose(tr, g[k + "_trace"], rtol=1e-9, atol=0)
