"""Regenerates tests/golden/*.npz.  Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

bunny_pair.npz   config 1 of BASELINE.json: the reference's own sample scans samples/Bunny_RealData/cloudXYZ_{0,1}.xyz
                 (x y z nx ny nz per line; the loader's trailing garbage point, common.h:224-239, is NOT reproduced)
                 with poses_{0,1}.txt (4x4 row-major).  Stored as float64 exactly as parsed.
                 + nn_idx / nn_d2: output of the REFERENCE's nanoflann (oracle/_ref/libref_nanoflann.so, compiled from
                 /root/reference/include/nanoflann.hpp) for edge 1 -> 0 under those poses, i.e. the pinned answer of
                 Frame::computeClosestPointsToNeighbours' inner loop (frame.cpp:129-138);
                 + first/second/dist/weight of frame.cpp:140-176 from those.
dino_pair.npz    the reference's second sample set, samples/dinosaur/cloud_{1,2}.xyz with pose_{1,2}.txt: millimetre units
                 (|coordinates| up to 686, NN distances of tens of mm), 5-digit pose matrices (not exactly orthonormal):
                 edge 2 -> 1, reference nanoflann answer + frame.cpp:140-176 at cutoff 25 (the default 0.05 m is meaningless
                 in mm); pins the fp32 screening bound at a different coordinate magnitude.
lm_golden.npz    oracle LM outputs (final poses, iteration trace) on a small seeded synthetic scene for every
                 parameterisation x cost; pins the oracle against accidental change (NOT against Ceres: unpinned).
sophus_vectors.npz  the SE3 group elements / tangents of ext/sophus-ceres/test/core/test_se3.cpp:40-82 (values only).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from mv_lm_icp_b200 import synth  # noqa: E402

REF = "/root/reference/samples/Bunny_RealData"
OUT = os.path.dirname(os.path.abspath(__file__))


def bunny_pair():
    c = [np.loadtxt(f"{REF}/cloudXYZ_{i}.xyz") for i in (0, 1)]
    P = [np.loadtxt(f"{REF}/poses_{i}.txt") for i in (0, 1)]
    ref = O.KdIndex(c[0][:, :3], "ref")
    idx, d2 = ref.closest_points(c[1][:, :3], P[1], P[0])
    first, second, dist, w, med = O.filter_edge(idx, d2, np.float32(0.05))
    np.savez_compressed(os.path.join(OUT, "bunny_pair.npz"), pts0=c[0][:, :3], nor0=c[0][:, 3:6], pts1=c[1][:, :3],
                        nor1=c[1][:, 3:6], pose0=P[0], pose1=P[1], nn_idx=idx, nn_d2=d2, first=first, second=second,
                        dist=dist, weight=w, median=med)
    print("bunny_pair:", c[0].shape, c[1].shape, "inliers", len(first), "weight", w)


def dino_pair():
    D = "/root/reference/samples/dinosaur"
    c = [np.loadtxt(f"{D}/cloud_{i}.xyz") for i in (1, 2)]
    P = [np.loadtxt(f"{D}/pose_{i}.txt") for i in (1, 2)]
    ref = O.KdIndex(c[0][:, :3], "ref")
    idx, d2 = ref.closest_points(c[1][:, :3], P[1], P[0])
    first, second, dist, w, med = O.filter_edge(idx, d2, np.float32(25.0))
    np.savez_compressed(os.path.join(OUT, "dino_pair.npz"), pts0=c[0][:, :3], nor0=c[0][:, 3:6], pts1=c[1][:, :3], nor1=c[1][:, 3:6],
                        pose0=P[0], pose1=P[1], nn_idx=idx, nn_d2=d2, first=first, second=second, dist=dist, weight=w, median=med)
    print("dino_pair:", c[0].shape, c[1].shape, "inliers", len(first), "weight", w)


def lm_golden():
    sc = synth.make_scene(4, 3000, config_id=7)
    edges = synth.ring_edges(4, 2)
    idxs = [O.KdIndex(p, "kd") for p in sc["pts"]]
    corr, weights = [], []
    for s, d in edges:
        i, d2 = idxs[d].closest_points(sc["pts"][s], sc["poses_init"][s], sc["poses_init"][d])
        f, sec, dist, w, _ = O.filter_edge(i, d2, np.float32(0.05))
        corr.append((f, sec)); weights.append(w)
    out = {}
    for param in (0, 1, 2):
        for cost in (0, 1, 2):
            for robust in (0, 1):
                P, s, tr = O.optimize(sc["pts"], sc["nor"], sc["poses_init"], edges, corr, weights, param=param, cost=cost,
                                      robust=bool(robust), se3_autodiff=True, threads=1)
                k = f"p{param}_c{cost}_r{robust}"
                out[k + "_poses"] = P; out[k + "_trace"] = tr
                out[k + "_summary"] = np.array([s["termination"], s["num_iterations"], s["num_successful_steps"]], np.int64)
                out[k + "_cost"] = np.array([s["initial_cost"], s["final_cost"]])
                print(k, O.TERMINATION[s["termination"]], s["num_iterations"], s["initial_cost"], s["final_cost"])
    np.savez_compressed(os.path.join(OUT, "lm_golden.npz"), **out)


def sophus_vectors():
    # ext/sophus-ceres/test/core/test_se3.cpp:40-65: SE3(SO3::exp(w), t) ; :67-82 tangents (upsilon, omega)
    pi = np.pi
    el = [([0.2, 0.5, 0.0], [0, 0, 0]), ([0.2, 0.5, -1.0], [10, 0, 0]), ([0.0, 0.0, 0.0], [0, 100, 5]),
          ([0.0, 0.0, 0.00001], [0, 0, 0]), ([0.0, 0.0, 0.00001], [0, -0.00000001, 0.0000000001]),
          ([0.0, 0.0, 0.00001], [0.01, 0, 0]), ([pi, 0, 0], [4, -5, 0]),
          ([0.2, 0.5, 0.0], [0, 0, 0]), ([0.3, 0.5, 0.1], [2, 0, -7])]
    tg = [[0, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0], [0, 1, 0, 1, 0, 0], [0, -5, 10, 0, 0, 0], [-1, 1, 0, 0, 0, 1],
          [20, -1, 0, -1, 1, 0], [30, 5, -1, 20, -1, 0]]
    np.savez_compressed(os.path.join(OUT, "sophus_vectors.npz"), so3_omega=np.array([e[0] for e in el], float),
                        trans=np.array([e[1] for e in el], float), tangents=np.array(tg, float))


if __name__ == "__main__":
    bunny_pair(); dino_pair(); lm_golden(); sophus_vectors()
