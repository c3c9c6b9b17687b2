"""Small end-to-end exercise of every kernel family, meant to be run under compute-sanitizer on a GPU box:
  compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck --error-exitcode 1 python tools/sanitize_run.py
(fp32 and fp64 storage, all three parameterisations, robust and plain, general non-rigid path, pairwise, normals, k-NN,
single closest point, every NN schedule).  Prints DONE at the end; results are not checked here (tests/ do that)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mv_lm_icp_b200 as mv
from mv_lm_icp_b200 import synth
from mv_lm_icp_b200.api import FLAG_HOST_BUILD, FLAG_NO_ADJ, FLAG_NO_OBB, FLAG_NO_SEED, ICP_Ceres, default_options

sc = synth.make_scene(4, 1501, config_id=1)
edges = synth.ring_edges(4, 2)
for flags in (0, FLAG_NO_SEED, FLAG_NO_OBB, FLAG_HOST_BUILD | FLAG_NO_ADJ):
    eng = mv.Engine(device=0, flags=flags)
    eng.set_frames(sc["pts"], sc["nor"]); eng.set_graph(edges); eng.set_poses(sc["poses_init"])
    for param in (mv.PARAM_SE3, mv.PARAM_QUAT, mv.PARAM_AA):
        for cost in (mv.COST_P2PLANE, mv.COST_P2P, mv.COST_MIXED):
            eng.set_poses(sc["poses_init"])
            eng.icp_round(0.05, param, cost, param != mv.PARAM_AA)
    eng.get_edge(len(edges) - 1); eng.get_nn(len(edges) - 1)
    eng.closest_point(1, sc["pts"][2][7])
    eng.recompute_normals(10); eng.knn_self(0, 5)
    eng.icp_round(0.05, mv.PARAM_SE3, mv.COST_P2PLANE, True)
    eng.close()
# converged rounds: guessed select, certificates written, then certified rounds (streaming kernel + searched list), both storages
none = default_options(); none.max_num_iterations = 0
for cloud in (sc["pts"], [p + np.random.default_rng(4).normal(0, 1e-9, p.shape) for p in sc["pts"]]):
    eng = mv.Engine(device=0)
    eng.set_frames(cloud, sc["nor"]); eng.set_graph(edges); eng.set_poses(sc["poses_gt"])
    eng.correspond(0.05); eng.optimize(options=none)
    for _ in range(4):
        eng.correspond(0.05)
    one = default_options(); one.max_num_iterations = 1
    for _ in range(3):
        eng.optimize(options=one); eng.correspond(0.05)
    st = eng.stats(); assert st["cert_rounds"] >= 2 and st["select_guess_rounds"] >= 2, st
    eng.pull_all_edges()
    eng.close()
# fp64 storage (coordinates not fp32-representable) + non-rigid poses (general LM path)
rng = np.random.default_rng(3)
pts = [p + rng.normal(0, 1e-9, p.shape) for p in sc["pts"][:3]]
poses = sc["poses_init"][:3].copy()
poses[1][:3, :3] *= 1.0 + 1e-4                      # not a rotation: takes the general frame model
eng = mv.Engine(device=0)
eng.set_frames(pts, sc["nor"][:3]); eng.set_graph(synth.ring_edges(3, 2)); eng.set_poses(poses)
for param in (mv.PARAM_SE3, mv.PARAM_QUAT):
    eng.set_poses(poses); eng.icp_round(0.05, param, mv.COST_P2PLANE, True)
    eng.set_poses(poses); eng.icp_round(0.05, param, mv.COST_P2P, False)
eng.close()
# 40 views: the Cholesky factor lives in global memory (n = 234 > 166)
sc40 = synth.make_scene(40, 400, config_id=2)
eng = mv.Engine(device=0)
eng.set_frames(sc40["pts"], sc40["nor"]); eng.set_graph(synth.ring_edges(40, 2)); eng.set_poses(sc40["poses_init"])
eng.icp_round(0.05, mv.PARAM_SE3, mv.COST_P2PLANE, True); eng.icp_round(0.05, mv.PARAM_QUAT, mv.COST_P2P, False)
eng.close()
# pairwise solvers (icp-ceres.h:30-36)
src = sc["pts"][0][:800]; T = sc["poses_init"][1]
dst = src @ T[:3, :3].T + T[:3, 3]
ICP_Ceres.pointToPoint_SophusSE3(src, dst)
ICP_Ceres.pointToPlane_SophusSE3(src, dst, sc["nor"][0][:800] @ T[:3, :3].T) if hasattr(ICP_Ceres, "pointToPlane_SophusSE3") else None
print("DONE")
