"""Development aid: per-phase clock64() cycles of lm_step_kernel's solving launch (MVICP_STEP_PROFILE=1).  usage (GPU box):
MVICP_STEP_PROFILE=1 python tools/step_profile.py [views] [points]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MVICP_STEP_PROFILE"] = "1"
import mv_lm_icp_b200 as mv
from mv_lm_icp_b200 import synth, _lib
M = int(sys.argv[1]) if len(sys.argv) > 1 else 20; N = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
sc = synth.make_scene(M, N, config_id=3)
eng = mv.Engine(); eng.set_frames(sc["pts"], sc["nor"]); eng.set_graph(synth.ring_edges(M, 2)); eng.set_poses(sc["poses_init"])
opt = mv.default_options(); opt.max_num_iterations = 1      # one solving launch, then the loop stops: the stamps are that launch's
for rnd in range(3):
    eng.correspond(0.05); eng.optimize(mv.PARAM_SE3, mv.COST_P2PLANE, True, opt)
    p = np.zeros(16, np.int64)
    _lib.check(_lib.lib().mvicp_debug_step_profile(eng._ctx, p.ctypes.data_as(C.POINTER(C.c_longlong))))
    names = ["gather", "accept/take", "diag+build", "cholesky", "mcc+candidate", "writeback", "flag"]
    d = np.diff(p[:8])
    print(f"round {rnd}: total {p[7] - p[0]} cycles; " + ", ".join(f"{n} {v}" for n, v in zip(names, d)) + f"; cholesky: panel {p[8]} look-ahead/update {p[9]} barriers {p[10]}")
