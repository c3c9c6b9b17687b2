import numpy as np, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle import oracle as O
import mv_lm_icp_b200 as mv
from mv_lm_icp_b200.api import default_options
g = np.load('/root/repo/tests/golden/bunny_pair.npz')
pts = [g["pts0"][:8], g["pts0"], g["pts1"]]; nor = [g["nor0"][:8], g["nor0"], g["nor1"]]
poses = np.stack([np.eye(4), g["pose0"], g["pose1"]])
edges = [(1, 2), (2, 1)]
for param in (1,2):
  for cost in (0,1):
    for robust in (False, True):
      eng = mv.Engine(); eng.set_frames(pts, nor); eng.set_graph(edges); eng.set_poses(poses); eng.correspond(0.05)
      corr=[]; w=[]
      for e in range(2):
          f,s,d,ww = eng.get_edge(e); corr.append((f,s)); w.append(ww)
      for it in (1,2,8):
          eng.set_poses(poses)
          o = default_options(); o.max_num_iterations = it
          s = eng.optimize(param, cost, robust, o); P = eng.get_poses()
          oo = O.default_options(); oo.max_num_iterations = it
          Pr, sr, tr = O.optimize(pts, nor, poses, edges, corr, w, param=param, cost=cost, robust=robust, se3_autodiff=True, threads=8, options=oo)
          print(param, cost, robust, 'maxit', it, 'iters', s['num_iterations'], sr['num_iterations'], 'cost', s['final_cost'], sr['final_cost'], 'relcost %.2e'%(abs(s['final_cost']-sr['final_cost'])/sr['final_cost']), 'pose %.2e'%np.abs(P-Pr).max(), 'term', s['termination'], sr['termination'])
      eng.close()
