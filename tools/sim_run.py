"""Drive tools/sim_search.cpp: per-query step counts of the NN search for a far (initial poses) and a near (converged
poses, seeded) round.  usage: python tools/sim_run.py [libsim.so]"""
import ctypes as C, sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import oracle as O
from mv_lm_icp_b200 import synth
lib = C.CDLL(sys.argv[1] if len(sys.argv) > 1 else '/tmp/libsim.so')
lib.sim_build.restype = C.c_void_p
import os
ml = int(os.environ.get('ML', '8')); ratio = float(os.environ.get('RATIO', '0.5')); lib.sim_config(C.c_int(ml), C.c_double(ratio)); print('config max_leaves', ml, 'ratio', ratio)
M, N = 20, 200000
import os
cache = '/tmp/sim_scene.npz'
if os.path.exists(cache):
    z = np.load(cache); pts = [z['p1'], z['p2']]; gt = z['gt']; init = z['init']
else:
    pts = []; gt = []; init = []
    for v in (1, 2):
        p, n, P = synth.make_view(v, M, N, 0xB200 + 3000 + v); pts.append(p); gt.append(P)
        rng = np.random.default_rng(0xA000 + 3000 + v); Q = P.copy(); Q[:3, :3] = P[:3, :3] @ synth._so3_exp(rng.normal(0, .02, 3)); Q[:3, 3] += rng.normal(0, .01, 3); init.append(Q)
    gt = np.stack(gt); init = np.stack(init); np.savez(cache, p1=pts[0], p2=pts[1], gt=gt, init=init)
dst = np.ascontiguousarray(pts[1]); h = C.c_void_p(lib.sim_build(dst.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(len(dst))))
rng = np.random.default_rng(0); ks = rng.choice(N, 4000, replace=False)
kd = O.KdIndex(dst, 'kd')
lib.sim_set_cap.argtypes = [C.c_double]
def run(name, poses, seed_idx, reseed=1, cap=np.inf):
    lib.sim_set_cap(cap)
    q = O.edge_queries(pts[0][ks], poses[0], poses[1])
    ri, rd = kd.closest_points(pts[0][ks], poses[0], poses[1], threads=8)
    cnt = (C.c_int64 * 4)(0, 0, 0, 0); bad = 0
    for j in range(len(ks)):
        sl = -1 if seed_idx is None or seed_idx[j] < 0 else lib.sim_leaf_of(h, int(seed_idx[j]))
        qq = np.ascontiguousarray(q[j]); r = lib.sim_query(h, qq.ctypes.data_as(C.POINTER(C.c_double)), sl, reseed, cnt)
        bad += int(r != ri[j] and rd[j] < 0.05 ** 2)
    n = len(ks)
    print('%-28s box tests %7.1f  point tests %7.1f  steps %7.1f  plane tests %5.1f  mismatches %d  (median nn dist %.2e)' % (name, cnt[0] / n, cnt[1] / n, cnt[2] / n, cnt[3] / n, bad, np.sqrt(np.median(rd))))
    print('      inlier fraction %.3f' % np.mean(rd < 0.05 ** 2))
    return np.where(rd < cap, ri, -1) if np.isfinite(cap) else ri
i0 = run('far, cold', init, None)
# stale seeds: NN under initial poses used as seed after poses moved half-way to GT
half = init.copy(); half[:, :3, 3] = 0.5 * (init[:, :3, 3] + gt[:, :3, 3])
i1 = run('mid, stale seed (reseed on)', half, i0, 1)
run('mid, stale seed (reseed off)', half, i0, 0)
i2 = run('near (GT), seed from mid', gt, i1, 1)
run('near (GT), seeded by itself', gt, i2, 1)

for c in (1, 2, 3):
    print('--- coarse levels', c, '(cold / stale queries use leaves of', 8 << c, 'points)')
    lib.sim_set_coarse(c)
    i0 = run('far, cold', init, None)
    i1 = run('mid, stale seed (reseed on)', half, i0, 1)
    i2 = run('near (GT), seed from mid', gt, i1, 1)
    run('near (GT), seeded by itself', gt, i2, 1)
