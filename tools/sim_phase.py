"""usage: python tools/sim_phase.py   (needs /tmp/sim_scene.npz from tools/sim_run.py)"""
import ctypes as C, sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import oracle as O
lib = C.CDLL('/tmp/libphase.so'); lib.ph_build.restype = C.c_void_p
z = np.load('/tmp/sim_scene.npz'); pts = [z['p1'], z['p2']]; gt = z['gt']; init = z['init']; N = len(pts[0])
dst = np.ascontiguousarray(pts[1]); h = C.c_void_p(lib.ph_build(dst.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(N)))
src = np.ascontiguousarray(pts[0]); hs = C.c_void_p(lib.ph_build(src.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(N)))
order = np.zeros(N, np.int32); lib.ph_order(hs, order.ctypes.data_as(C.POINTER(C.c_int)))
rng = np.random.default_rng(1); W = 96
starts = rng.choice(N // 32 - 1, W, replace=False) * 32
ks = np.concatenate([order[s:s + 32] for s in starts]); kd = O.KdIndex(dst, 'kd')
def run(name, poses, seed_idx, cap):
    q = np.ascontiguousarray(O.edge_queries(pts[0][ks], poses[0], poses[1]))
    ri, rd = kd.closest_points(pts[0][ks], poses[0], poses[1], threads=8)
    cnt = (C.c_int64 * 6)(*([0] * 6)); out = np.zeros(32, np.int32)
    for w in range(W):
        sl = np.array([-1 if seed_idx is None else lib.ph_leaf_of(h, int(seed_idx[32 * w + j])) for j in range(32)], np.int32)
        lib.ph_warp(h, q[32 * w:32 * w + 32].ctypes.data_as(C.POINTER(C.c_double)), sl.ctypes.data_as(C.POINTER(C.c_int)), cap, out.ctypes.data_as(C.POINTER(C.c_int)), cnt)
        assert np.array_equal(out, ri[32 * w:32 * w + 32])
    A, B, D, la, lb_, sw = [c / W for c in cnt]
    print('%-26s cap %2d: A slots %6.1f (lane box steps/query %5.1f)  B slots %5.1f (lane leaf scans/query %4.1f)  descent %4.1f  switches %4.1f  => est. instr/warp %6.0f' %
          (name, cap, A, la / 32, B, lb_ / 32, D, sw, A * 50 + B * 85 + D * 36 + 4 * 22 + 15 * 10))
    return ri
half = init.copy(); half[:, :3, 3] = 0.5 * (init[:, :3, 3] + gt[:, :3, 3])
for cap in (4, 8, 16):
    i0 = run('far, cold', init, None, cap)
    i1 = run('mid, stale seed', half, i0, cap)
    i2 = run('near (GT), seed from mid', gt, i1, cap)
    run('near (GT), seeded by itself', gt, i2, cap)
