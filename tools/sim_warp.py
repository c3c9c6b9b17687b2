"""Lock-step warp model on top of tools/sim_search.cpp: 32 consecutive src points (src tree order) form a warp; the main
loop costs one box-step slot if any lane does a box step and one leaf-step slot if any lane does a leaf step.
usage: python tools/sim_warp.py [libsim.so]"""
import ctypes as C, sys, os, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import oracle as O
lib = C.CDLL(sys.argv[1] if len(sys.argv) > 1 else '/tmp/libsim.so')
lib.sim_build.restype = C.c_void_p
z = np.load('/tmp/sim_scene.npz'); pts = [z['p1'], z['p2']]; gt = z['gt']; init = z['init']
N = len(pts[0])
dst = np.ascontiguousarray(pts[1]); h = C.c_void_p(lib.sim_build(dst.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(len(dst))))
src = np.ascontiguousarray(pts[0]); hs = C.c_void_p(lib.sim_build(src.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(len(src))))
order = np.zeros(N, np.int32); lib.sim_order(hs, order.ctypes.data_as(C.POINTER(C.c_int)))
rng = np.random.default_rng(1); W = 96
starts = rng.choice(N // 32 - 1, W, replace=False) * 32
ks = np.concatenate([order[s:s + 32] for s in starts])
kd = O.KdIndex(dst, 'kd')
buf = np.zeros(8192, np.uint8); lib.sim_set_trace(buf.ctypes.data_as(C.POINTER(C.c_ubyte)), 8192)
def run(name, poses, seed_idx):
    q = O.edge_queries(pts[0][ks], poses[0], poses[1])
    ri, rd = kd.closest_points(pts[0][ks], poses[0], poses[1], threads=8)
    cnt = (C.c_int64 * 4)(0, 0, 0, 0); traces = []
    for j in range(len(ks)):
        sl = -1 if seed_idx is None else lib.sim_leaf_of(h, int(seed_idx[j]))
        qq = np.ascontiguousarray(q[j]); r = lib.sim_query(h, qq.ctypes.data_as(C.POINTER(C.c_double)), sl, 1, cnt); assert r == ri[j]
        traces.append(buf[:lib.sim_trace_len()].copy())
    lane_steps = 0; slots_trip = 0; slots_type = 0; pro_lane = 0; pro_slots = 0
    for w in range(W):
        tw = traces[32 * w:32 * w + 32]
        main = [t[(t == 1) | (t == 2)] for t in tw]
        ml = max(len(m) for m in main)
        lane_steps += sum(len(m) for m in main); slots_trip += ml
        for i in range(ml):
            kinds = {m[i] for m in main if len(m) > i}
            slots_type += len(kinds)
        for code in (3, 4, 5):
            c = [int(np.sum(t == code)) for t in tw]; pro_lane += sum(c); pro_slots += max(c)
    print('%-26s main loop: lane-steps/query %6.1f | warp slots/warp: trip-only %6.1f  type-aware %6.1f | lanes busy %4.1f/32 (trip-only %4.1f) | prologue slots/warp %5.1f (lane-steps/query %4.1f) + 15 sweep' %
          (name, lane_steps / (32 * W), slots_trip / W, slots_type / W, lane_steps / slots_type, lane_steps / slots_trip, pro_slots / W, pro_lane / (32 * W)))
    return ri
i0 = run('far, cold', init, None)
half = init.copy(); half[:, :3, 3] = 0.5 * (init[:, :3, 3] + gt[:, :3, 3])
i1 = run('mid, stale seed', half, i0)
i2 = run('near (GT), seed from mid', gt, i1)
run('near (GT), seeded by itself', gt, i2)
