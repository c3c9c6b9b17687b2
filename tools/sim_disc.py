"""usage: python tools/sim_disc.py   (needs /tmp/sim_scene.npz from tools/sim_run.py; builds /tmp/libsimdisc.so itself)"""
import ctypes as C, subprocess, sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I/usr/local/cuda/include", "-o", "/tmp/libsimdisc.so", os.path.join(ROOT, "tools", "sim_disc.cpp")], check=True)
lib = C.CDLL("/tmp/libsimdisc.so"); lib.sd_build.restype = C.c_void_p
z = np.load('/tmp/sim_scene.npz'); pts = [z['p1'], z['p2']]; gt = z['gt']; init = z['init']; N = len(pts[0])
dst = np.ascontiguousarray(pts[1]); h = C.c_void_p(lib.sd_build(dst.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(N)))
rng = np.random.default_rng(0); ks = rng.choice(N, 3000, replace=False); kd = O.KdIndex(dst, 'kd')
def run(name, poses, seed_idx):
    q = np.ascontiguousarray(O.edge_queries(pts[0][ks], poses[0], poses[1])); ri, rd = kd.closest_points(pts[0][ks], poses[0], poses[1], threads=8)
    for mode, mname in ((0, 'AABB'), (1, 'disc'), (2, 'max(AABB, disc)')):
        lib.sd_mode(mode); cnt = (C.c_int64 * 3)(0, 0, 0); bad = 0
        for j in range(len(ks)):
            sl = -1 if seed_idx is None else lib.sd_leaf_of(h, int(seed_idx[j]))
            bad += int(lib.sd_query(h, q[j].ctypes.data_as(C.POINTER(C.c_double)), sl, cnt) != ri[j])
        n = len(ks)
        print('%-24s %-16s bound tests %7.1f  point tests %7.1f  plane tests %5.1f  mismatches %d' % (name, mname, cnt[0] / n, cnt[1] / n, cnt[2] / n, bad))
    return ri
i0 = run('far, cold', init, None)
half = init.copy(); half[:, :3, 3] = 0.5 * (init[:, :3, 3] + gt[:, :3, 3])
i1 = run('mid, stale seed', half, i0)
i2 = run('near (GT), seed from mid', gt, i1)
run('near (GT), own seed', gt, i2)
