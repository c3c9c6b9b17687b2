"""Seeded synthetic range scans of a bunny-shaped object (SURVEY 8(d) "synthetic inputs").

A union of ellipsoids (body, head, ears, tail, legs; ~0.15 m across) is ray-cast from M pinhole cameras on a ring
of radius 0.45 m, as a range scanner would see it: only the camera-facing surface, scan-line point order, a little
depth noise.  Exactly N points per view; every coordinate and normal component is rounded to fp32 so that the
engine's float4 storage is lossless and the CPU oracle sees bit-identical inputs.  Ground-truth poses are exactly
orthonormal; initial poses add the reference's noise model (common.h:38-67 addNoise: rotation sigma 0.02 rad per
axis applied on the right, translation sigma 0.01 m in world axes; main_multiview.cpp:42-43,78-84), frame 0 = GT.
No file of the reference is read: the generator is self-contained and runs on the GPU box.
"""
import numpy as np


def _rot(axis, a):
    c, s = np.cos(a), np.sin(a)
    if axis == 0:
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    if axis == 1:
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


# (centre, radii, rotation)
_PARTS = [
    ((0.000, 0.000, 0.000), (0.075, 0.040, 0.048), _rot(1, 0.25)),                   # body (elongated, pitched)
    ((0.070, 0.004, 0.050), (0.030, 0.024, 0.027), _rot(2, 0.3)),                    # head
    ((0.060, -0.016, 0.105), (0.010, 0.006, 0.042), _rot(0, 0.30) @ _rot(1, -0.35)), # ear
    ((0.064, 0.020, 0.100), (0.010, 0.006, 0.038), _rot(0, -0.45) @ _rot(1, -0.10)), # ear
    ((-0.078, 0.000, 0.020), (0.014, 0.014, 0.014), np.eye(3)),                      # tail
    ((-0.030, -0.036, -0.030), (0.040, 0.016, 0.022), _rot(2, 0.25)),                # hind leg
    ((-0.030, 0.036, -0.030), (0.040, 0.016, 0.022), _rot(2, -0.25)),                # hind leg
    ((0.050, -0.022, -0.040), (0.024, 0.010, 0.011), _rot(2, -0.1)),                 # front paw
    ((0.052, 0.020, -0.040), (0.022, 0.010, 0.011), _rot(2, 0.15)),                  # front paw
    ((0.098, 0.006, 0.046), (0.010, 0.012, 0.010), np.eye(3)),                       # nose
]
_CENTRE = np.array([0.01, 0.0, 0.03])


def _add_bumps():
    """Deterministic surface detail (small spheres half-sunk into body and head) so that the shape has no sliding
    directions: a union of a few smooth ellipsoids alone lets point-to-plane ICP rotate freely about their axes."""
    rng = np.random.default_rng(0xB077)
    for (c, r, R) in list(_PARTS[:2]):
        k = 0
        while k < 14:
            u = rng.normal(size=3); u /= np.linalg.norm(u)
            pos = np.asarray(c) + R @ (np.asarray(r) * u)
            rad = rng.uniform(0.006, 0.013)
            _PARTS.append((tuple(pos), (rad, rad * rng.uniform(0.6, 1.0), rad * rng.uniform(0.6, 1.0)),
                           _rot(0, rng.uniform(0, 3)) @ _rot(2, rng.uniform(0, 3))))
            k += 1


_add_bumps()


def _cast(o, d):
    """First hit of rays o + t d with the union; returns t (inf if miss) and outward unit normals."""
    n_rays = d.shape[0]
    best_t = np.full(n_rays, np.inf)
    best_n = np.zeros((n_rays, 3))
    for c, r, R in _PARTS:
        A = (R / np.asarray(r)).T          # unit-sphere space: A (x - c), A = diag(1/r) R^T
        oo = A @ (o - np.asarray(c))
        dd = d @ A.T
        a = np.einsum("ij,ij->i", dd, dd)
        b = dd @ oo
        cc = oo @ oo - 1.0
        disc = b * b - a * cc
        ok = disc > 0
        t = np.where(ok, (-b - np.sqrt(np.where(ok, disc, 0.0))) / a, np.inf)
        ok &= t > 0
        upd = ok & (t < best_t)
        if upd.any():
            x = o + t[upd, None] * d[upd]
            g = (A @ (x - np.asarray(c)).T).T @ A     # gradient of |A(x-c)|^2 (up to 2)
            g /= np.linalg.norm(g, axis=1, keepdims=True)
            best_t[upd] = t[upd]
            best_n[upd] = g
    return best_t, best_n


def _camera_pose(az, radius=0.45, height=0.05):
    """camera -> world; camera z looks at the object, x right, y down (range-image convention)."""
    pos = np.array([radius * np.cos(az), radius * np.sin(az), height])
    z = _CENTRE - pos
    z /= np.linalg.norm(z)
    up = np.array([0.0, 0.0, 1.0])
    x = np.cross(z, up); x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z], axis=1)
    # one Gram-Schmidt sweep in fp64 keeps |R^T R - I| at rounding level
    R[:, 0] /= np.linalg.norm(R[:, 0])
    R[:, 1] -= R[:, 0] * (R[:, 0] @ R[:, 1]); R[:, 1] /= np.linalg.norm(R[:, 1])
    R[:, 2] = np.cross(R[:, 0], R[:, 1])
    P = np.eye(4); P[:3, :3] = R; P[:3, 3] = pos
    return P


def _so3_exp(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)


def make_view(v, n_views, n_points, seed, depth_sigma=5e-5, half_fov=0.30):
    rng = np.random.default_rng(seed)
    P = _camera_pose(2.0 * np.pi * v / n_views)
    R, pos = P[:3, :3], P[:3, 3]

    def rays(res, jitter):
        u = (np.arange(res) + 0.5) / res * 2 - 1
        uu, vv = np.meshgrid(u, u)     # row-major scan lines
        uu = uu.ravel() * half_fov; vv = vv.ravel() * half_fov
        if jitter is not None:
            uu = uu + jitter[0]; vv = vv + jitter[1]
        d = np.stack([uu, vv, np.ones_like(uu)], axis=1)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        return d @ R.T

    t0, _ = _cast(pos, rays(256, None))
    frac = max(np.isfinite(t0).mean(), 1e-3)
    res = int(np.ceil(np.sqrt(n_points * 1.25 / frac)))
    pts, nor = [], []
    while True:
        pix = 2.0 * half_fov / res
        total = res * res
        got = 0
        pts, nor = [], []
        chunk = 1 << 20
        jit = rng.uniform(-0.35 * pix, 0.35 * pix, size=(2, total))
        d_all = rays(res, jit)
        for s in range(0, total, chunk):
            d = d_all[s:s + chunk]
            t, n = _cast(pos, d)
            hit = np.isfinite(t)
            tt = t[hit] + rng.normal(0.0, depth_sigma, size=int(hit.sum()))
            x = pos + tt[:, None] * d[hit]
            pts.append((x - pos) @ R)        # camera-local coordinates: R^T (x - pos)
            nor.append(n[hit] @ R)
            got += int(hit.sum())
        if got >= n_points:
            break
        res = int(res * 1.2) + 1
    pts = np.concatenate(pts); nor = np.concatenate(nor)
    keep = np.sort(rng.choice(len(pts), size=n_points, replace=False))   # exactly N, scan-line order kept
    pts = pts[keep].astype(np.float32).astype(np.float64)
    nor = nor[keep].astype(np.float32).astype(np.float64)
    return pts, nor, P


def scene_poses(n_views, config_id=3, sigma_rot=0.02, sigma_tra=0.01):
    """Ground-truth ring poses and the seeded noisy initial poses of a scene (no ray casting): (gt[M,4,4], init[M,4,4])."""
    gt, init = [], []
    for v in range(n_views):
        P = _camera_pose(2.0 * np.pi * v / n_views)
        gt.append(P)
        if v == 0:
            init.append(P.copy())
        else:
            rng = np.random.default_rng(0xA000 + 1000 * config_id + v)
            Q = P.copy()
            Q[:3, :3] = P[:3, :3] @ _so3_exp(rng.normal(0.0, sigma_rot, 3))
            Q[:3, 3] = P[:3, 3] + rng.normal(0.0, sigma_tra, 3)
            init.append(Q)
    return np.stack(gt), np.stack(init)


def make_scene(n_views, n_points, config_id=3, sigma_rot=0.02, sigma_tra=0.01):
    """Returns dict(pts, nor, poses_gt, poses_init); seeds 0xB200 + 1000*config + view."""
    pts, nor = [], []
    for v in range(n_views):
        p, n, _ = make_view(v, n_views, n_points, 0xB200 + 1000 * config_id + v)
        pts.append(p); nor.append(n)
    gt, init = scene_poses(n_views, config_id, sigma_rot, sigma_tra)
    return {"pts": pts, "nor": nor, "poses_gt": gt, "poses_init": init}


def ring_edges(n_views, knn=2):
    """The graph Frame::computePoseNeighboursKnn yields on an evenly spaced ring: each view's knn nearest views
    (ties broken by lower index, as a stable sort on float distances does)."""
    edges = []
    for i in range(n_views):
        cand = sorted(((min((i - j) % n_views, (j - i) % n_views), j) for j in range(n_views) if j != i))
        for q in range(min(knn, len(cand))):
            edges.append((i, cand[q][1]))
    return edges
