"""GPU parity of the correspondence step (mvicp_correspond) against the CPU oracle, through the C ABI.
Bar: nearest-neighbour indices and squared distances bit-exact, inlier lists identical, float weight bit-exact."""
import numpy as np
import pytest

from helpers import oracle_correspond, scene
from mv_lm_icp_b200 import Engine, synth
from mv_lm_icp_b200.api import FLAG_HOST_BUILD, FLAG_NO_ADJ, FLAG_NO_OBB, FLAG_NO_SEED, FLAG_NO_CERT, FLAG_NO_SELECT_GUESS, FLAG_STEP_LOOP, default_options

pytestmark = pytest.mark.gpu


def _check_edges(eng, ref, edges):
    ties = 0
    for e, (s, d) in enumerate(edges):
        if ref[e] is None:
            cnt, w = eng.get_edge(e, arrays=False)
            assert cnt == 0
            continue
        idx, d2 = eng.get_nn(e)
        r = ref[e]
        assert np.array_equal(d2.view(np.uint64), r["nn_d2"].view(np.uint64)), f"edge {e}: d2 not bit-exact"
        diff = idx != r["nn_idx"]
        ties += int(diff.sum())   # equal d2 but different index = exact tie (tie rule differs only vs nanoflann)
        assert not diff.any(), f"edge {e}: {diff.sum()} index mismatches"
        f, sec, dist, w = eng.get_edge(e)
        assert np.array_equal(f, r["first"]) and np.array_equal(sec, r["second"])
        assert np.array_equal(dist.view(np.uint64), r["dist"].view(np.uint64))
        assert np.float32(w).view(np.uint32) == np.float32(r["weight"]).view(np.uint32), (w, r["weight"])
    return ties


@pytest.mark.parametrize("n_views,n_points,cfg", [(4, 5000, 21), (6, 20011, 22)])
def test_synthetic_bit_exact(oracle, n_views, n_points, cfg):
    sc = scene(n_views, n_points, cfg)
    edges = synth.ring_edges(n_views, 2)
    eng = Engine()
    eng.set_frames(sc["pts"], sc["nor"])
    eng.set_graph(edges)
    for poses in (sc["poses_init"], sc["poses_gt"]):   # far and near queries; second call is seeded by the first
        eng.set_poses(poses)
        eng.correspond(0.05)
        ref = oracle_correspond(oracle, sc["pts"], poses, edges)
        _check_edges(eng, ref, edges)
    eng.close()


SCHEDULES = (FLAG_NO_SEED, FLAG_NO_OBB, FLAG_NO_OBB | FLAG_NO_SEED, FLAG_HOST_BUILD, FLAG_HOST_BUILD | FLAG_NO_SEED, FLAG_HOST_BUILD | FLAG_NO_OBB,
             FLAG_NO_ADJ, FLAG_NO_ADJ | FLAG_HOST_BUILD, FLAG_NO_ADJ | FLAG_NO_OBB | FLAG_NO_SEED,
             FLAG_STEP_LOOP, FLAG_STEP_LOOP | FLAG_NO_SEED, FLAG_STEP_LOOP | FLAG_NO_ADJ | FLAG_NO_OBB)


def test_seed_and_schedule_do_not_change_results(oracle, extra_flags=()):
    """Seeded / unseeded, with / without the per-leaf neighbour lists that let a seeded query skip the tree walk (adjacency.h), with /
    without the oriented node boxes that the far rounds search (far.cuh), search trees built on the device (tree_gpu.cuh, default)
    or on the host (tree_build.h): the same exact search, bit-identical output.  Odd cloud sizes leave partially filled warps and
    padding leaves in play."""
    sc = scene(4, 5003, 21)
    edges = synth.ring_edges(4, 2)
    res = []
    for flags in (0,) + SCHEDULES + tuple(extra_flags):
        eng = Engine(flags=flags)
        eng.set_frames(sc["pts"], sc["nor"]); eng.set_graph(edges)
        for poses in (sc["poses_init"], sc["poses_gt"], sc["poses_init"], sc["poses_init"]):   # cold, stale seeds twice, exact seeds
            eng.set_poses(poses); eng.correspond(0.05)
        res.append([eng.get_nn(e) for e in range(len(edges)) if edges[e][0] != 0])
        eng.close()
    for other in res[1:]:
        for (i0, d0), (i1, d1) in zip(res[0], other):
            assert np.array_equal(i0, i1) and np.array_equal(d0.view(np.uint64), d1.view(np.uint64))


def test_all_edges_in_one_call_equal_the_per_edge_lists(oracle):
    """mvicp_get_all_edges (device-side compaction into the reference's 16-byte Correspondance records, one copy back) against
    mvicp_get_edge per edge: same (first, second, dist) in the same order, same weights; edges of the fixed frame are empty."""
    sc = scene(4, 5003, 21)
    edges = synth.ring_edges(4, 2)
    eng = Engine(); eng.set_frames(sc["pts"], sc["nor"]); eng.set_graph(edges)
    for poses, cut in ((sc["poses_init"], 0.05), (sc["poses_gt"], 0.0004)):     # nearly all inliers / a sparse subset
        eng.set_poses(poses); eng.correspond(cut)
        nb = eng.pull_all_edges()
        assert nb == int(eng.edge_offsets[-1]) * 16 + 4 * len(edges) + 8 * (len(edges) + 1)
        for e, (s, d) in enumerate(edges):
            rec, w = eng.host_edges[e]
            if s == 0:
                assert len(rec) == 0
                continue
            f, sec, dist, ww = eng.get_edge(e)
            assert np.array_equal(rec["first"], f) and np.array_equal(rec["second"], sec)
            assert np.array_equal(rec["dist"].view(np.uint64), dist.view(np.uint64))
            assert np.float32(w).view(np.uint32) == np.float32(ww).view(np.uint32)
        eng.pull_all_edges(records=False)
        assert all(r is None for r, _ in eng.host_edges)
    eng.close()


def test_device_built_tree_is_a_valid_left_balanced_kd_tree():
    """The structure tree_gpu.cuh builds, read back through the search: every point of a cloud is its own nearest neighbour at
    distance 0 (ragged sizes around the leaf / level boundaries, a cloud with many equal coordinates, fp64 storage)."""
    rng = np.random.default_rng(17)
    clouds = [rng.normal(size=(n, 3)).astype(np.float32).astype(np.float64) * 0.05 for n in (1, 8, 9, 17, 64, 65, 1000, 4097, 30011)]
    grid = np.stack(np.meshgrid(np.arange(20), np.arange(20), np.arange(3)), -1).reshape(-1, 3).astype(np.float64) * 0.01   # ties along every axis
    clouds.append(grid)
    clouds.append(rng.normal(size=(5000, 3)) * 0.05)      # not fp32-representable: every frame takes the fp64 records
    for fp64 in (False, True):
        cs = clouds if fp64 else clouds[:-1]
        M = len(cs)
        eng = Engine(); eng.set_frames(cs + cs, None)      # frame M + i is a copy of frame i: edge (M + i -> i) must match index to index
        eng.set_graph([(M + i, i) for i in range(M)]); eng.set_poses([np.eye(4)] * (2 * M))
        eng.correspond(0.05)
        for i in range(M):
            idx, d2 = eng.get_nn(i)
            assert np.all(d2 == 0.0) and np.array_equal(idx, np.arange(len(cs[i]))), (fp64, i)
        eng.close()


def test_real_bunny_pair_fp64_storage(oracle, golden_dir):
    """Config 1: real scans (not fp32-representable -> fp64 storage path), non-rigid sample poses, general inverse.
    Golden answer = the reference's own nanoflann (tests/golden/make_golden.py)."""
    g = np.load(f"{golden_dir}/bunny_pair.npz")
    eng = Engine()
    eng.set_frames([g["pts0"], g["pts1"]], [g["nor0"], g["nor1"]])
    eng.set_graph([(1, 0)])
    eng.set_poses([g["pose0"], g["pose1"]])
    eng.correspond(0.05)
    idx, d2 = eng.get_nn(0)
    assert np.array_equal(d2.view(np.uint64), g["nn_d2"].view(np.uint64))
    assert np.array_equal(idx, g["nn_idx"])
    f, s, dist, w = eng.get_edge(0)
    assert np.array_equal(f, g["first"]) and np.array_equal(s, g["second"])
    assert np.array_equal(dist.view(np.uint64), g["dist"].view(np.uint64))
    assert np.float32(w).view(np.uint32) == np.float32(g["weight"]).view(np.uint32)
    eng.close()


def test_edge_cases(oracle):
    rng = np.random.default_rng(5)
    # tiny clouds (1, 7, 9 points: below / around one leaf), ragged sizes, a frame far away (all outliers)
    pts = [rng.normal(size=(n, 3)).astype(np.float32).astype(np.float64) * 0.01 for n in (1, 7, 9, 300)]
    pts.append(pts[3][:50] + 10.0)
    pts[4] = pts[4].astype(np.float32).astype(np.float64)
    poses = [np.eye(4) for _ in pts]
    edges = [(1, 0), (2, 1), (3, 2), (1, 3), (4, 3), (3, 4), (2, 0)]
    eng = Engine()
    eng.set_frames(pts, None)
    eng.set_graph(edges); eng.set_poses(poses)
    eng.correspond(0.05)
    ref = oracle_correspond(oracle, pts, poses, edges, kind="brute")
    _check_edges(eng, ref, edges)
    cnt, w = eng.get_edge(4, arrays=False)   # frame 4 is 10 m away: no inliers; oracle choice weight = 0
    assert cnt == 0 and w == 0.0
    # exact duplicates in the dst cloud: tie -> lowest index, same as the brute-force oracle
    dup = np.repeat(pts[3][:20], 2, axis=0)
    eng2 = Engine(); eng2.set_frames([dup, pts[3]], None); eng2.set_graph([(1, 0)]); eng2.set_poses([np.eye(4)] * 2)
    eng2.correspond(0.05)
    ref2 = oracle_correspond(oracle, [dup, pts[3]], [np.eye(4)] * 2, [(1, 0)], kind="brute")
    _check_edges(eng2, ref2, [(1, 0)])
    eng.close(); eng2.close()


def test_median_with_masses_of_near_equal_distances(oracle):
    """6000 query points on a thin spherical shell around a one-point dst cloud: every squared distance shares its
    leading 22 bits, which overflows the median select's candidate buffer and takes its scan-the-edge path; plus an
    all-equal case (exact ties in the order statistic)."""
    rng = np.random.default_rng(9)
    u = rng.normal(size=(6000, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    shell = (u * (0.01 * (1 + rng.uniform(0, 1e-5, size=(6000, 1))))).astype(np.float32).astype(np.float64)
    same = np.tile(np.array([[0.0078125, 0.0, 0.0]]), (5000, 1))
    dst = np.zeros((1, 3))
    pts = [dst, shell, same]
    edges = [(1, 0), (2, 0)]
    eng = Engine(); eng.set_frames(pts, None); eng.set_graph(edges); eng.set_poses([np.eye(4)] * 3)
    eng.correspond(0.05)
    ref = oracle_correspond(oracle, pts, [np.eye(4)] * 3, edges, kind="brute")
    _check_edges(eng, ref, edges)
    eng.close()


def _edge_meta(eng, edges):
    return [eng.get_edge(e, arrays=False) for e in range(len(edges))]


def test_guessed_median_select_is_exact(oracle):
    """In a round that follows a one-iteration solve the NN kernel's epilogue checks the previous median's 22-bit prefix as a guess and
    the three select passes are skipped (select.cuh).  Inlier counts and weights must be bit-identical to the full select: (a) on a
    run that converges (guesses hit), (b) while the poses still jump (one LM iteration per round from the noisy start: guesses miss,
    the edge is redone from scratch), and the final round is checked against the oracle."""
    sc = scene(4, 5003, 21)
    edges = synth.ring_edges(4, 2)
    one = default_options(); one.max_num_iterations = 1
    for warm in (8, 0):      # (a) 8 full solves first, then one-iteration rounds that barely move; (b) one-iteration rounds from the start
        engs = [Engine(flags=f) for f in (0, FLAG_NO_SELECT_GUESS)]
        for eng in engs:
            eng.set_frames(sc["pts"], sc["nor"]); eng.set_graph(edges); eng.set_poses(sc["poses_init"])
        for rnd in range(warm + 6):
            metas = []
            for eng in engs:
                eng.correspond(0.05)
                metas.append(_edge_meta(eng, edges))
                eng.optimize(options=None if rnd < warm else one)
            for (c0, w0), (c1, w1) in zip(*metas):
                assert c0 == c1 and np.float32(w0).view(np.uint32) == np.float32(w1).view(np.uint32), (rnd, c0, c1, w0, w1)
        assert np.array_equal(engs[0].get_poses(), engs[1].get_poses())
        st = [eng.stats() for eng in engs]
        assert st[0]["select_guess_rounds"] >= 3 and st[1]["select_guess_rounds"] == 0, st
        if warm:
            assert st[0]["select_guess_misses"] <= 2, st[0]        # converged: the prefix hardly moves any more
        else:
            assert st[0]["select_guess_misses"] > 0, st[0]         # still moving: this run is the test of the fallback
        poses = engs[0].get_poses()
        engs[0].correspond(0.05)
        ref = oracle_correspond(oracle, sc["pts"], poses, edges)
        _check_edges(engs[0], ref, edges)
        for eng in engs:
            eng.close()


def test_certified_matches_are_exact(oracle):
    """Converged rounds keep a query's previous match without searching when the previous search left a margin -- its match beats
    every other point by m -- and the edge has moved by less than m/2 since (knn.cuh, CERT).  Matches and distances must be
    bit-identical to searching every query: (a) after a run of full solves, with the poses then standing still (every query with a
    positive margin is kept), (b) with one LM iteration per round from the noisy start, where the clouds keep moving by more than
    many margins (those queries are collected per CTA, searched and re-certified); the final round is checked against the oracle."""
    sc = scene(4, 5003, 21)
    edges = synth.ring_edges(4, 2)
    one = default_options(); one.max_num_iterations = 1
    none = default_options(); none.max_num_iterations = 0
    for warm, late in ((8, none), (0, one)):
        engs = [Engine(flags=f) for f in (0, FLAG_NO_CERT, FLAG_NO_CERT | FLAG_NO_SELECT_GUESS)]
        for eng in engs:
            eng.set_frames(sc["pts"], sc["nor"]); eng.set_graph(edges); eng.set_poses(sc["poses_init"])
        for rnd in range(warm + 7):
            nn = []
            for eng in engs:
                eng.correspond(0.05)
                nn.append([eng.get_nn(e) for e in range(len(edges)) if edges[e][0] != 0] + [_edge_meta(eng, edges)])
                eng.optimize(options=None if rnd < warm else late)
            for other in nn[1:]:
                for (i0, d0), (i1, d1) in zip(nn[0][:-1], other[:-1]):
                    assert np.array_equal(i0, i1) and np.array_equal(d0.view(np.uint64), d1.view(np.uint64)), rnd
                assert nn[0][-1] == other[-1], rnd
        for eng in engs[1:]:
            assert np.array_equal(engs[0].get_poses(), eng.get_poses())
        st = [eng.stats() for eng in engs]
        n_q = st[0]["queries"]
        assert st[0]["cert_rounds"] >= 3 and st[1]["cert_rounds"] == 0 and st[1]["cert_reused"] == 0, st
        frac = st[0]["cert_reused"] / (n_q * st[0]["cert_rounds"])
        if warm:
            assert frac > 0.9, st[0]             # the poses stand still: kept unless a pruned box sat just outside the bound (tiny margin)
        else:
            assert 0 < frac < 0.5, st[0]         # still moving by more than most margins: those queries are searched again
        poses = engs[0].get_poses()
        engs[0].correspond(0.05)
        ref = oracle_correspond(oracle, sc["pts"], poses, edges)
        _check_edges(engs[0], ref, edges)
        for eng in engs:
            eng.close()


def test_certificates_with_ties_and_duplicates(oracle):
    """Exact ties never certify: a dst cloud in which every point exists twice (the lower index must win every round), queried
    by a copy of itself moved by a rigid transform that the solve then removes."""
    rng = np.random.default_rng(5)
    base = rng.uniform(-0.1, 0.1, size=(3000, 3)).astype(np.float32).astype(np.float64)
    dup = np.concatenate([base, base[::-1]])
    nor = np.tile(np.array([[0.0, 0.0, -1.0]]), (6000, 1))
    pts = [dup, base.copy()]; nors = [nor, nor[:3000]]
    edges = [(1, 0), (0, 1)]
    P = [np.eye(4), np.eye(4)]; P[1][:3, 3] = [1e-4, -2e-4, 1e-4]
    none = default_options(); none.max_num_iterations = 0
    engs = [Engine(flags=f) for f in (0, FLAG_NO_CERT)]
    res = []
    for eng in engs:
        eng.set_frames(pts, nors); eng.set_graph(edges); eng.set_poses(P)
        eng.correspond(0.05); eng.optimize(cost=0, options=none)
        out = []
        for _ in range(4):
            eng.correspond(0.05); out.append(eng.get_nn(0))
        res.append(out)
    for (i0, d0), (i1, d1) in zip(*res):
        assert np.array_equal(i0, i1) and np.array_equal(d0.view(np.uint64), d1.view(np.uint64))
    idx = res[0][-1][0]
    assert np.array_equal(idx, np.arange(3000))           # of the two copies the lower index
    st = engs[0].stats()
    assert st["cert_rounds"] >= 1 and st["cert_reused"] == 0, st   # every margin is zero
    for eng in engs:
        eng.close()


def test_guessed_median_select_degenerate_buckets(oracle):
    """The shell / all-equal clouds of the test above under the guessed select: the guessed bucket holds more keys than the candidate
    buffer (miss by overflow -> scan of the edge), and the empty edge of a fixed src frame stays empty."""
    rng = np.random.default_rng(9)
    u = rng.normal(size=(6000, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    shell = (u * (0.01 * (1 + rng.uniform(0, 1e-5, size=(6000, 1))))).astype(np.float32).astype(np.float64)
    same = np.tile(np.array([[0.0078125, 0.0, 0.0]]), (5000, 1))
    pts = [np.zeros((1, 3)), shell, same]
    edges = [(1, 0), (2, 0), (0, 1)]
    none = default_options(); none.max_num_iterations = 0
    eng = Engine(); eng.set_frames(pts, None); eng.set_graph(edges); eng.set_poses([np.eye(4)] * 3)
    eng.correspond(0.05)
    eng.optimize(cost=0, options=none)       # zero iterations: the poses stay, the next rounds count as converged
    for _ in range(3):
        eng.correspond(0.05)
    assert eng.stats()["select_guess_rounds"] >= 1
    ref = oracle_correspond(oracle, pts, [np.eye(4)] * 3, edges, kind="brute")
    _check_edges(eng, ref, edges)
    eng.close()


def test_cutoff_boundary_is_the_reference_comparison(oracle):
    """`sqrt(d2) < cutoff` (frame.cpp:142,156) is evaluated on the device as d2 <= the largest double whose square root is below the
    cutoff: points within a few ulps of the cutoff distance, on both sides, must be classified as the oracle's sqrt comparison does."""
    for cutoff in (np.float32(0.05), np.float32(0.3), np.float32(1e-3)):
        t = float(cutoff)
        xs = [t * (1.0 + k * 2.0 ** -52) for k in range(-6, 7)] + [np.nextafter(t, 0), np.nextafter(t, 1), t, 0.5 * t, 2 * t]
        src = np.array([[x, 0.0, 0.0] for x in xs] + [[0.0, x, 0.0] for x in xs] + [[x / np.sqrt(3), x / np.sqrt(3), x / np.sqrt(3)] for x in xs])
        pts = [np.zeros((1, 3)), src]
        edges = [(1, 0)]
        eng = Engine(); eng.set_frames(pts, None); eng.set_graph(edges); eng.set_poses([np.eye(4)] * 2)
        eng.correspond(float(cutoff))
        ref = []
        for s_, d_ in edges:
            idx = oracle.KdIndex(pts[d_], "brute")
            i, d2 = idx.closest_points(pts[s_], np.eye(4), np.eye(4))
            f, sec, dist, w, med = oracle.filter_edge(i, d2, cutoff)
            ref.append(dict(first=f, second=sec, dist=dist, weight=w, nn_idx=i, nn_d2=d2, median=med))
        _check_edges(eng, ref, edges)
        assert 0 < len(ref[0]["first"]) < len(src)
        eng.close()


def test_closest_point_api(oracle):
    sc = scene(4, 5000, 21)
    eng = Engine(); eng.set_frames(sc["pts"], sc["nor"])
    kd = oracle.KdIndex(sc["pts"][2], "kd")
    rng = np.random.default_rng(1)
    for _ in range(20):
        q = sc["pts"][2][rng.integers(5000)] + rng.normal(0, 0.003, 3)
        i, d2 = eng.closest_point(2, q)
        ri, rd = kd.closest_points(q[None], np.eye(4), np.eye(4))
        assert i == ri[0] and d2 == rd[0]
    eng.close()


def test_full_size_properties(oracle):
    """BASELINE config 3 shape (20 x 200k): size-independent checks + oracle on a sample of queries."""
    M, N = 20, 200_000
    sc = scene(M, N, 3)
    edges = synth.ring_edges(M, 2)
    eng = Engine(); eng.set_frames(sc["pts"], sc["nor"]); eng.set_graph(edges); eng.set_poses(sc["poses_init"])
    eng.correspond(0.05)
    rng = np.random.default_rng(0)
    for e in rng.choice(len(edges), 4, replace=False):
        s, d = edges[e]
        if s == 0:
            continue
        idx, d2 = eng.get_nn(e)
        assert idx.min() >= 0 and idx.max() < N
        # (a) the reported distance is the distance to the reported point, in the reference's arithmetic
        ks = rng.choice(N, 2000, replace=False)
        q = oracle.edge_queries(sc["pts"][s][ks], sc["poses_init"][s], sc["poses_init"][d])
        p = sc["pts"][d][idx[ks]]
        dd = q - p
        assert np.array_equal(((dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1]) + dd[:, 2] * dd[:, 2]).view(np.uint64), d2[ks].view(np.uint64))
        # (b) exactness on the sample against the oracle tree
        ri, rd = oracle.KdIndex(sc["pts"][d], "kd").closest_points(sc["pts"][s][ks], sc["poses_init"][s], sc["poses_init"][d], threads=8)
        assert np.array_equal(ri, idx[ks]) and np.array_equal(rd.view(np.uint64), d2[ks].view(np.uint64))
        # (c) list is sorted by src index, weight = float(1.5 * upper median of the inlier distances)
        f, sec, dist, w = eng.get_edge(e)
        assert np.all(np.diff(f) > 0)
        med = np.sort(dist)[len(dist) // 2]
        assert np.float32(w) == np.float32(med * 1.5)
    eng.close()
