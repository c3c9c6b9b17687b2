"""Second reference dataset on the device (samples/dinosaur: millimetre units).  Golden = the reference's own nanoflann
(tests/golden/make_golden.py:dino_pair).  Added after the round's GPU budget was spent (checked on the host model and by
tests/test_knn_host.py); the file is named to run after the established GPU tests."""
import numpy as np
import pytest

from mv_lm_icp_b200 import Engine

pytestmark = pytest.mark.gpu


def test_real_dinosaur_pair_mm_units(golden_dir):
    """The reference's second sample set: millimetre units (|x| up to 686, so the fp32 screening allowance is ~1e-3 mm and
    matches are tens of mm away), 5-digit non-orthonormal poses, cutoff 25.  Golden = the reference's own nanoflann."""
    g = np.load(f"{golden_dir}/dino_pair.npz")
    eng = Engine()
    eng.set_frames([g["pts0"], g["pts1"]], [g["nor0"], g["nor1"]])
    eng.set_graph([(1, 0)])
    eng.set_poses([g["pose0"], g["pose1"]])
    for rnd in range(2):                                     # second pass is seeded by the first: same answer
        eng.correspond(25.0)
        idx, d2 = eng.get_nn(0)
        assert np.array_equal(d2.view(np.uint64), g["nn_d2"].view(np.uint64))
        assert np.array_equal(idx, g["nn_idx"])
        f, s, dist, w = eng.get_edge(0)
        assert np.array_equal(f, g["first"]) and np.array_equal(s, g["second"])
        assert np.array_equal(dist.view(np.uint64), g["dist"].view(np.uint64))
        assert np.float32(w).view(np.uint32) == np.float32(g["weight"]).view(np.uint32)
    eng.close()
