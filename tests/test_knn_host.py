"""The engine's own NN search text (csrc/knn.cuh: nn_query_init, nn_search, nn_search_warp; csrc/tree_build.h) compiled by g++
through tools/hostshim and run against a brute force in the reference's operation order: exact index (lowest on ties) and
bit-exact d^2 for near / far / on-point / out-of-box queries, no seed / right seed / stale seed, both storage modes, cloud
sizes around the leaf size.  Exercises the search LOGIC on the CPU; the hardware's roundings are covered by the GPU tests."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("knnhost") / "knn_host_check")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    r = subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I" + os.path.join(ROOT, "tools", "hostshim"), "-I" + cuda_inc,
                        "-o", exe, os.path.join(ROOT, "tools", "knn_host_check.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.parametrize("mode", [0, 1], ids=["fp32-storage", "fp64-storage"])
@pytest.mark.parametrize("n", [1, 5, 8, 9, 17, 1000, 12345])
def test_search_text_is_exact_on_host(harness, n, mode):
    r = subprocess.run([harness, str(n), "600", str(11 + n), str(mode)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "0 mismatches" in r.stdout


@pytest.mark.parametrize("name", ["bunny_pair", "dino_pair"])
def test_search_text_reproduces_reference_nanoflann_goldens(harness, oracle, golden_dir, tmp_path, name):
    """Same text, the reference's own scans and the reference's own nanoflann answers (tests/golden): metre-scale bunny and
    millimetre-scale dinosaur, unseeded pass then a pass seeded with the first one's matches, both schedules."""
    g = np.load(f"{golden_dir}/{name}.npz")
    q = np.ascontiguousarray(oracle.edge_queries(g["pts1"], g["pose1"], g["pose0"]))   # frame.cpp:117-118,131,136 arithmetic
    path = tmp_path / f"{name}.bin"
    with open(path, "wb") as f:
        f.write(np.int64(len(g["pts0"])).tobytes()); f.write(np.int64(len(q)).tobytes())
        f.write(np.ascontiguousarray(g["pts0"], dtype=np.float64).tobytes()); f.write(q.tobytes())
        f.write(g["nn_idx"].astype(np.int32).tobytes()); f.write(g["nn_d2"].astype(np.float64).tobytes())
    r = subprocess.run([harness, "--file", str(path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "0 mismatches" in r.stdout, r.stdout[-2000:]
