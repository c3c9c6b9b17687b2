"""The engine's own NN search text (csrc/knn.cuh: nn_query_init, nn_search, nn_search_warp; csrc/tree_build.h) compiled by g++
through tools/hostshim and run against a brute force in the reference's operation order: exact index (lowest on ties) and
bit-exact d^2 for near / far / on-point / out-of-box queries, no seed / right seed / stale seed, both storage modes, cloud
sizes around the leaf size.  Exercises the search LOGIC on the CPU; the hardware's roundings are covered by the GPU tests."""
import os
import subprocess

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
