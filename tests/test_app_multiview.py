"""Headless C++ driver (apps/multiview_main.cpp): the reference's on-disk formats + loop over the C ABI."""
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "apps", "multiview_b200")
BIN_PAIR = os.path.join(ROOT, "apps", "pairwise_b200")


def _build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "apps")], check=True, capture_output=True)
    assert os.path.exists(BIN) and os.path.exists(BIN_PAIR)


def _write_scene(d, sc, with_gt=True, stride=1):
    """cloud%d.xyz / pose%d.txt / groundtruth%d.txt as the reference stores them (samples/Bunny_RealData)."""
    for i, (p, n) in enumerate(zip(sc["pts"], sc["nor"])):
        with open(os.path.join(d, f"cloud{i * stride}.xyz"), "w") as f:
            for a, b in zip(p, n):
                f.write(" ".join(repr(float(x)) for x in (*a, *b)) + "\n")
        np.savetxt(os.path.join(d, f"pose{i * stride}.txt"), sc["poses_init"][i], fmt="%.17g")
        if with_gt:
            np.savetxt(os.path.join(d, f"groundtruth{i * stride}.txt"), sc["poses_gt"][i], fmt="%.17g")


def test_driver_loads_reference_formats(tmp_path):
    """No GPU needed: file discovery (length-then-lexicographic order, --limit/--step), parsing, and the error path."""
    _build()
    sc = scene(12, 50, 3)
    _write_scene(str(tmp_path), sc)
    (tmp_path / "notes.txt").write_text("ignored: no cloud/pose prefix\n")
    r = subprocess.run([BIN, f"--dir={tmp_path}", "--limit=5", "--step=2", "--rounds=1", "--norecomputeNormals"], capture_output=True, text=True)
    assert "loaded 5 frames" in r.stdout                     # files 0,2,4,6,8 (cloud10/11 sort after cloud9)
    assert r.returncode in (0, 2)                            # 2 = no CUDA device here: reported, not a crash
    if r.returncode == 2:
        assert "mvicp:" in r.stderr
    r = subprocess.run([BIN, f"--dir={tmp_path}/missing"], capture_output=True, text=True)
    assert r.returncode == 1 and "Could not open directory" in r.stderr


@pytest.mark.gpu
def test_driver_matches_python_api(tmp_path):
    from mv_lm_icp_b200 import Frame, ICP_Ceres
    _build()
    sc = scene(5, 4000, 17)
    _write_scene(str(tmp_path), sc)
    out = tmp_path / "out"; out.mkdir()
    r = subprocess.run([BIN, f"--dir={tmp_path}", "--limit=40", "--step=1", "--knn=2", "--rounds=4", f"--out={out}"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert len(re.findall(r"=====  TIMING\[closest pts \d+\] is [0-9.e+-]+ s", r.stdout)) == 4
    assert len(re.findall(r"=====  TIMING\[global \d+\] is [0-9.e+-]+ s", r.stdout)) == 4
    got = [np.loadtxt(out / f"pose_out_{i}.txt") for i in range(5)]

    frames = [Frame(p, n, pose=P) for p, n, P in zip(sc["pts"], sc["nor"], sc["poses_init"])]
    icp = ICP_Ceres(frames)
    icp.recomputeNormals(10)
    frames[0].fixed = True
    icp.computePoseNeighbours(2)
    for _ in range(4):
        icp.computeClosestPoints(0.05)
        icp.ceresOptimizer_sophusSE3(True, True)
    for i in range(5):
        assert np.array_equal(got[i], frames[i].pose), i     # same library, same inputs (text round trip is exact at %.17g)
    icp.engine.close()


def _write_cloud(path, pts, nor):
    with open(path, "w") as f:
        for a, b in zip(pts, nor):
            f.write(" ".join(repr(float(x)) for x in (*a, *b)) + " \n")     # trailing space as in the reference's files


def test_pairwise_driver_cli(tmp_path):
    _build()
    r = subprocess.run([BIN_PAIR, f"--cloud={tmp_path}/none.xyz"], capture_output=True, text=True)
    assert r.returncode == 1 and "could not be opened" in r.stderr
    sc = scene(2, 200, 5)
    _write_cloud(tmp_path / "c.xyz", sc["pts"][0], sc["nor"][0])
    r = subprocess.run([BIN_PAIR, f"--cloud={tmp_path}/c.xyz"], capture_output=True, text=True)
    assert r.returncode in (0, 2)                            # 2 = no CUDA device here
    assert len(r.stdout.splitlines()) >= 10                  # the first ten points are echoed (main_pairwise.cpp:36-39)
    if r.returncode == 2:
        assert "mvicp:" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("p2plane", [False, True])
def test_pairwise_driver_recovers_known_transform(tmp_path, p2plane):
    """main_pairwise.cpp's known-answer benchmark (README.md:141-150: ~1e-10 translation, ~1e-6 degrees)."""
    _build()
    sc = scene(2, 3000, 9)
    _write_cloud(tmp_path / "c.xyz", sc["pts"][0], sc["nor"][0])
    args = [BIN_PAIR, f"--cloud={tmp_path}/c.xyz", f"--out={tmp_path}"] + (["--pointToPlane"] if p2plane else [])
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "=====  TIMINGS ====" in r.stdout and "=====  Accurracy ====" in r.stdout
    assert len(re.findall(r"=====  TIMING\[ceres (CeresAngleAxis|EigenQuaternion|SophusSE3)\] is", r.stdout)) == 3
    P = np.loadtxt(tmp_path / "P_true.txt")
    for k in range(3):
        E = np.loadtxt(tmp_path / f"P_est_{k}.txt")
        assert np.linalg.norm(E[:3, 3] - P[:3, 3]) < 1e-8, (k, E, P)
        assert np.degrees(np.arccos(np.clip((np.trace(E[:3, :3].T @ P[:3, :3]) - 1) / 2, -1, 1))) < 1e-5
    m = re.findall(r"diff_tra:([0-9.e+-]+)\t diff_rot_degrees:([0-9.e+-]+)", r.stdout)
    # the three LM solvers; the closed-form row in front of them is checked by tests/test_gpu_zz_closed_form.py
    assert len(m) >= 3 and all(float(a) < 1e-8 and float(b) < 1e-4 for a, b in m[-3:])
