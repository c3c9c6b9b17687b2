"""A few seeds of tools/fuzz_hostemu.py in the CPU suite: the engine's own kernels (host model, tools/hostemu) against the
oracle on random cloud shapes (surfaces, quantised grids with exact ties, collinear / identical points, tiny clouds), scales,
offsets, rigid and non-rigid poses, cutoffs and NN schedules; the LM step on random well-posed scenes; and the converged-round shortcuts (guessed median select,
certified matches) bit for bit against an engine that has them switched off.  The tool itself
has been run over several hundred seeds."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fuzz_seeds():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_hostemu.py"), "12", "2000"], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    assert r.stdout.count(" ok ") >= 12 and r.stdout.count("lm ok") >= 3 and r.stdout.count("steady ok") >= 6
