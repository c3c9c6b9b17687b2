"""The LM step kernel's blocked envelope Cholesky (csrc/lm_step.cuh:chol_solve), executed on the host: the function's text
is compiled with 512 std::threads standing in for the CTA and std::barrier for __syncthreads/__syncwarp, on a ring-graph
normal matrix that is NaN outside the row profiles (tools/chol_host_check.py)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_blocked_cholesky_text_solves_ring_system_on_host():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "chol_host_check.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok 1" in r.stdout
