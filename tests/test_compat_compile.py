"""The C++ drop-in shim (compat/mvicp_compat.hpp) compiles against the reference's container shapes and links with
libmvicp.so (no GPU needed: nothing is executed)."""
import os
import subprocess

import mv_lm_icp_b200 as mv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_compat_shim_compiles_and_links(tmp_path):
    so = mv.build()
    obj, exe = str(tmp_path / "cc.o"), str(tmp_path / "cc")
    subprocess.run(["/usr/bin/g++", "-std=c++11", "-Wall", "-I" + os.path.join(ROOT, "compat", "eigen_stub"), "-c",
                    os.path.join(ROOT, "compat", "compile_check.cpp"), "-o", obj], check=True)
    subprocess.run(["/usr/bin/g++", obj, "-L" + os.path.dirname(so), "-lmvicp", "-Wl,-rpath," + os.path.dirname(so), "-o", exe], check=True)
    assert subprocess.run([exe]).returncode == 0
