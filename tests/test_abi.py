"""The C-ABI library loads and exports every symbol include/mvicp.h declares; no compute call needs a GPU here."""
import ctypes as C
import os
import re

import pytest

import mv_lm_icp_b200 as mv
from mv_lm_icp_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "mvicp.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(mvicp_[a-z0-9_]+)\s*\(", h)))


def test_every_declared_symbol_is_exported():
    lib = mv.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.EXPORTS) == names
    assert lib.mvicp_abi_version() == 1


def test_default_options_are_ceres_defaults():
    o = mv.LmOptions()
    mv.lib().mvicp_default_lm_options(C.byref(o))
    assert o.max_num_iterations == 50 and o.max_num_consecutive_invalid_steps == 5 and o.jacobi_scaling == 1
    assert (o.initial_trust_region_radius, o.max_trust_region_radius, o.min_trust_region_radius) == (1e4, 1e16, 1e-32)
    assert (o.min_relative_decrease, o.min_lm_diagonal, o.max_lm_diagonal) == (1e-3, 1e-6, 1e32)
    assert (o.function_tolerance, o.gradient_tolerance, o.parameter_tolerance) == (1e-6, 1e-10, 1e-8)


def test_no_cpu_fallback_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: creation succeeds")
    with pytest.raises(mv.MvicpError) as ei:
        mv.Engine()
    assert ei.value.code == 2 and "no CPU path" in str(ei.value)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "mv_lm_icp_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle/" not in txt and "import oracle" not in txt and "from oracle" not in txt, f
                assert "liboracle" not in txt, f


def test_invalid_arguments_are_rejected_without_a_device():
    lib = mv.lib()
    assert lib.mvicp_create(None, None) != 0
    assert lib.mvicp_set_graph(None, 0, None, None) != 0
    assert lib.mvicp_nccl_unique_id(None) != 0
