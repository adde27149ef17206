"""CPU tests of the drop-in boundary: libnhp.so loads, exports every symbol include/nhp.h declares,
and fails loudly (no CPU fallback) when no GPU is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions(names=("nhp.h", "nhp_devel.h")):
    found = set()
    for name in names:
        src = open(os.path.join(ROOT, "include", name)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        found |= set(re.findall(r"\b(nhp_[a-z0-9_]+)\s*\(", src))
    return sorted(found)


def test_boundary_header_holds_no_measurement_hooks():
    boundary = header_functions(("nhp.h",))
    for hook in ("nhp_bench_fp64", "nhp_test_fastmath", "nhp_launch_count", "nhp_cont_params_save"):
        assert hook not in boundary


def test_library_exports_every_declared_symbol():
    from nhp_b200 import _lib
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libnhp.so does not export {n}"
    assert set(names) == set(_lib.PROTOTYPES), "ctypes prototypes and include/nhp.h disagree"
    assert lib.nhp_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import nhp_b200
    with pytest.raises(nhp_b200.NHPError) as e:
        nhp_b200.Context(0)
    assert e.value.code == -3  # NHP_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "networkhawkesprocesses.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle"
