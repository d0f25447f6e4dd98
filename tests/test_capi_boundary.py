"""The drop-in boundary: libdlsc_b200.so loads and exports every symbol include/dlsc_b200.h declares
(no compute calls: this tier has no GPU), and the product refuses to run without CUDA."""
import ctypes
import os
import re

import pytest

import _parity
from dlsc_gc_planner_b200 import capi


def header_symbols():
    src = open(os.path.join(_parity.ROOT, "include", "dlsc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dlsc_[a-z_0-9]+)\s*\(", src)))


def test_header_lists_match_python_binding():
    assert sorted(capi.EXPORTS) == header_symbols()


def test_library_builds_and_exports_every_symbol():
    capi.build_library()
    lib = ctypes.CDLL(capi.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.dlsc_cuda_build() > 0
    assert lib.dlsc_abi_version() == 4


def test_loader_refuses_a_cpu_stand_in(hostsim, monkeypatch):
    """DLSC_B200_LIB may select another CUDA build, never the test-only host simulator."""
    assert not hasattr(hostsim, "dlsc_cuda_build")
    monkeypatch.setattr(capi, "LIB_PATH", _parity.HOSTSIM_SO)
    monkeypatch.setattr(capi, "_LIB", None)
    with pytest.raises(RuntimeError) as e:
        capi.load_library()
    assert "not a CUDA build" in str(e.value)


def test_no_cpu_fallback():
    """Without a CUDA device dlsc_create must fail loudly (never route through a CPU path)."""
    lib = capi.load_library()
    if lib.dlsc_device_count() > 0:
        pytest.skip("a GPU is visible; the refusal path is exercised on the CPU tier")
    cfg, m = _parity.load_case("empty10")
    with pytest.raises(capi.DlscError) as e:
        capi.SwarmPlanner(cfg, m, lib=lib)
    assert "no CUDA device" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(_parity.ROOT, "dlsc_gc_planner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_py" not in txt and "liboracle" not in txt and "hostsim" not in txt.replace(
                    "tests/hostsim", "").replace("host simulator", ""), (f,)
