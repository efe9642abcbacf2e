"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/erp_b200.h declares, refuses to run without a device, and its host-side logic
(random_array replay) matches the oracle.  No compute calls are made here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import erp_match_eightpoint_test_b200 as erp
import oracle as O
from erp_match_eightpoint_test_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "erp_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(erp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = erp.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in erp_b200.h but not exported"
    assert set(names) == set(binding._SIGNATURES), "binding.py and erp_b200.h disagree"
    assert lib.erp_version() == 200


def test_exports_are_plain_c_and_torch_free():
    out = subprocess.run(["nm", "-D", "--defined-only", erp.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert all(not s.startswith("_Z") for s in exported), "C++-mangled symbols leak from the C ABI"
    deps = subprocess.run(["ldd", erp.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in deps and "c10" not in deps


def test_dmatch_layout_is_cv_dmatch():
    assert binding.DMATCH.itemsize == 16
    assert [binding.DMATCH.fields[k][1] for k in ("queryIdx", "trainIdx", "imgIdx", "distance")] == [0, 4, 8, 12]
    assert ctypes.sizeof(binding.RansacResult) == 8 + 8 + 4 + 4 + 72 + 72 + 48


def test_no_device_is_an_error_not_a_fallback():
    if erp.lib().erp_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(erp.ErpError) as ei:
        erp.Context(0)
    assert ei.value.status == binding.E_NO_DEVICE and "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "erp_match_eightpoint_test_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle|#include\s+[\"<].*erp_oracle", text, re.M), f
    host = os.path.join(ROOT, "host")
    for dirpath, dirs, files in os.walk(host):
        dirs[:] = [d for d in dirs if d != "_build"]
        for f in files:
            text = open(os.path.join(dirpath, f), errors="ignore").read()
            assert "erp_oracle" not in text and not re.search(r"^\s*(import|from)\s+oracle", text, re.M), f


@pytest.mark.parametrize("m,H,S", [(37, 3, 37), (200, 80, 50), (1000, 80, 250)])
def test_random_array_replay_matches_glibc(m, H, S):
    """erp_libstdcxx_sample_table restates glibc rand() + libstdc++ random_shuffle; the oracle
    calls the real rand().  (src/eight_point.hpp:54-58, src/eight_point.cpp:99-111)"""
    assert np.array_equal(erp.libstdcxx_sample_table(m, H, S, 1), O.ref_sample_table(m, H, S, 1))


def test_sample_table_argument_errors():
    with pytest.raises(erp.ErpError) as ei:
        erp.libstdcxx_sample_table(10, 2, 11)
    assert ei.value.status == binding.E_ARG


def test_shard_range_is_the_python_helper():
    """erp_shard_range (what the library cuts query rows and hypothesis ids with) == sharding.shard_range (what the gloo
    tests of the host logic use): contiguous, ordered, sizes differing by at most one."""
    from erp_match_eightpoint_test_b200 import sharding
    for n in (0, 1, 7, 100, 100003, 1000000):
        for w in (1, 2, 3, 8, 64):
            parts = [erp.shard_range(n, r, w) for r in range(w)]
            assert parts == [sharding.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    with pytest.raises(erp.ErpError):
        erp.shard_range(10, 3, 3)


def test_multi_gpu_entry_points_fail_loudly_without_devices():
    if erp.lib().erp_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(erp.ErpError) as ei:
        erp.Group([0, 1])
    assert ei.value.status in (binding.E_NO_DEVICE, binding.E_NCCL)
