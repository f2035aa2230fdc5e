"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/ef_track.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import instancefusion_b200 as ef
from instancefusion_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "ef_track.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ef_[a-z0-9_]+)\s*\(", text)))


def test_library_exists_and_loads():
    assert os.path.exists(binding.lib_path()), "build with __graft_entry__.build()"
    L = binding.lib()
    assert L.ef_abi_version() == 1


def test_every_declared_symbol_is_exported():
    L = binding.lib()
    names = _declared()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(binding.EXPORTED) == names


def test_no_cpp_or_torch_types_in_signatures():
    text = open(os.path.join(ROOT, "include", "ef_track.h")).read()
    for bad in ("std::", "torch", "at::", "Eigen", "template"):
        assert bad not in re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def test_product_does_not_link_oracle():
    import subprocess
    out = subprocess.run(["ldd", binding.lib_path()], capture_output=True, text=True).stdout
    assert "ef_oracle" not in out and "ef_ref" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", binding.lib_path()], capture_output=True, text=True).stdout
    assert "efo_" not in syms and "efr_" not in syms


def test_product_sources_never_reference_oracle():
    pkg = os.path.join(ROOT, "instancefusion_b200")
    for dp, _, files in os.walk(pkg):
        if "build" in dp:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "ef_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


def test_fails_loudly_without_gpu():
    L = binding.lib()
    if L.ef_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ef.EFError) as e:
        ef.RGBDOdometry(640, 480, 320, 240, 528, 528)
    assert e.value.code == -2
    # Tier 2 also refuses
    rc = L.ef_op_pyr_down_u16(None, C.c_size_t(0), 480, 640, None, C.c_size_t(0), None)
    assert rc == -2


def test_invalid_arguments():
    L = binding.lib()
    h = C.c_void_p()
    assert L.ef_tracker_create(0, 480, C.c_float(1), C.c_float(1), C.c_float(1), C.c_float(1), C.c_float(0.1), C.c_float(0.3), None, C.byref(h)) == -1
    assert L.ef_tracker_create(8, 480, C.c_float(1), C.c_float(1), C.c_float(1), C.c_float(1), C.c_float(0.1), C.c_float(0.3), None, C.byref(h)) == -1
    # sizes that are not multiples of 4 are accepted like the reference accepts them (the check that follows is the device check)
    if L.ef_device_count() == 0:
        assert L.ef_tracker_create(642, 481, C.c_float(1), C.c_float(1), C.c_float(1), C.c_float(1), C.c_float(0.1), C.c_float(0.3), None, C.byref(h)) == -2
    assert L.ef_tracker_destroy(None) == 0
    assert abs(L.ef_default_dist_thresh() - 0.10) < 1e-7
    assert abs(L.ef_default_angle_thresh() - np.sin(np.float32(20.0) * np.float32(3.14159254) / np.float32(180.0))) < 1e-6
