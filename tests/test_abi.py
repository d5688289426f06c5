"""CPU: the C-ABI library loads, exports every symbol include/bpmf_b200.h declares, refuses to
compute without a CUDA device (no CPU fallback), and its host-side helper symbols behave like the
reference's (libc.c:389-439, 488-514)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from seismic_bpmf_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="bpmf_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    names = re.findall(r"\b([A-Za-z_]\w*)\s*\([^;{]*\)\s*;", text)
    return sorted(set(names))


def test_header_symbols_are_exported():
    declared = _declared_symbols()
    assert len(declared) >= 30, declared
    out = subprocess.check_output(["nm", "-D", "--defined-only", build.lib_path()], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [n for n in declared if n not in exported]
    assert not missing, "declared in include/bpmf_b200.h but not exported: %s" % missing
    # the ctypes table binds exactly the declared surface
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared


def test_features_header_symbols_are_exported():
    from seismic_bpmf_b200 import features

    declared = _declared_symbols("bpmf_b200_features.h")
    out = subprocess.check_output(["nm", "-D", "--defined-only", build.features_lib_path()], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert not [n for n in declared if n not in exported]
    assert sorted(features.EXPORTED_SYMBOLS) == declared
    features.load()


def test_reference_libc_symbol_set_is_covered():
    """Every symbol of the reference's libc.so (BPMF/libc.h:1-11 + find_similar_moveouts2) exists."""
    ref_syms = ["kurtosis", "find_similar_moveouts", "find_similar_moveouts2", "select_cc_indexes",
                "time_dependent_threshold", "swap", "selectionSort", "sort_indexes", "_compute_average",
                "_compute_std"]
    lib = _lib.load()
    for s in ref_syms:
        assert hasattr(lib, s)
    link = os.path.join(os.path.dirname(build.lib_path()), "libc.so")
    assert os.path.exists(link)  # usable as BPMF/lib/libc.so (clib.py:12-15)


def test_library_loads_and_reports():
    lib = _lib.load()
    assert b"sm_100a" in lib.bpmf_version()
    assert lib.bpmf_mf_n_corr(8640000, 800, 1) == 8639201
    assert lib.bpmf_mf_n_corr(1000, 100, 7) == (1000 - 100) // 7 + 1
    assert lib.bpmf_mf_n_corr(10, 100, 1) == 0
    assert lib.bpmf_mf_workspace_bytes(800, 8640000, 4, 16, 3, 1, 0, 1, _lib.MF_FP32) > 4 * 48 * 8639201
    assert lib.bpmf_beam_workspace_bytes(1000, 8, 2, 2, 5000) >= 4 * 8 * 2 * 5000


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from seismic_bpmf_b200 import clib, fast_matched_filter
    from seismic_bpmf_b200.beampower import beampower

    with pytest.raises(_lib.BpmfError):
        fast_matched_filter.matched_filter(np.zeros((1, 1, 1, 8), np.float32), np.zeros((1, 1, 1), np.int32),
                                           np.ones((1, 1, 1), np.float32), np.zeros((1, 1, 64), np.float32), 1)
    with pytest.raises(_lib.BpmfError):
        beampower.beamform(np.zeros((2, 1, 64), np.float32), np.zeros((3, 2, 1), np.int64),
                           np.ones((2, 1, 1), np.float32), np.ones((3, 2), np.float32))
    with pytest.raises(_lib.BpmfError):
        clib.time_dependent_threshold(np.zeros(1000, np.float32), 100, 8.0)
    # the raw C entry point reports the CUDA failure instead of computing on the host
    lib = _lib.load()
    out = np.zeros((1, 57), np.float32)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    rc = lib.bpmf_matched_filter(np.zeros(8, np.float32).ctypes.data_as(fp), np.zeros(1, np.int32).ctypes.data_as(ip),
                                 np.ones(1, np.float32).ctypes.data_as(fp), np.zeros(64, np.float32).ctypes.data_as(fp),
                                 1, 8, 64, 1, 1, 1, 0, 1, 0, out.ctypes.data_as(fp))
    assert rc == -1 and len(lib.bpmf_last_error()) > 0


def test_argument_validation_is_device_independent():
    from seismic_bpmf_b200 import fast_matched_filter
    from seismic_bpmf_b200.beampower import beampower

    for bad in [dict(templates=np.zeros((1, 2, 1, 8), np.float32))]:
        with pytest.raises((ValueError, _lib.BpmfError)):
            fast_matched_filter.matched_filter(bad["templates"], np.zeros((1, 1, 1), np.int32),
                                               np.ones((1, 1, 1), np.float32), np.zeros((1, 1, 64), np.float32), 1)
    with pytest.raises((ValueError, _lib.BpmfError)):
        beampower.beamform(np.zeros((2, 1, 64), np.float32), np.zeros((3, 2, 1), np.int64),
                           np.ones((2, 1, 1), np.float32), np.ones((3, 2), np.float32), reduce="sum")


def test_host_helper_symbols():
    lib = _lib.load()
    fp = C.POINTER(C.c_float)
    rng = np.random.default_rng(0)
    a = rng.standard_normal(17).astype(np.float32)
    b = a.copy()
    lib.selectionSort(b.ctypes.data_as(fp), b.size)
    assert np.array_equal(b, np.sort(a))
    idx = np.zeros(a.size, dtype=np.uintp)
    lib.sort_indexes(a.ctypes.data_as(fp), idx.ctypes.data_as(C.POINTER(C.c_size_t)), a.size)
    assert np.array_equal(a[idx], np.sort(a))
    x = np.array([1.0, 2.0], dtype=np.float32)
    lib.swap(x[0:].ctypes.data_as(fp), x[1:].ctypes.data_as(fp))
    assert x.tolist() == [2.0, 1.0]
    m = lib._compute_average(a.ctypes.data_as(fp), a.size)
    assert abs(m - a.mean()) < 1e-6
    assert abs(lib._compute_std(a.ctypes.data_as(fp), m, a.size) - a.std()) < 1e-6


def test_reference_clib_py_loads_against_this_library():
    """The reference's own BPMF/clib.py (unchanged) binds every symbol it needs from libbpmf_b200.so
    installed as lib/libc.so (only where /root/reference exists; compute calls need a GPU)."""
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("/root/reference is not present on this box")
    pkg_dir = os.path.dirname(os.path.dirname(build.lib_path()))  # .../seismic_bpmf_b200 holds lib/libc.so
    clib_ref = ref_loader.load_reference_clib(package_dir=pkg_dir)
    assert clib_ref.cpu_loaded
    for sym in ("kurtosis", "find_similar_moveouts", "find_similar_moveouts2", "select_cc_indexes",
                "time_dependent_threshold"):
        assert getattr(clib_ref._libc, sym).argtypes is not None
