"""libmgcn.so builds (cross-compiled for sm_100a without a GPU), loads, and exports every symbol
include/mgcn.h declares; argument validation works without touching the device."""
import ctypes
import os
import re
import subprocess

from meta_gcn_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mgcn.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mgcn_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in mgcn.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_sass_is_sm_100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", build.LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_version_and_error_strings():
    lib = _lib.load()
    assert lib.mgcn_version() == 100
    assert lib.mgcn_error_string(0) == b"ok"
    assert b"NULL" in lib.mgcn_error_string(-1)


def test_argument_errors_without_device():
    lib = _lib.load()
    n = ctypes.c_size_t(0)
    caps = [ctypes.c_int64(0) for _ in range(3)]
    assert lib.mgcn_csr_capacities(1000, 100, 2, 256, *[ctypes.byref(c) for c in caps]) == 0
    assert caps[0].value == 1100 and caps[1].value == 1100 // 256 + 1
    # workspace query never launches
    rc = lib.mgcn_csr_build(None, 1000, 100, 1, 0, None, None, None, ctypes.byref(n), None)
    assert rc == 0 and n.value > 4 * 1000 * 4
    assert lib.mgcn_csr_build(None, 10, 10, 7, 0, None, None, None, ctypes.byref(n), None) == -3
    assert lib.mgcn_csr_build(None, 1 << 31, 10, 0, 0, None, None, None, ctypes.byref(n), None) == -2
    assert lib.mgcn_spmm(None, None, 0, 32, 0, None, None, None, 0, None, None, 0, None, None,
                         ctypes.byref(n), None) == -1
    s = _lib.MgcnCsr()
    s.n_rows = 4
    assert lib.mgcn_spmm(ctypes.byref(s), None, 4, 0, 0, None, None, None, 0, None, None, 0, None,
                         None, ctypes.byref(n), None) == -3         # H = 0
    assert lib.mgcn_aggregate_prescaled(ctypes.byref(s), None, 4, 48, None, 0, None, None, 0, None,
                                        None, ctypes.byref(n), None) == -3   # H not in {16,32,64,128}
    assert lib.mgcn_linear(None, 5, 0, None, 1, 1, 4, None, None, 0, None, None) == -3
    assert lib.mgcn_linear(None, 5, 4, None, 1, 1, 4, None, None, 0, None, None) == -1
    assert lib.mgcn_gcn_norm(None, 5, 3, None, None) == -3
    assert lib.mgcn_segment_reduce(None, 8, None, 3, 100, 5, None, None, ctypes.byref(n), None) == -3
    rc = lib.mgcn_linear_wgrad(None, 1000, 32, None, 32, None, 32, 1, None, None, ctypes.byref(n), None)
    assert rc == 0 and n.value >= 32 * 32 * 4


def test_python_check_raises():
    import pytest
    with pytest.raises(RuntimeError, match="mgcn"):
        _lib.check(-3)
