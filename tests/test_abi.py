"""CPU: the C-ABI shared library builds, loads and exports every symbol include/gstex_b200.h declares
(no compute calls here - there is no GPU in the build container)."""
import ctypes
import os
import re

from gstex_cuda_b200 import _lib
from gstex_cuda_b200.build import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gstex_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gstex_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    path = build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gstex_b200.h but not exported"
    # every declared entry point has a ctypes prototype in the loader and vice versa
    assert set(names) == set(_lib.exported_symbols())


def test_host_only_entry_points():
    lib = _lib.load()
    assert lib.gstex_abi_version() == 1
    assert lib.gstex_scan_temp_bytes(1 << 20) >= 4 * ((1 << 20) // 2048)
    assert lib.gstex_sort_temp_bytes(1 << 20) >= 12 * (1 << 20)
    n, x = 1000, 16000
    assert lib.gstex_texture_forward_temp_bytes(n, x, 3, 5 * n) >= 128 * n + 16 * x + 32 * 5 * n
    assert lib.gstex_texture_forward_temp_bytes(n, x, 5, 0) >= 128 * n
    assert lib.gstex_texture_backward_temp_bytes(n, x, 3) >= 128 * n + 16 * x
    assert lib.gstex_bin_tiles_temp_bytes(8160, 4 * n) >= 8 * 4 * n + 12 * 8160
    # argument validation happens on the host before any CUDA call: error code + message, no exception
    rc = lib.gstex_sh_forward(10, 7, 0, None, None, None, None)
    assert rc == -1 and "degree" in _lib.last_error()
    rc = lib.gstex_sort_pairs(10, None, None, None, None, 0, None, None, 0, None)
    assert rc == -1 and "end_bit" in _lib.last_error()


def test_product_does_not_touch_the_oracle():
    """The shipped package must never import / call anything under oracle/."""
    pkg = os.path.join(ROOT, "gstex_cuda_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt and "gstex_oracle" not in txt, f
