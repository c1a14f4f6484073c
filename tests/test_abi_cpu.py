"""No-GPU checks of the drop-in boundary: libweedcu.so loads and exports every symbol that
include/weedcu.h declares; argument validation paths return error codes instead of crashing."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from weed_b200 import weedcu
    from weed_b200._lib import SYMBOLS
    lib = weedcu()
    hdr = open(os.path.join(ROOT, "include", "weedcu.h")).read()
    declared = set(re.findall(r"\b(weedcu_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(SYMBOLS)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"declared in weedcu.h but not exported: {missing}"


def test_view_struct_layout_matches_header():
    from weed_b200 import View, Mat
    assert C.sizeof(View) == 8 + 4 + 4 + 8 * 4 + 8 * 4  # offset, rank(+pad), shape[8], stride[8]
    assert C.sizeof(Mat) == 24


def test_bad_arguments_return_codes_not_crashes():
    from weed_b200 import weedcu
    lib = weedcu()
    assert lib.weedcu_fill_real(None, C.c_uint64(4), C.c_float(0), None) == -1
    assert lib.weedcu_device_count(None) == -1
    assert b"invalid" in lib.weedcu_error_string(-1)


def test_missing_extension_fails_loudly(monkeypatch, tmp_path):
    import weed_b200._lib as L
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "_HERE", str(tmp_path))
    with pytest.raises(L.WeedcuError):
        L.weedcu()
