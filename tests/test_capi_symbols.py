"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = "".join(open(os.path.join(ROOT, "include", h)).read() for h in sorted(os.listdir(os.path.join(ROOT, "include")))
                  if h.endswith(".h"))
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ap_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_documented_entry_points():
    d = _declared()
    for name in ["ap_netg_create", "ap_netg_destroy", "ap_netg_load_weights", "ap_netg_workspace_bytes",
                 "ap_netg_forward", "ap_netg_forward_host", "ap_netg_debug_read", "ap_last_error", "ap_version",
                 "ap_cond_draw_landmarks", "ap_cond_motion256", "ap_cond_motion256_workspace_bytes", "ap_cond_kp_to_map",
                 "ap_cond_matte_photo"]:
        assert name in d


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/ap_netg.h but not exported"
    lib.ap_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.ap_version()


def test_python_binding_covers_the_header(built_lib):
    from animateportrait_b200 import _capi
    assert sorted(_capi.SYMBOLS) == _declared()
    _capi.lib()  # sets restype/argtypes on every symbol; raises if one is missing


def test_sass_is_blackwell_native(built_lib):
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip("cuobjdump not available")
    assert "UTCHMMA" in sass, "no tcgen05.mma in the library"
    assert "UTMALDG" in sass, "no TMA loads in the library"
    assert "LDTM" in sass, "no tcgen05.ld in the library"
    assert "HMMA.16816" not in sass and "HGMMA" not in sass  # no legacy mma.sync / wgmma paths


def test_invalid_arguments_return_error_codes_without_a_gpu(built_lib):
    from animateportrait_b200 import _capi
    lib = _capi.lib()
    h = ctypes.c_void_p()
    assert lib.ap_netg_create(ctypes.byref(h), 2, 0, 0) == -3  # AP_ERR_UNSUPPORTED: output_nc must be 1 or 3
    assert b"output_nc" in lib.ap_last_error()
    assert lib.ap_netg_create(ctypes.byref(h), 1, 9, 0) == -1  # AP_ERR_INVALID
    assert lib.ap_netg_forward(None, 1, None, None, None, None, None, None, None, None) == -1
    # conditioning producers (include/ap_cond.h): argument checks come before any CUDA call
    assert lib.ap_cond_draw_landmarks(0, 1, 68, 255, 3, None, None, None) == -1
    assert lib.ap_cond_draw_landmarks(0, 1, 68, 256, 16, ctypes.c_void_p(16), ctypes.c_void_p(16), None) == -1
    assert b"radius" in lib.ap_last_error()
    assert lib.ap_cond_motion256(0, 1, ctypes.c_void_p(16), 0, ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 8,
                                 None, None) == -1
    assert b"workspace" in lib.ap_last_error()
    n = ctypes.c_size_t()
    assert lib.ap_cond_motion256_workspace_bytes(3, ctypes.byref(n)) == 0 and n.value >= 3 * 512 * 16
    assert lib.ap_cond_kp_to_map(0, 1, 68, 222, 4.0, ctypes.c_void_p(16), ctypes.c_void_p(16), None) == -1
    assert lib.ap_cond_matte_photo(0, 1, 3, 64, None, ctypes.c_void_p(16), None, None, None) == -1
