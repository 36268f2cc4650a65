"""ctypes binding of libapnetg.so (include/ap_netg.h, include/ap_cond.h, include/ap_flow.h). No fallback: a missing or unloadable library raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
# AP_NETG_LIB selects another BUILD of the same library (A/B timing of kernel variants); there is still no fallback
LIB_PATH = os.environ.get("AP_NETG_LIB") or os.path.join(_PKG, "lib", "libapnetg.so")

AP_PREC_FP32X3 = 0
AP_PREC_BF16 = 1
AP_PREC_FP32_SIMT = 2
PRECISIONS = {"fp32": AP_PREC_FP32X3, "fp32x3": AP_PREC_FP32X3, "bf16": AP_PREC_BF16, "fp32_simt": AP_PREC_FP32_SIMT}

# every symbol include/*.h declares: (restype, argtypes)
_FP = C.POINTER(C.c_float)
SYMBOLS = {
    "ap_netg_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]),
    "ap_netg_destroy": (C.c_int, [C.c_void_p]),
    "ap_netg_load_weights": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_int64), C.c_int, C.c_void_p]),
    "ap_netg_workspace_bytes": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]),
    "ap_netg_forward": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [C.c_void_p]),
    "ap_netg_forward_host": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [C.c_void_p]),
    "ap_netg_forward_host_async": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [C.c_void_p]),
    "ap_netg_host_sync": (C.c_int, [C.c_void_p]),
    "ap_netg_forward_shared_photo": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [C.c_void_p]),
    "ap_netg_compose": (C.c_int, [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 7),
    "ap_netg_last_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "ap_netg_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "ap_netg_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "ap_device_enable_peer_access": (C.c_int, [C.c_int, C.c_int]),
    "ap_peer_alloc": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]),
    "ap_peer_open": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]),
    "ap_peer_close": (C.c_int, [C.c_void_p]),
    "ap_peer_free": (C.c_int, [C.c_void_p]),
    "ap_netg_get_profile": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                      C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "ap_netg_debug_read": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int64), C.c_void_p]),
    "ap_conv2d_debug": (C.c_int, [C.c_int] * 12 + [C.c_void_p] * 5),
    # include/ap_cond.h
    "ap_cond_draw_landmarks": (C.c_int, [C.c_int] * 5 + [C.c_void_p] * 3),
    "ap_cond_motion256": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                    C.c_void_p, C.c_void_p]),
    "ap_cond_motion256_workspace_bytes": (C.c_int, [C.c_int, C.POINTER(C.c_size_t)]),
    "ap_cond_kp_to_map": (C.c_int, [C.c_int] * 4 + [C.c_float] + [C.c_void_p] * 3),
    "ap_cond_matte_photo": (C.c_int, [C.c_int] * 4 + [C.c_void_p] * 5),
    # include/ap_flow.h
    "ap_flow_create": (C.c_int, [C.POINTER(C.c_void_p)] + [C.c_int] * 8),
    "ap_flow_destroy": (C.c_int, [C.c_void_p]),
    "ap_flow_load_weights": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_int64), C.c_int, C.c_void_p]),
    "ap_flow_forward": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_void_p]),
    "ap_flow_warp_landmarks": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 4),
    "ap_flow_last_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "ap_last_error": (C.c_char_p, []),
    "ap_version": (C.c_char_p, []),
}

_lib = None
_lock = threading.Lock()


class ApNetgError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load libapnetg.so once. Raises ApNetgError when it has not been built (python -m animateportrait_b200.build)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ApNetgError(f"{LIB_PATH} not found: build it with `python -m animateportrait_b200.build` "
                                  "(there is no CPU or PyTorch fallback for the generator)")
            l = C.CDLL(LIB_PATH)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(l, name)  # AttributeError if the symbol is not exported
                fn.restype = res
                fn.argtypes = args
            _lib = l
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().ap_last_error()
        raise ApNetgError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
