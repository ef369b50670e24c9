"""ctypes binding of libnvalchemi_nl_b200.so (the C ABI in include/nvalchemi_nl_b200.h).

There is deliberately no fallback: if the library is missing or fails to load the import error
propagates, and every op checks that its tensors live on a CUDA device.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "csrc", "libnvalchemi_nl_b200.so"))

ABI_VERSION = 6
_lib = None

c_void_p, c_int, c_int32, c_int64, c_double, c_size_t = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_size_t)

_SIGNATURES = {
    "nvnl_abi_version": (c_int, []),
    "nvnl_last_error": (ctypes.c_char_p, []),
    "nvnl_launch_count": (c_int64, []),
    "nvnl_set_rows_budget": (None, [c_int64, c_int64]),
    "nvnl_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "nvnl_build": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_double, c_int64,
                           c_void_p, c_size_t, c_void_p]),
    "nvnl_import_cache": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int32, c_double, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_size_t,
                                  c_void_p]),
    "nvnl_count": (c_int, [c_void_p, c_int, c_int64, c_int32, c_void_p, c_double, c_int, c_int, c_void_p, c_void_p,
                           c_void_p]),
    "nvnl_status": (c_int, [c_void_p, c_int, c_int64, c_int32, ctypes.POINTER(c_int64), ctypes.POINTER(c_int32),
                            ctypes.POINTER(c_int32), ctypes.POINTER(c_int32), ctypes.POINTER(c_int32),
                            ctypes.POINTER(c_int32), ctypes.POINTER(c_int32), c_void_p]),
    "nvnl_count_rows": (c_int, [c_void_p, c_int, c_int64, c_int32, c_void_p, c_double, c_int, c_int, c_void_p, c_void_p,
                                c_void_p, c_int64, c_int32, c_void_p]),
    "nvnl_fill_rows": (c_int, [c_void_p, c_int, c_int64, c_int32, c_void_p, c_double, c_int, c_int, c_void_p, c_void_p,
                               c_int64, c_int64, c_void_p, c_int32, c_int32, c_void_p]),
    "nvnl_fill_rows_speculative": (c_int, [c_void_p, c_int, c_int64, c_int32, c_void_p, c_void_p, c_int64, c_void_p, c_int32,
                                           c_void_p]),
    "nvnl_fill_coo": (c_int, [c_void_p, c_int, c_int64, c_int32, c_void_p, c_double, c_int, c_int, c_void_p, c_void_p,
                              c_int64, c_int64, c_void_p, c_int32, c_int32, c_void_p]),
    "nvnl_fill_matrix": (c_int, [c_void_p, c_int, c_int64, c_int32, c_void_p, c_double, c_int, c_int, c_void_p,
                                 c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "nvnl_export_cache": (c_int, [c_void_p, c_int, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "nvnl_refresh_positions": (c_int, [c_void_p, c_int, c_int64, c_int32, c_void_p, c_void_p]),
    "nvnl_cells_changed": (c_int, [c_void_p, c_int, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nvnl_cells_changed_cache": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                         c_void_p, c_void_p]),
    "nvnl_moved_beyond": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_double, c_void_p, c_void_p]),
    "nvnl_get_grid": (c_int, [c_void_p, c_int, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "nvnl_coulomb_fused": (c_int, [c_void_p, c_int, c_int64, c_int32, c_void_p, c_double, c_int, c_void_p, c_double, c_double,
                                   c_void_p, c_void_p, c_void_p]),
    "nvnl_coulomb_list": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int32, c_void_p, c_void_p, c_double, c_double, c_void_p,
                                  c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "nvnl_pack_shifts": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "nvnl_pack_shifts_word": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "nvnl_expand_padded": (c_int, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "nvnl_expand_padded_ranges": (c_int, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nvnl_expand_gathered": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
}


def library_path() -> str:
    return _LIB_PATH


def lib():
    """Load the CUDA library (once).  Raises if it is missing: there is no other implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(
                f"{_LIB_PATH} not found: build it with `python nvalchemi-toolkit-ops_b200/build.py` "
                "(nvcc, sm_100a).  nvalchemiops_b200 has no CPU or PyTorch fallback."
            )
        L = ctypes.CDLL(_LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the ABI lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if L.nvnl_abi_version() != ABI_VERSION:
            raise ImportError(f"ABI mismatch: library {L.nvnl_abi_version()} != binding {ABI_VERSION}")
        _lib = L
    return _lib


def declared_symbols():
    return sorted(_SIGNATURES)


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib().nvnl_last_error().decode()}")


def launch_count() -> int:
    """Kernels launched by the library since it was loaded."""
    return int(lib().nvnl_launch_count())
