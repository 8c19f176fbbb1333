"""ctypes binding of libagpknn.so (include/agpknn.h).

The library is the product: there is no Python/CPU fallback.  If the shared object is missing
this module raises at first use with the build command to run; if no sm_100 GPU is visible the
C entry points return AGP_ENODEV and the wrappers raise ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_void_p
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libagpknn.so"

MEM_HOST, MEM_DEVICE = 0, 1
PRECISION = {"auto": 0, "fp32_simt": 1, "3xtf32": 2, "exact_diff": 3, "3xfp16": 4, "fp16_screen": 5}
MAX_K = 512

# every symbol include/agpknn.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "agp_index_create": (c_int, [c_int, c_int, c_int, POINTER(c_void_p)]),
    "agp_index_create_metric": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    "agp_index_create_multi": (c_int, [c_int, c_int, POINTER(c_int), c_int, c_int, POINTER(c_void_p)]),
    "agp_index_n_shards": (c_int, [c_void_p]),
    "agp_index_metric": (c_int, [c_void_p]),
    "agp_index_free": (None, [c_void_p]),
    "agp_index_add": (c_int, [c_void_p, c_int64, c_void_p, c_int]),
    "agp_index_search": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int]),
    "agp_index_reset": (c_int, [c_void_p]),
    "agp_index_ntotal": (c_int64, [c_void_p]),
    "agp_index_dim": (c_int, [c_void_p]),
    "agp_index_reserve": (c_int, [c_void_p, c_int64]),
    "agp_index_set_stream": (c_int, [c_void_p, c_void_p, c_int]),
    "agp_index_set_id_base": (c_int, [c_void_p, c_int64]),
    "agp_index_set_profiling": (c_int, [c_void_p, c_int]),
    "agp_index_get_profile": (c_int, [c_void_p, POINTER(c_double), POINTER(c_int64), c_int]),
    "agp_index_get_profile_phases": (c_int, [c_void_p, POINTER(c_double), POINTER(c_int64), c_int]),
    "agp_index_get_stats": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64)]),
    "agp_index_set_knob": (c_int, [c_void_p, c_char_p, c_int]),
    "agp_index_screen_probe": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "agp_index_search_masked": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "agp_index_search_subset": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "agp_best_of_lists": (c_int, [c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "agp_merge_topk": (c_int, [c_int, c_void_p, c_int64, c_int, c_int, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "agp_merge_topk_metric": (c_int, [c_int, c_void_p, c_int64, c_int, c_int, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "agp_recall_at_n": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "agp_radius_count": (c_int, [c_int, c_int64, c_int, c_void_p, c_int64, c_void_p, c_double, c_void_p]),
    "agp_radius_fill": (c_int, [c_int, c_int64, c_int, c_void_p, c_int64, c_void_p, c_double, c_void_p, c_void_p]),
    "agp_plan_screen": (c_int, [c_int64, c_int64, c_int, c_int, c_int64, c_int, POINTER(c_int)]),
    "agp_plan_screen_piece": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_int)]),
    "agp_plan_host_chunks": (c_int, [c_int64, c_int, c_int, c_int64, c_int, c_int, c_int, POINTER(c_int64), c_int]),
    "agp_last_error": (c_char_p, []),
    "agp_device_count": (c_int, []),
    "agp_kernel_launches": (c_int64, []),
    "agp_version": (c_char_p, []),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load libagpknn.so and bind every declared symbol (raises if the build is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build the CUDA extension with `python -m agplace_b200.build` "
            "(agplace_b200 has no CPU fallback)")
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().agp_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def kernel_launches() -> int:
    return int(load().agp_kernel_launches())


def device_count() -> int:
    return int(load().agp_device_count())
