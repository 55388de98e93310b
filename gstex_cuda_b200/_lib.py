"""ctypes loader for libgstex_b200.so (the C ABI declared in include/gstex_b200.h).

There is no fallback: if the shared library is missing it is built in-tree with nvcc, and if that is
impossible the import of any op fails loudly with the build error.  Nothing here touches the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgstex_b200.so")

_lock = threading.Lock()
_lib = None

c_fp = C.c_void_p  # device pointers travel as integers
c_i = C.c_int
c_i64 = C.c_int64
c_f = C.c_float
c_sz = C.c_size_t

_PROTOS = {
    "gstex_last_error": (C.c_char_p, []),
    "gstex_abi_version": (c_i, []),
    "gstex_get_aabb_2d": (c_i, [c_i, c_fp, c_fp, c_f, c_fp, c_fp, c_f, c_f, c_f, c_f, c_fp, c_fp, c_fp]),
    "gstex_num_tiles_hit_2d": (c_i, [c_i, c_fp, c_fp, c_i, c_i, c_i, c_fp, c_fp]),
    "gstex_project_points": (c_i, [c_i, c_fp, c_fp, c_f, c_f, c_f, c_f, c_fp, c_fp, c_fp]),
    "gstex_project_aabb_count": (c_i, [c_i, c_fp, c_fp, c_f, c_fp, c_fp, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_fp,
                                       c_fp, c_fp, c_fp, c_fp, c_fp]),
    "gstex_scan_temp_bytes": (c_sz, [c_i]),
    "gstex_cumsum_i32": (c_i, [c_i, c_fp, c_fp, c_fp, c_sz, c_fp]),
    "gstex_map_gaussian_to_intersects": (c_i, [c_i, c_i64, c_fp, c_fp, c_fp, c_fp, c_i, c_i, c_i, c_fp, c_fp, c_fp]),
    "gstex_map_gaussian_to_intersects_wrapped": (c_i, [c_i, c_i64, c_fp, c_fp, c_fp, c_fp, c_i, c_i, c_i, c_fp, c_fp, c_fp]),
    "gstex_sort_temp_bytes": (c_sz, [c_i64]),
    "gstex_sort_pairs": (c_i, [c_i64, c_fp, c_fp, c_fp, c_fp, c_i, c_fp, c_fp, c_sz, c_fp]),
    "gstex_get_tile_bin_edges": (c_i, [c_i64, c_fp, c_fp, c_fp, c_fp]),
    "gstex_bin_tiles_temp_bytes": (c_sz, [c_i, c_i64]),
    "gstex_bin_tiles": (c_i, [c_i, c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i64, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_sz, c_fp]),
    "gstex_texture_forward_temp_bytes": (c_sz, [c_i, c_i64, c_i, c_i64]),
    "gstex_texture_backward_temp_bytes": (c_sz, [c_i, c_i64, c_i]),
    "gstex_texture_backward_stateless_temp_bytes": (c_sz, [c_i, c_i64, c_i, c_i64]),
    "gstex_texture_forward": (c_i, [c_i, c_i, c_i, c_i, c_i64, c_i, c_i64] + [c_fp] * 7 + [c_f] + [c_fp] * 7 + [c_f] * 4
                              + [c_i] + [c_fp] * 10 + [c_fp, c_sz, c_fp]),
    "gstex_texture_backward": (c_i, [c_i, c_i, c_i, c_i, c_i64, c_i, c_i64] + [c_fp] * 7 + [c_f] + [c_fp] * 7 + [c_f] * 4
                               + [c_i] + [c_fp] * 11 + [c_fp] * 9 + [c_i, c_fp, c_fp, c_sz, c_fp]),
    "gstex_texture_edit_temp_bytes": (c_sz, [c_i]),
    "gstex_texture_edit": (c_i, [c_i, c_i, c_i, c_i, c_i64, c_i] + [c_fp] * 10 + [c_f] + [c_fp] * 6 + [c_f] * 4 + [c_i]
                           + [c_fp, c_fp, c_sz, c_fp]),
    "gstex_sh_forward": (c_i, [c_i, c_i, c_i, c_fp, c_fp, c_fp, c_fp]),
    "gstex_sh_backward": (c_i, [c_i, c_i, c_i, c_fp, c_fp, c_fp, c_i, c_fp]),
    "gstex_texture_sample_forward": (c_i, [c_i, c_i, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "gstex_texture_sample_backward": (c_i, [c_i, c_i, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "gstex_image_loss": (c_i, [c_i, c_i] + [c_fp] * 11 + [c_fp]),
    "gstex_pad_texture": (c_i, [c_i64, c_fp, c_fp, c_fp]),
    "gstex_unpad_texture_grad": (c_i, [c_i64, c_fp, c_fp, c_i, c_fp]),
    "gstex_pack_records": (c_i, [c_i] + [c_fp] * 5 + [c_f] + [c_fp] * 6 + [c_f] * 4 + [c_fp, c_fp, c_fp, c_fp]),
    "gstex_fill_zero": (c_i, [c_fp, c_sz, c_fp]),
    "gstex_raster_forward": (c_i, [c_i] * 5 + [c_fp] * 7 + [c_f] * 4 + [c_fp] * 10 + [c_fp, c_i64, c_fp] + [c_fp]),
    "gstex_raster_masks": (c_i, [c_i] * 4 + [c_fp] * 6 + [c_f] * 4 + [c_fp] * 3 + [c_i64, c_fp] + [c_fp]),
    "gstex_raster_backward": (c_i, [c_i] * 5 + [c_fp] * 7 + [c_f] * 4 + [c_fp] * 11 + [c_fp, c_fp, c_fp] + [c_fp]),
    "gstex_raster_epilogue": (c_i, [c_i, c_fp, c_fp, c_f] + [c_fp] * 5 + [c_f] * 4 + [c_fp] * 10 + [c_i, c_fp]),
    "gstex_sh_colors_forward": (c_i, [c_i, c_i, c_i] + [c_fp] * 5 + [c_fp]),
    "gstex_sh_colors_backward": (c_i, [c_i, c_i, c_i] + [c_fp] * 5 + [c_i, c_fp]),
    "gstex_preprocess_forward": (c_i, [c_i] + [c_fp] * 12 + [c_fp]),
    "gstex_preprocess_backward": (c_i, [c_i] + [c_fp] * 17 + [c_fp]),
    "gstex_sigmoid_pad_texture": (c_i, [c_i64, c_fp, c_fp, c_fp]),
    "gstex_unpad_texture_grad_sigmoid": (c_i, [c_i64, c_fp, c_fp, c_fp, c_i, c_fp]),
    "gstex_adam_step": (c_i, [c_i64, c_fp, c_fp, c_fp, c_fp] + [C.c_double] * 4 + [c_i, c_f, c_fp]),
    "gstex_adam_state_bytes": (c_sz, []),
    "gstex_adam_prepare_device": (c_i, [c_fp, C.c_double, C.c_double, C.c_double, C.c_double, c_f, c_fp]),
    "gstex_adam_apply_rows_device": (c_i, [c_i64, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_i, c_fp, c_fp]),
    "gstex_adam_step_device": (c_i, [c_i64, c_fp, c_fp, c_fp, c_fp, C.c_double, C.c_double, C.c_double, C.c_double, c_fp,
                                     c_f, c_fp]),
}

# entry points that may be absent from an older build of the library (checked lazily)
_OPTIONAL = set()


def exported_symbols():
    """Names every build of the library must export (tests/test_abi.py checks them against the header)."""
    return sorted(_PROTOS)


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            from .build import build  # builds with nvcc; raises with the compiler output on failure

            build()  # takes an inter-process lock: the ranks of a torchrun launch must not compile into the same files
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                if name in _OPTIONAL:
                    continue
                raise RuntimeError(f"libgstex_b200.so does not export {name}; rebuild with "
                                   f"`python -m gstex_cuda_b200.build --force`") from e
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    msg = load().gstex_last_error()
    return msg.decode() if msg else ""


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")
