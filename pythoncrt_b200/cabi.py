"""ctypes binding of include/crt_b200.h (the C-ABI shared library).

The library is built in-tree by `pythoncrt_b200.build` (nvcc, sm_100a) as
`pythoncrt_b200/libcrt_b200.so`.  There is no fallback: if it is missing or no
CUDA device is present the import of the compute path raises.
"""
from __future__ import annotations

import ctypes as C
import os

LIB_NAME = "libcrt_b200.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

ABI_VERSION = 2
(TABLE_TRIAD_COLS, TABLE_LUT_FWD, TABLE_LUT_INV, TABLE_GAUSS_TAPS, TABLE_PIXELATE_X, TABLE_PIXELATE_Y,
 TABLE_VIGNETTE_PLANE, TABLE_TEXT_RGBA) = range(8)
VARIANT_GUI, VARIANT_EXPORT = 0, 1
ORDER_RGB, ORDER_BGR = 0, 1
POLICY_AUTO, POLICY_STAGED, POLICY_FUSED = 0, 1, 2

# every symbol include/crt_b200.h declares
EXPORTS = ("crt_abi_version", "crt_create", "crt_destroy", "crt_last_error", "crt_set_params", "crt_set_table",
           "crt_set_policy", "crt_set_shards", "crt_process", "crt_process_static", "crt_process_host", "crt_reset_state", "crt_resize_state",
           "crt_generate_noise", "crt_generate_glitch", "crt_profile_begin", "crt_profile_end", "crt_profile_sample_every")


class CrtParamsC(C.Structure):
    """struct crt_params"""
    _fields_ = [
        ("brightness", C.c_double), ("contrast", C.c_double), ("gamma", C.c_double), ("saturation", C.c_double),
        ("temperature", C.c_double),
        ("aberration_px", C.c_int32), ("pixel_size", C.c_int32),
        ("bloom_sigma", C.c_double), ("bloom_strength", C.c_double), ("bloom_threshold", C.c_double),
        ("fast_bloom", C.c_int32),
        ("triad_on", C.c_int32), ("triad_gamma", C.c_double), ("triad_preserve_luma", C.c_int32),
        ("scanline_strength", C.c_double), ("scanline_period_px", C.c_double), ("scanline_angle", C.c_double),
        ("scanline_thickness", C.c_double),
        ("vignette_on", C.c_int32), ("vignette_strength", C.c_double),
        ("flicker_strength", C.c_double), ("flicker_hz", C.c_double),
        ("noise_strength", C.c_double), ("grain_size", C.c_int32), ("noise_mode", C.c_int32),
        ("noise_seed", C.c_uint64),
        ("warp_strength", C.c_double),
        ("glitch_amp_px", C.c_int32), ("glitch_height_frac", C.c_double), ("glitch_mode", C.c_int32),
        ("text_mode", C.c_int32),
        ("persistence", C.c_double),
        ("variant", C.c_int32),
        ("channel_order", C.c_int32),
        ("reserved", C.c_int32 * 6),
    ]


class CrtFrameC(C.Structure):
    """struct crt_frame"""
    _fields_ = [
        ("phase_px", C.c_double), ("time_sec", C.c_double), ("frame_index", C.c_uint64),
        ("d_noise", C.c_void_p), ("d_glitch_offs", C.c_void_p),
        ("glitch_y0", C.c_int32), ("glitch_seg_len", C.c_int32), ("glitch_segments", C.c_int32),
        ("glitch_rows", C.c_int32),
    ]


class CrtLaunchInfoC(C.Structure):
    """struct crt_launch_info"""
    _fields_ = [("kernels_launched", C.c_int32), ("fused", C.c_int32), ("reserved", C.c_int32 * 6)]


class CrtError(RuntimeError):
    pass


_lib = None


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """dlopen the C-ABI library and declare every prototype.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(path):
        raise CrtError(f"{path} not found: build it with `python -m pythoncrt_b200.build` "
                       "(the CRT chain has no CPU fallback)")
    lib = C.CDLL(path)
    vp, i32, u64, sz = C.c_void_p, C.c_int, C.c_uint64, C.c_size_t
    lib.crt_abi_version.restype = C.c_int
    lib.crt_abi_version.argtypes = []
    lib.crt_create.restype = C.c_int
    lib.crt_create.argtypes = [i32, i32, i32, C.POINTER(vp)]
    lib.crt_destroy.restype = C.c_int
    lib.crt_destroy.argtypes = [vp]
    lib.crt_last_error.restype = C.c_char_p
    lib.crt_last_error.argtypes = [vp]
    lib.crt_set_params.restype = C.c_int
    lib.crt_set_params.argtypes = [vp, C.POINTER(CrtParamsC)]
    lib.crt_set_table.restype = C.c_int
    lib.crt_set_table.argtypes = [vp, i32, vp, sz]
    lib.crt_set_policy.restype = C.c_int
    lib.crt_set_policy.argtypes = [vp, i32]
    lib.crt_set_shards.restype = C.c_int
    lib.crt_set_shards.argtypes = [vp, i32]
    lib.crt_process.restype = C.c_int
    lib.crt_process.argtypes = [vp, vp, vp, vp, i32, C.POINTER(CrtFrameC), i32, vp, C.POINTER(CrtLaunchInfoC)]
    lib.crt_process_static.restype = C.c_int
    lib.crt_process_static.argtypes = [vp, vp, vp, C.POINTER(CrtFrameC), i32, vp, C.POINTER(CrtLaunchInfoC)]
    lib.crt_process_host.restype = C.c_int
    lib.crt_process_host.argtypes = [vp, vp, vp, C.POINTER(CrtFrameC), i32, C.POINTER(CrtLaunchInfoC)]
    lib.crt_reset_state.restype = C.c_int
    lib.crt_reset_state.argtypes = [vp]
    lib.crt_resize_state.restype = C.c_int
    lib.crt_resize_state.argtypes = [vp, vp, i32, i32, vp, vp]
    lib.crt_generate_noise.restype = C.c_int
    lib.crt_generate_noise.argtypes = [vp, u64, vp, vp]
    lib.crt_generate_glitch.restype = C.c_int
    lib.crt_generate_glitch.argtypes = [vp, C.POINTER(CrtFrameC), vp, vp]
    lib.crt_profile_begin.restype = C.c_int
    lib.crt_profile_begin.argtypes = [vp, i32]
    lib.crt_profile_sample_every.restype = C.c_int
    lib.crt_profile_sample_every.argtypes = [vp, i32]
    lib.crt_profile_end.restype = C.c_int
    lib.crt_profile_end.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    if lib.crt_abi_version() != ABI_VERSION:
        raise CrtError(f"{path}: ABI version {lib.crt_abi_version()} != {ABI_VERSION}; rebuild the library")
    _lib = lib
    return lib


def check(lib: C.CDLL, ctx, rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.crt_last_error(ctx)
        raise CrtError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")
