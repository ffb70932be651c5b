"""tests/host_emu — TEST INFRASTRUCTURE ONLY (never imported by pythoncrt_b200).

Builds tests/host_emu/emu.cpp with g++ (the kernels' own per-pixel arithmetic
headers compiled for the host) and drives it with the SAME parameter/table
translation the product uses (pythoncrt_b200.config.build_config), so the
arithmetic and the host-side tables can be checked against the oracle on a
machine without a GPU.  What it cannot check — thread mapping, shared-memory
tiling, the fused kernel — is covered by the `gpu` tests on the B200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from pythoncrt_b200 import cabi, tables
from pythoncrt_b200.config import build_config
from pythoncrt_b200.params import CrtParams

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_emu.so")
SRC = os.path.join(HERE, "emu.cpp")
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "pythoncrt_b200", "csrc")
_lib = None


def _stale() -> bool:
    if not os.path.isfile(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("crt_math.cuh", "crt_stages.cuh", "crt_derive.h", "crt_fused.cuh", "crt_fused_ps2.cuh",
                                                   "crt_fused_warp_src.cuh", "crt_launch.h", "crt_tma.cuh", "crt_policy.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def lib():
    global _lib
    if _lib is None:
        if _stale():
            subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I/usr/local/cuda/include",
                            "-o", SO, SRC], check=True)
        _lib = C.CDLL(SO)
        _lib.emu_process.restype = C.c_int
    return _lib


def oracle_to_product_params(p) -> CrtParams:
    """oracle.ChainParams -> pythoncrt_b200.CrtParams (same field names)."""
    names = {f for f in CrtParams.__dataclass_fields__}
    return CrtParams(**{k: v for k, v in vars(p).items() if k in names})


def run_case(case, variant: str, static: bool = False, channel_order: str = "rgb"):
    """Run an oracle Case through the host build of the kernel arithmetic.  With channel_order="bgr" the
    frames are fed channel-swapped and the results swapped back (must equal the "rgb" run exactly)."""
    from oracle import harness
    from oracle.cases import case_frames, case_text_layer
    L = lib()
    p = oracle_to_product_params(case.params)
    H, W = case.h, case.w
    text = case_text_layer(case)
    c, tabs = build_config(p, W, H, variant=variant, text_rgba=text, text_after=(case.text != "before"), channel_order=channel_order)
    assert L.emu_sizeof_params() == C.sizeof(cabi.CrtParamsC) and L.emu_sizeof_frame() == C.sizeof(cabi.CrtFrameC)
    ptrs = (C.c_void_p * 8)()
    sizes = (C.c_size_t * 8)()
    for k, a in tabs.items():
        ptrs[k] = a.ctypes.data
        sizes[k] = a.nbytes
    frames = np.ascontiguousarray(np.stack(case_frames(case)))
    if channel_order == "bgr":
        frames = np.ascontiguousarray(frames[..., ::-1])
    n = frames.shape[0]
    planes = harness.noise_planes(case)
    recs = (cabi.CrtFrameC * n)()
    keep = []
    for j in range(n):
        phase, tsec = harness.frame_scalars(case, j)
        recs[j].phase_px, recs[j].time_sec, recs[j].frame_index = phase, tsec, case.first_index + j
        if planes is not None:
            pl = np.ascontiguousarray(planes[j], np.float32)
            keep.append(pl)
            recs[j].d_noise = pl.ctypes.data
        t = tables.glitch_offsets(variant, H, W, int(p.glitch_amp_px), float(p.glitch_height_frac), phase)
        if t is not None:
            keep.append(t)
            recs[j].d_glitch_offs = t.ctypes.data
            recs[j].glitch_y0, recs[j].glitch_rows, recs[j].glitch_seg_len, recs[j].glitch_segments = \
                tables.glitch_geometry(variant, H, W, p.glitch_height_frac)
    out = np.empty_like(frames)
    state = np.zeros((H, W, 3), np.float32)
    img = np.empty(frames.shape, np.float32) if static else None
    err = C.create_string_buffer(512)
    rc = L.emu_process(C.byref(c), W, H, ptrs, sizes, frames.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                       state.ctypes.data_as(C.c_void_p), 0, img.ctypes.data_as(C.c_void_p) if static else None,
                       recs, n, err, 512)
    if rc != 0:
        raise RuntimeError(f"emu_process failed ({rc}): {err.value.decode()}")
    if channel_order == "bgr":
        out, state = np.ascontiguousarray(out[..., ::-1]), np.ascontiguousarray(state[..., ::-1])
        img = None if img is None else np.ascontiguousarray(img[..., ::-1])
    return (img if static else list(out)), state


def plan(params: CrtParams, W: int, H: int, variant: str = "export"):
    """Tile plan the library would choose for the fused kernel: dict(ok, th, cap_px, cap_aux, smem, why)."""
    L = lib()
    c, tabs = build_config(params, W, H, variant=variant)
    ptrs = (C.c_void_p * 8)()
    sizes = (C.c_size_t * 8)()
    for k, a in tabs.items():
        ptrs[k] = a.ctypes.data
        sizes[k] = a.nbytes
    ps = int(params.pixel_size)
    uni = 0
    if ps > 1:
        xs, ys = tabs[cabi.TABLE_PIXELATE_X], tabs[cabi.TABLE_PIXELATE_Y]
        if np.array_equal(xs, ps * (np.arange(W) // ps)) and np.array_equal(ys, ps * (np.arange(H) // ps)):
            uni = ps
    out = (C.c_longlong * 5)()
    why = C.create_string_buffer(256)
    rc = L.emu_plan(C.byref(c), W, H, ptrs, sizes, uni, out, why, 256)
    if rc:
        raise RuntimeError(why.value.decode())
    return dict(ok=bool(out[0]), th=out[1], cap_px=out[2], cap_aux=out[3], smem=out[4], why=why.value.decode())


def check_warp_src(params: CrtParams, W: int, H: int, variant: str = "export"):
    """Planner of the source-driven warp kernel on the host: dict(ok, tiles, violations, max_quads, box_px, why)."""
    L = lib()
    c, tabs = build_config(params, W, H, variant=variant)
    ptrs = (C.c_void_p * 8)()
    sizes = (C.c_size_t * 8)()
    for k, a in tabs.items():
        ptrs[k] = a.ctypes.data
        sizes[k] = a.nbytes
    ps = int(params.pixel_size)
    uni = ps if ps > 1 and W % ps == 0 and H % ps == 0 else 0
    out = (C.c_longlong * 5)()
    why = C.create_string_buffer(256)
    rc = L.emu_check_warp_src(C.byref(c), W, H, ptrs, sizes, uni, out, why, 256)
    if rc:
        raise RuntimeError(why.value.decode())
    return dict(ok=bool(out[0]), tiles=out[1], violations=out[2], max_quads=out[3], box_px=out[4], why=why.value.decode())


def policy(w: int, h: int, *, sms: int = 148, per_sm: int = 4, halo_blocks: int = 1, gaussian: bool = False, shards_wanted: int = 0,
           n_frames: int = 600, persistence: float = 0.2) -> dict:
    """csrc/crt_policy.h: the C ABI's scheduling decisions for a frame size / kernel family (pure functions, no GPU)."""
    out = (C.c_int * 6)()
    lib().emu_policy(w, h, sms, per_sm, halo_blocks, int(gaussian), shards_wanted, n_frames, C.c_double(persistence), out)
    return dict(tile_h=out[0], clip_tile_h=out[1], clip_size_ok=bool(out[2]), shards=out[3], auto_prefers_clip=bool(out[4]), halo=out[5])
