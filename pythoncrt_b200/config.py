"""Translate the reference's parameter surface into the C-ABI inputs
(`crt_params` + host-built tables).  Pure host code: no CUDA needed."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from . import cabi, tables
from .params import CrtParams


def build_config(params: CrtParams, W: int, H: int, *, variant: str = "export", triad_cols="auto", vignette="auto",
                 text_rgba: Optional[np.ndarray] = None, text_after: bool = True, noise_mode: str = "inject",
                 glitch_mode: str = "inject", seed: int = 0, channel_order: str = "rgb") -> Tuple[cabi.CrtParamsC, Dict[int, np.ndarray]]:
    """Return (crt_params struct, {crt_table id: contiguous ndarray}).  See CrtEngine.configure."""
    if channel_order not in ("rgb", "bgr"):
        raise ValueError("channel_order must be 'rgb' (index 0 = R, like the reference) or 'bgr'")
    tabs: Dict[int, np.ndarray] = {}
    p = params
    c = cabi.CrtParamsC()
    for name in ("brightness", "contrast", "gamma", "saturation", "temperature", "bloom_sigma", "bloom_strength",
                 "bloom_threshold", "triad_gamma", "scanline_strength", "scanline_period_px", "scanline_angle",
                 "scanline_thickness", "flicker_strength", "flicker_hz", "noise_strength", "warp_strength",
                 "glitch_height_frac", "persistence"):
        setattr(c, name, float(getattr(p, name)))
    c.aberration_px, c.pixel_size = int(p.aberration_px), max(1, int(p.pixel_size))
    c.fast_bloom = int(bool(p.fast_bloom))
    c.triad_preserve_luma = int(bool(p.triad_preserve_luma))
    c.grain_size = int(p.grain_size)
    c.glitch_amp_px = int(p.glitch_amp_px)
    c.variant = cabi.VARIANT_GUI if variant == "gui" else cabi.VARIANT_EXPORT
    c.noise_mode = 1 if noise_mode == "generate" else 0
    c.glitch_mode = 1 if glitch_mode == "generate" else 0
    c.noise_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    c.channel_order = cabi.ORDER_BGR if channel_order == "bgr" else cabi.ORDER_RGB

    # triad column table + LUTs
    if isinstance(triad_cols, str):
        triad_cols = tables.triad_columns(W, p.triad_strength, p.triad_softness) if p.triad_strength > 0.0 else None
    if triad_cols is not None:
        t = np.asarray(triad_cols)
        if t.ndim == 3:
            t = t[0]
        if t.shape != (W, 3):
            raise ValueError(f"triad table must be [W={W}][3], got {t.shape}")
        tabs[cabi.TABLE_TRIAD_COLS] = t.astype(np.float32)
        c.triad_on = 1
        g = float(p.triad_gamma)
        if not (((not p.triad_preserve_luma) and abs(g - 1.0) < 1e-3) or g <= 0.0):
            fwd, inv = tables.triad_luts(g)
            tabs[cabi.TABLE_LUT_FWD] = fwd
            tabs[cabi.TABLE_LUT_INV] = inv
    # vignette
    if isinstance(vignette, str):
        vignette = float(p.vignette_strength) if p.vignette_strength > 0.0 else None
    if vignette is None:
        c.vignette_on = 0
    elif np.isscalar(vignette):
        c.vignette_on, c.vignette_strength = 1, float(vignette)
    else:
        s = tables.infer_vignette_strength(vignette)
        if s is not None:
            c.vignette_on, c.vignette_strength = 1, s
        else:
            v = np.asarray(vignette)
            if v.shape != (H, W):
                raise ValueError(f"vignette plane must be [H={H}][W={W}], got {v.shape}")
            tabs[cabi.TABLE_VIGNETTE_PLANE] = v.astype(np.float32)
            c.vignette_on = 2
    # bloom taps
    if p.bloom_strength > 0.0 and p.bloom_sigma > 0.0 and not p.fast_bloom:
        k = tables.bloom_ksize(p.bloom_sigma)
        tabs[cabi.TABLE_GAUSS_TAPS] = tables.gaussian_taps(k, p.bloom_sigma)
    # pixelate
    if int(p.pixel_size) > 1:
        tabs[cabi.TABLE_PIXELATE_X] = tables.pixelate_table(W, int(p.pixel_size))
        tabs[cabi.TABLE_PIXELATE_Y] = tables.pixelate_table(H, int(p.pixel_size))
    # text layer
    if text_rgba is not None:
        ov = np.asarray(text_rgba)
        if ov.dtype != np.uint8:
            ov = np.clip(ov, 0, 255).astype(np.uint8)
        if ov.shape != (H, W, 4):
            raise ValueError("text layer must be rasterised at frame size [H][W][4] (resizing is host-side, out of scope)")
        tabs[cabi.TABLE_TEXT_RGBA] = ov
        c.text_mode = 2 if text_after else 1
    return c, {k: np.ascontiguousarray(v) for k, v in tabs.items()}
