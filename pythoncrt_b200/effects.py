"""Drop-in replacements for the reference's two per-frame entry points.

    apply_crt_effect      /root/reference/crt_filter.py:531-699  (GUI tick, :1810 / :1972)
    apply_static_effects  /root/reference/crt_filter.py:702-861  (export worker, :1045)

Same positional/keyword arguments, same meaning, same return shapes.  Frames are
numpy uint8 H x W x 3 on the host (possibly read-only, :501); they are copied to
the GPU, run through the C-ABI chain and copied back.  Differences a caller can
observe, all deliberate:

  * the persistence state returned by apply_crt_effect is a `DeviceState` (the
    float32 state kept in HBM) instead of a float ndarray; it has `.shape`,
    converts with `np.asarray(state)` and is accepted back as `state_prev`, which
    is all the reference's callers do with it (:1810-1823, :689);
  * noise uses the device's counter-based generator (the reference's cv2.randn
    stream is thread-local and not reproducible) unless draws
    are injected with the extra keyword `noise_plane=`;
  * glitch offsets are drawn on the host from numpy's PCG64 exactly like the
    reference, so glitch output is identical.

One engine (C-ABI context) is cached per (thread, device, frame size), which is
what makes the reference's two-worker export pool (:1015-1017) safe.
"""
from __future__ import annotations

import threading
from typing import Optional, Tuple

import numpy as np

from . import tables
from .params import CrtParams

_local = threading.local()
_frame_counter = threading.local()


class DeviceState:
    """float32 H x W x 3 persistence state resident on the GPU."""

    def __init__(self, tensor):
        self.tensor = tensor

    @property
    def shape(self) -> Tuple[int, ...]:
        return tuple(self.tensor.shape)

    dtype = np.dtype(np.float32)

    def numpy(self) -> np.ndarray:
        return self.tensor.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a


def _engine_for(h: int, w: int, device: int):
    from .engine import CrtEngine
    cache = getattr(_local, "engines", None)
    if cache is None:
        cache = _local.engines = {}
    key = (device, h, w)
    if key not in cache:
        import torch
        eng = CrtEngine(w, h, device)
        eng._cfg_key = None
        eng._pin_in = torch.empty((h, w, 3), dtype=torch.uint8).pin_memory()
        eng._pin_out = torch.empty((h, w, 3), dtype=torch.uint8).pin_memory()
        eng._dev_in = torch.empty((1, h, w, 3), dtype=torch.uint8, device=f"cuda:{device}")
        eng._dev_out = torch.empty((1, h, w, 3), dtype=torch.uint8, device=f"cuda:{device}")
        cache[key] = eng
    return cache[key]


def _mask_key(mask) -> Optional[tuple]:
    if mask is None:
        return None
    s = getattr(mask, "strength", None)
    if s is not None:
        return ("made", float(s), float(getattr(mask, "softness", 0.0)))
    a = np.asarray(mask)
    row = np.ascontiguousarray(a[0] if a.ndim == 3 else a[:: max(1, a.shape[0] // 16)])
    return ("raw", a.shape, hash(row.tobytes()))


def _configure(eng, p: CrtParams, variant: str, triad_mask, vignette_mask, text, text_after, inject_noise: bool):
    key = (p, variant, _mask_key(triad_mask), _mask_key(vignette_mask), None if text is None else hash(np.asarray(text).tobytes()),
           bool(text_after), inject_noise)
    if eng._cfg_key == key:
        return
    triad_cols = None if triad_mask is None else np.asarray(triad_mask)[0]
    eng.configure(p, variant=variant, triad_cols=triad_cols, vignette=vignette_mask, text_rgba=text, text_after=text_after,
                  noise_mode="inject" if inject_noise else "generate", glitch_mode="inject", seed=0x5EED)
    eng._cfg_key = key


def _params(scanline_strength, triad_gamma, triad_preserve_luma, aberration_px, bloom_sigma, bloom_strength, bloom_threshold,
            noise_strength, persistence, scanline_period_px, fast_bloom, pixel_size, glitch_amp_px, glitch_height_frac,
            brightness, contrast, gamma, saturation, temperature, flicker_strength, flicker_hz, grain_size, scanline_angle,
            scanline_thickness, warp_strength) -> CrtParams:
    return CrtParams(
        scanline_strength=float(scanline_strength), triad_gamma=float(triad_gamma), triad_preserve_luma=bool(triad_preserve_luma),
        aberration_px=int(aberration_px), bloom_sigma=float(bloom_sigma), bloom_strength=float(bloom_strength),
        bloom_threshold=float(bloom_threshold), noise_strength=float(noise_strength), persistence=float(persistence),
        scanline_period_px=float(scanline_period_px), fast_bloom=bool(fast_bloom), pixel_size=int(pixel_size),
        glitch_amp_px=int(glitch_amp_px), glitch_height_frac=float(glitch_height_frac), brightness=float(brightness),
        contrast=float(contrast), gamma=float(gamma), saturation=float(saturation), temperature=float(temperature),
        flicker_strength=float(flicker_strength), flicker_hz=float(flicker_hz), grain_size=int(grain_size),
        scanline_angle=float(scanline_angle), scanline_thickness=float(scanline_thickness), warp_strength=float(warp_strength),
        # strengths of the masks travel with the mask arrays, not with the scalars
        triad_strength=0.0, triad_softness=0.0, vignette_strength=0.0, scanline_speed_px_s=0.0)


def _next_index() -> int:
    i = getattr(_frame_counter, "i", 0)
    _frame_counter.i = i + 1
    return i


def _check_text(text_overlay_rgba, h, w):
    if text_overlay_rgba is None:
        return None
    ov = np.asarray(text_overlay_rgba)
    if ov.shape[0] != h or ov.shape[1] != w:
        from PIL import Image   # same host-side resize as the reference (:594)
        if ov.dtype != np.uint8:
            ov = np.clip(ov, 0, 255).astype(np.uint8)
        ov = np.asarray(Image.fromarray(ov, mode="RGBA").resize((w, h), Image.BILINEAR))
    return ov


def apply_crt_effect(frame, scanline_strength, triad_mask, triad_gamma, triad_preserve_luma, aberration_px, bloom_sigma,
                     bloom_strength, bloom_threshold, noise_strength, vignette_mask, persistence, state_prev,
                     scanline_period_px, scanline_phase_px, fast_bloom, pixel_size, glitch_amp_px=0, glitch_height_frac=0.0,
                     time_sec=0.0, brightness=0.0, contrast=1.0, gamma=1.0, saturation=1.0, temperature=0.0,
                     flicker_strength=0.0, flicker_hz=0.0, grain_size=1, scanline_angle=0.0, scanline_thickness=1.0,
                     warp_strength=0.0, text_overlay_rgba=None, text_overlay_after=True, *, noise_plane=None, device: int = 0):
    """GUI chain incl. persistence and uint8 quantise; returns (uint8 H x W x 3, DeviceState)."""
    import torch
    frame = np.asarray(frame)
    h, w = frame.shape[0], frame.shape[1]
    eng = _engine_for(h, w, device)
    p = _params(scanline_strength, triad_gamma, triad_preserve_luma, aberration_px, bloom_sigma, bloom_strength, bloom_threshold,
                noise_strength, persistence, scanline_period_px, fast_bloom, pixel_size, glitch_amp_px, glitch_height_frac,
                brightness, contrast, gamma, saturation, temperature, flicker_strength, flicker_hz, grain_size, scanline_angle,
                scanline_thickness, warp_strength)
    text = _check_text(text_overlay_rgba, h, w)
    _configure(eng, p, "gui", triad_mask, vignette_mask, text, text_overlay_after, noise_plane is not None)
    # state handling (:687-694): blend only when state_prev is given and persistence > 0
    if isinstance(state_prev, DeviceState) and state_prev.shape == (h, w, 3):
        state, valid = state_prev.tensor, True
    elif state_prev is not None:
        prev = np.asarray(state_prev, dtype=np.float32)
        if prev.shape != (h, w, 3):
            import cv2  # the reference resizes a stale state on the host (:689-690)
            prev = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)
        state, valid = torch.from_numpy(np.ascontiguousarray(prev)).to(eng._dev_in.device), True
    else:
        state, valid = eng.new_state(), False
    eng._pin_in.numpy()[...] = frame
    eng._dev_in[0].copy_(eng._pin_in, non_blocking=True)
    kw = dict(phases=[float(scanline_phase_px)], times=[float(time_sec)], first_index=_next_index())
    if noise_plane is not None:
        kw["noise_planes"] = torch.from_numpy(np.ascontiguousarray(noise_plane, np.float32)).to(eng._dev_in.device)[None]
    eng.process(eng._dev_in, eng._dev_out, state=state, state_valid=valid, **kw)
    eng._pin_out.copy_(eng._dev_out[0], non_blocking=True)
    torch.cuda.current_stream(eng._dev_in.device).synchronize()
    return eng._pin_out.numpy().copy(), DeviceState(state)


def apply_static_effects(frame, scanline_strength, triad_mask, triad_gamma, triad_preserve_luma, aberration_px, bloom_sigma,
                         bloom_strength, bloom_threshold, noise_strength, vignette_mask, scanline_period_px, scanline_phase_px,
                         fast_bloom, pixel_size, glitch_amp_px, glitch_height_frac, time_sec=0.0, brightness=0.0, contrast=1.0,
                         gamma=1.0, saturation=1.0, temperature=0.0, flicker_strength=0.0, flicker_hz=0.0, grain_size=1,
                         scanline_angle=0.0, scanline_thickness=1.0, warp_strength=0.0, text_overlay_rgba=None,
                         text_overlay_after=True, *, noise_plane=None, device: int = 0) -> np.ndarray:
    """Export chain, stateless; returns the float32 image that process_video blends (:1084-1098)."""
    import torch
    frame = np.asarray(frame)
    h, w = frame.shape[0], frame.shape[1]
    eng = _engine_for(h, w, device)
    p = _params(scanline_strength, triad_gamma, triad_preserve_luma, aberration_px, bloom_sigma, bloom_strength, bloom_threshold,
                noise_strength, 0.0, scanline_period_px, fast_bloom, pixel_size, glitch_amp_px, glitch_height_frac,
                brightness, contrast, gamma, saturation, temperature, flicker_strength, flicker_hz, grain_size, scanline_angle,
                scanline_thickness, warp_strength)
    text = _check_text(text_overlay_rgba, h, w)
    _configure(eng, p, "export", triad_mask, vignette_mask, text, text_overlay_after, noise_plane is not None)
    eng._pin_in.numpy()[...] = frame
    eng._dev_in[0].copy_(eng._pin_in, non_blocking=True)
    kw = dict(phases=[float(scanline_phase_px)], times=[float(time_sec)], first_index=_next_index())
    if noise_plane is not None:
        kw["noise_planes"] = torch.from_numpy(np.ascontiguousarray(noise_plane, np.float32)).to(eng._dev_in.device)[None]
    img = eng.process_static(eng._dev_in, **kw)
    return img[0].cpu().numpy()


def install(reference_module) -> None:
    """Monkey-patch an imported reference module (crt_filter) so that its unmodified
    GUI (:1810, :1972) and export pool (:1045) run the chain on the GPU."""
    reference_module.apply_crt_effect = apply_crt_effect
    reference_module.apply_static_effects = apply_static_effects
    reference_module.make_triad_mask = tables.make_triad_mask
    reference_module.make_vignette = tables.make_vignette
