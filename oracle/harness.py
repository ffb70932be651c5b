"""Drive a parity case through the oracle, or through the real reference where
it is importable (TEST INFRASTRUCTURE ONLY).

The reference loop mirrors what its two callers do around the chain:
  variant "gui"    CRTWindow.on_tick -> apply_crt_effect with the state passed
                   back in (crt_filter.py:1810-1852);
  variant "export" process_video: apply_static_effects, then the ordered
                   persistence blend and convertScaleAbs (crt_filter.py:1043,
                   :1064, :1086-1098).  process_video itself cannot run here
                   (no moviepy/ffmpeg), so those 13 lines are restated.

RNG: before each frame the calling thread's OpenCV RNG is seeded with
`case_noise_seed(case, j)`; the same seed regenerates the plane that is injected
into the oracle / CUDA path (`noise_planes`).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import crt_oracle as O
from .cases import Case, case_frames, case_noise_seed, case_text_layer

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def noise_planes(case: Case) -> Optional[List[np.ndarray]]:
    if not case.params.noise_strength > 0.0:
        return None
    planes = []
    for j in range(case.frames):
        cv2.setRNGSeed(case_noise_seed(case, j))
        planes.append(O.draw_noise_plane(case.h, case.w, case.params.grain_size))
    return planes


def frame_scalars(case: Case, j: int):
    i = case.first_index + j
    return (i / float(case.fps)) * case.params.scanline_speed_px_s, i / float(case.fps)


def run_oracle(case: Case, variant: str, backend: str = "cv2"):
    p = case.params
    frames = case_frames(case)
    planes = noise_planes(case)
    text = case_text_layer(case)
    tri = O.triad_mask(case.h, case.w, p.triad_strength, p.triad_softness, backend) if p.triad_strength > 0.0 else None
    vig = O.vignette_mask(case.h, case.w, p.vignette_strength) if p.vignette_strength > 0.0 else None
    outs, state = [], None
    for j, frame in enumerate(frames):
        phase, tsec = frame_scalars(case, j)
        out, state = O.frame_step(frame, p, state, phase_px=phase, time_sec=tsec, variant=variant, backend=backend,
                                  triad=tri, vignette=vig, noise_plane=None if planes is None else planes[j],
                                  text_rgba=text, text_after=(case.text != "before"), build_masks=False)
        outs.append(out)
    return outs, state


def run_reference(case: Case, variant: str):
    """Same case through the unmodified reference functions."""
    from . import ref_loader
    ref = ref_loader.load()
    p = case.params
    frames = case_frames(case)
    text = case_text_layer(case)
    tri = ref.make_triad_mask(case.h, case.w, p.triad_strength, p.triad_softness) if p.triad_strength > 0.0 else None
    vig = ref.make_vignette(case.h, case.w, p.vignette_strength) if p.vignette_strength > 0.0 else None
    common = dict(time_sec=0.0, brightness=float(p.brightness), contrast=float(p.contrast), gamma=float(p.gamma),
                  saturation=float(p.saturation), temperature=float(p.temperature),
                  flicker_strength=float(p.flicker_strength), flicker_hz=float(p.flicker_hz),
                  grain_size=int(p.grain_size), scanline_angle=float(p.scanline_angle),
                  scanline_thickness=float(p.scanline_thickness), warp_strength=float(p.warp_strength),
                  text_overlay_rgba=text, text_overlay_after=(case.text != "before"))
    outs, state = [], None
    for j, frame in enumerate(frames):
        phase, tsec = frame_scalars(case, j)
        common["time_sec"] = tsec
        cv2.setRNGSeed(case_noise_seed(case, j))
        if variant == "gui":
            out, state = ref.apply_crt_effect(
                frame, p.scanline_strength, tri, float(p.triad_gamma), bool(p.triad_preserve_luma), int(p.aberration_px),
                p.bloom_sigma, p.bloom_strength, float(p.bloom_threshold), p.noise_strength, vig, p.persistence, state,
                p.scanline_period_px, phase, p.fast_bloom, int(p.pixel_size), int(p.glitch_amp_px),
                float(p.glitch_height_frac), **common)
        else:
            img = ref.apply_static_effects(
                frame, p.scanline_strength, tri, float(p.triad_gamma), bool(p.triad_preserve_luma), int(p.aberration_px),
                p.bloom_sigma, p.bloom_strength, float(p.bloom_threshold), p.noise_strength, vig,
                p.scanline_period_px, phase, p.fast_bloom, int(p.pixel_size), int(p.glitch_amp_px),
                float(p.glitch_height_frac), **common)
            if state is not None and p.persistence > 0.0:                       # crt_filter.py:1086-1092
                img = np.clip(p.persistence * state + (1.0 - p.persistence) * img, 0.0, 1.0)
            state = img                                                          # :1096
            out = cv2.convertScaleAbs(img, alpha=255.0, beta=0)                  # :1098
        outs.append(out)
    return outs, state


def diff_stats(a: np.ndarray, b: np.ndarray) -> dict:
    """max |delta| in LSB, fraction of samples off by more than 1, PSNR (dB)."""
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    mse = float(np.mean(d.astype(np.float64) ** 2))
    psnr = float("inf") if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)
    return {"max": int(d.max()) if d.size else 0, "frac_gt1": float((d > 1).mean()) if d.size else 0.0,
            "frac_ne": float((d > 0).mean()) if d.size else 0.0, "psnr": psnr}
