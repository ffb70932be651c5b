"""Named parity cases shared by the golden-vector generator, the oracle tests
and the GPU parity tests (TEST INFRASTRUCTURE ONLY).

Each case = frame size, number of frames, fps, frame source and a ChainParams
override on top of the CLI defaults with noise off (BASELINE.json configs[0]).
The `cfgN` cases are the BASELINE.json configurations at reduced frame size.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict

import numpy as np

from .crt_oracle import ChainParams

BASE = ChainParams(noise_strength=0.0)  # CLI defaults, noise off (configs[0])

GRADE = dict(brightness=0.05, contrast=1.15, gamma=1.1, saturation=1.1, temperature=0.1)
GAUSS = dict(fast_bloom=False, bloom_sigma=1.5, bloom_threshold=0.7, bloom_strength=0.35)
WARP = dict(warp_strength=0.15, scanline_angle=3.0, scanline_thickness=1.2)
LIVE = dict(noise_strength=1.5, grain_size=2, flicker_strength=0.25, flicker_hz=60.0,
            glitch_amp_px=16, glitch_height_frac=0.25)


@dataclass
class Case:
    name: str
    h: int
    w: int
    params: ChainParams
    frames: int = 3
    fps: float = 30.0
    source: str = "noise"      # "noise" | "structured"
    first_index: int = 0
    text: str = ""             # "" | "before" | "after"
    tags: tuple = field(default_factory=tuple)


def _c(name, over: Dict, h=96, w=128, **kw) -> Case:
    return Case(name, h, w, BASE.but(**over), **kw)


CASES = [
    _c("identity", dict(scanline_strength=0.0, triad_strength=0.0, aberration_px=0, pixel_size=1,
                        bloom_strength=0.0, vignette_strength=0.0, persistence=0.0)),
    _c("cfg1_cli_default", {}),
    _c("cfg1_cli_default_structured", {}, source="structured"),
    _c("cfg1_gui_default", dict(triad_preserve_luma=True, scanline_speed_px_s=60.0)),
    _c("cfg2_gauss_grade", {**GAUSS, **GRADE}),
    _c("cfg2_gauss_grade_structured", {**GAUSS, **GRADE}, source="structured"),
    _c("cfg3_warp", WARP),
    _c("cfg3_warp_structured", WARP, source="structured"),
    _c("cfg4_full", {**GAUSS, **GRADE, **WARP, **LIVE}, fps=60.0, first_index=17),
    _c("cfg4_full_fastbloom", {**GRADE, **WARP, **LIVE}, fps=60.0, first_index=5),
    _c("cfg5_gauss_sigma4", {**GRADE, **WARP, **LIVE, **dict(fast_bloom=False, bloom_sigma=4.0, bloom_strength=0.3)}, h=112, w=160),
    _c("neg_aberration_ps3", dict(aberration_px=-5, pixel_size=3)),
    _c("warp_negative", dict(warp_strength=-0.4)),
    _c("warp_strong", dict(warp_strength=1.0, pixel_size=1)),
    _c("noise_grain1", dict(noise_strength=4.0, grain_size=1)),
    _c("noise_grain5", dict(noise_strength=8.0, grain_size=5)),
    _c("triad_gamma1_plain", dict(triad_gamma=1.0)),
    _c("triad_gamma1_luma", dict(triad_gamma=1.0, triad_preserve_luma=True)),
    _c("triad_hard_nosoft", dict(triad_strength=1.0, triad_softness=0.0, triad_gamma=2.4, triad_preserve_luma=True)),
    _c("no_persistence", dict(persistence=0.0)),
    _c("high_persistence", dict(persistence=0.9), frames=6),
    _c("gauss_k3", dict(fast_bloom=False, bloom_sigma=0.4, bloom_strength=0.5)),
    _c("gauss_k5", dict(fast_bloom=False, bloom_sigma=0.7, bloom_strength=0.5)),
    _c("gauss_k25_thr", dict(fast_bloom=False, bloom_sigma=4.0, bloom_strength=1.0, bloom_threshold=0.3)),
    _c("gauss_k13", dict(fast_bloom=False, bloom_sigma=2.0, bloom_strength=0.4, bloom_threshold=0.2)),
    _c("gauss_k61", dict(fast_bloom=False, bloom_sigma=10.0, bloom_strength=0.6), h=112, w=160),
    _c("warp_gauss_noglitch", dict(fast_bloom=False, bloom_sigma=1.5, bloom_strength=0.35, warp_strength=0.3)),
    _c("scan_period3_angle", dict(scanline_period_px=3.0, scanline_angle=-20.0, scanline_thickness=0.5, scanline_speed_px_s=47.0)),
    _c("scan_period5_1d", dict(scanline_period_px=5.0, scanline_speed_px_s=13.0)),
    _c("flicker_only", dict(flicker_strength=1.0, flicker_hz=7.0)),
    _c("glitch_big", dict(glitch_amp_px=64, glitch_height_frac=0.6), first_index=40),
    _c("text_before", dict(), text="before"),
    _c("text_after_warp", dict(warp_strength=0.2), text="after"),
    _c("odd_size", {**WARP, **dict(noise_strength=2.0, grain_size=3)}, h=75, w=101),
    _c("odd_size_gauss", dict(fast_bloom=False, bloom_sigma=1.2), h=75, w=101, tags=("simd_tail",)),
    _c("vga_cfg1", {}, h=480, w=640, frames=2, tags=("large",)),
]
CASES_BY_NAME = {c.name: c for c in CASES}


def case_frames(case: Case):
    from .crt_oracle import structured_frame, synthetic_frame
    make = structured_frame if case.source == "structured" else synthetic_frame
    return [make(case.first_index + j, case.h, case.w) for j in range(case.frames)]


def case_text_layer(case: Case):
    """Deterministic RGBA uint8 layer standing in for the rasterised text."""
    if not case.text:
        return None
    rng = np.random.default_rng(99)
    layer = np.zeros((case.h, case.w, 4), np.uint8)
    y0, x0 = case.h // 5, case.w // 6
    box = layer[y0:y0 + case.h // 3, x0:x0 + case.w // 2]
    box[:, :, :3] = rng.integers(0, 256, box[:, :, :3].shape, dtype=np.uint8)
    box[:, :, 3] = rng.integers(0, 256, box.shape[:2], dtype=np.uint8)
    return layer


def case_noise_seed(case: Case, j: int) -> int:
    return 7000 + 31 * j + (sum(map(ord, case.name)) % 1000)
