"""CPU oracle for the per-frame CRT effect chain (TEST INFRASTRUCTURE ONLY).

Restates, stage by stage and in the reference's arithmetic (same numpy dtypes,
same operation order, same OpenCV entry points), the hot path of
jaylikesbunda/PythonCRT:

    apply_crt_effect      /root/reference/crt_filter.py:531-699   (GUI chain)
    apply_static_effects  /root/reference/crt_filter.py:702-861   (export chain)
    persistence + quantise in process_video          :1086-1098

The chain is expressed as a list of small stage functions driven by one
`ChainParams` record instead of the reference's two long positional-argument
functions; `variant` selects the only two places the GUI and export chains
differ (the glitch offset generator, and who blends the persistence state).

Two interchangeable back ends implement the OpenCV primitives:
  backend="cv2"    calls the same cv2 functions the reference calls (default;
                   this is also what `bench.py` times as the CPU baseline);
  backend="numpy"  uses the pure-numpy restatements in oracle/cv_restated.py.

RNG-driven stages accept injected draws: `noise_plane` (the float32 N(0,1)
plane cv2.randn would have produced, BEFORE the grain up-scale) and
`glitch_table` (integer row/segment offsets), so the CUDA path can be compared
on identical draws.

Pinned by tests/test_oracle_vs_reference.py (runs where /root/reference
exists) and tests/test_oracle_golden.py (committed fixtures made from the
reference by oracle/make_golden.py).
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Optional, Tuple

import numpy as np

from . import cv_restated as R

try:  # the same third-party library the reference calls
    import cv2
except Exception:  # pragma: no cover - cv2 is part of the image
    cv2 = None

F32 = np.float32


@dataclass
class ChainParams:
    """Scalar effect parameters; defaults are the CLI defaults
    (crt_filter.py:1160-1205).  Field names follow process_video's kwargs."""
    scanline_strength: float = 0.6
    scanline_period_px: float = 2.0
    scanline_speed_px_s: float = 30.0
    scanline_angle: float = 0.0
    scanline_thickness: float = 1.0
    triad_strength: float = 0.35
    triad_gamma: float = 2.2
    triad_softness: float = 0.5
    triad_preserve_luma: bool = False
    aberration_px: int = 1
    pixel_size: int = 2
    bloom_sigma: float = 1.2
    bloom_strength: float = 0.25
    bloom_threshold: float = 0.0
    fast_bloom: bool = True
    noise_strength: float = 1.5
    grain_size: int = 1
    vignette_strength: float = 0.25
    flicker_strength: float = 0.0
    flicker_hz: float = 0.0
    persistence: float = 0.2
    glitch_amp_px: int = 0
    glitch_height_frac: float = 0.0
    brightness: float = 0.0
    contrast: float = 1.0
    gamma: float = 1.0
    saturation: float = 1.0
    temperature: float = 0.0
    warp_strength: float = 0.0

    def but(self, **kw) -> "ChainParams":
        return replace(self, **kw)


# ------------------------------------------------------------------ masks ---
def triad_mask(h: int, w: int, strength: float, softness_px: float = 0.0, backend: str = "cv2") -> np.ndarray:
    """Phosphor triad mask, H x W x 3 float32 (crt_filter.py:220-235): channel c
    is 1.0 on columns x % 3 == c and 1-strength elsewhere, optionally softened
    by a horizontal gaussian with REPLICATE border."""
    col = np.arange(w)[None, :]
    base = 1.0 - float(strength)
    planes = [base + float(strength) * (col % 3 == c).astype(F32) for c in range(3)]
    mask = np.repeat(np.stack(planes, axis=2).astype(F32), h, axis=0)
    soft = float(max(0.0, softness_px))
    if soft > 0.0:
        k = max(3, int(round(soft * 3)) * 2 + 1)
        if backend == "cv2":
            mask = cv2.GaussianBlur(mask, (k, 1), sigmaX=soft, sigmaY=0, borderType=cv2.BORDER_REPLICATE)
        else:
            mask = R.gaussian_blur(mask, k, 1, soft, 0)
    return mask.astype(F32)


def vignette_mask(h: int, w: int, strength: float) -> np.ndarray:
    """Radial vignette, H x W float64 (crt_filter.py:266-276)."""
    yy, xx = np.mgrid[0:h, 0:w]
    nx = (xx - (w - 1) / 2.0) / max(1.0, w / 2.0)
    ny = (yy - (h - 1) / 2.0) / max(1.0, h / 2.0)
    return 1.0 - strength * np.clip(nx * nx + ny * ny, 0.0, 1.0)


def scanline_rows(h: int, strength: float, period_px: float, phase_px: float) -> np.ndarray:
    """Per-row sinusoid, float32 (crt_filter.py:213-217)."""
    rows = np.arange(h, dtype=F32)
    wave = 0.5 * (1.0 + np.sin((2.0 * np.pi / max(1e-6, period_px)) * (rows + phase_px)))
    return 1.0 - strength * wave


def scanline_plane(h: int, w: int, strength: float, period_px: float, phase_px: float,
                   angle_deg: float, thickness: float) -> np.ndarray:
    """Slanted / shaped scanline mask, float64 math cast to float32
    (crt_filter.py:308-328)."""
    if strength <= 0.0:
        return np.ones((h, w), dtype=F32)
    yy, xx = np.mgrid[0:h, 0:w]
    slanted = yy + np.tan(np.deg2rad(float(angle_deg))) * xx
    omega = 2.0 * np.pi / max(1e-6, float(period_px))
    wave = 0.5 * (1.0 + np.sin(omega * (slanted + float(phase_px))))
    shaped = np.power(wave, 1.0 / np.clip(float(thickness), 0.1, 4.0))
    return (1.0 - float(strength) * shaped).astype(F32)


# ----------------------------------------------------------------- stages ---
def to_unit_float(frame: np.ndarray) -> np.ndarray:
    """uint8 -> float32 by true division (crt_filter.py:569, :738)."""
    return frame.astype(F32) / 255.0


def aberration(img: np.ndarray, px: int) -> np.ndarray:
    """Channel 0 rolled by +px, channel 2 by -px, circular (crt_filter.py:207-210, :571-577)."""
    if px == 0:
        return img
    return np.stack([np.roll(img[:, :, 0], px, axis=1), img[:, :, 1], np.roll(img[:, :, 2], -px, axis=1)], axis=2)


def pixelate(img: np.ndarray, pixel_size: int, backend: str) -> np.ndarray:
    """NEAREST down then up (crt_filter.py:578-584)."""
    if pixel_size <= 1:
        return img
    h, w = img.shape[:2]
    if backend == "cv2":
        small = cv2.resize(img, (max(1, w // int(pixel_size)), max(1, h // int(pixel_size))), interpolation=cv2.INTER_NEAREST)
        return cv2.resize(small, (w, h), interpolation=cv2.INTER_NEAREST)
    ys, xs = R.pixelate_index(h, pixel_size), R.pixelate_index(w, pixel_size)
    return img[ys][:, xs]


def colour(img: np.ndarray, p: ChainParams) -> np.ndarray:
    """Saturation -> temperature -> brightness/contrast -> gamma, each clipped,
    each only when non-identity (crt_filter.py:279-305).  All float32."""
    if p.saturation != 1.0:
        y = 0.2126 * img[:, :, 0] + 0.7152 * img[:, :, 1] + 0.0722 * img[:, :, 2]
        y = y[:, :, None]
        img = np.clip(y + (img - y) * float(p.saturation), 0.0, 1.0)
    if p.temperature != 0.0:
        t = float(p.temperature)
        gain0 = float(np.clip(1.0 + 0.5 * t, 0.5, 1.5))
        gain2 = float(np.clip(1.0 - 0.5 * t, 0.5, 1.5))
        img[:, :, 0] = np.clip(img[:, :, 0] * gain0, 0.0, 1.0)
        img[:, :, 2] = np.clip(img[:, :, 2] * gain2, 0.0, 1.0)
    if p.brightness != 0.0 or p.contrast != 1.0:
        img = np.clip((img - 0.5) * float(p.contrast) + 0.5 + float(p.brightness), 0.0, 1.0)
    if p.gamma != 1.0 and p.gamma > 0.0:
        img = np.clip(np.power(img, 1.0 / float(p.gamma), dtype=F32), 0.0, 1.0)
    return img


def text_blend(img: np.ndarray, overlay_rgba: np.ndarray) -> np.ndarray:
    """Alpha-blend a same-size RGBA uint8 layer (crt_filter.py:588-598, :653-663)."""
    ov = overlay_rgba
    if ov.dtype != np.uint8:
        ov = np.clip(ov, 0, 255).astype(np.uint8)
    if ov.shape[0] != img.shape[0] or ov.shape[1] != img.shape[1]:
        raise ValueError("oracle expects the text layer at frame size (rasterisation/resizing is out of scope)")
    alpha = ov[:, :, 3:4].astype(F32) / 255.0
    rgb = ov[:, :, :3].astype(F32) / 255.0
    return np.clip(img * (1.0 - alpha) + rgb * alpha, 0.0, 1.0)


def bloom_kernel_size(sigma: float) -> int:
    """k = max(1, int(round(3 sigma))*2+1) with Python's round-half-even (crt_filter.py:609)."""
    return max(1, int(round(sigma * 3)) * 2 + 1)


def bloom(img: np.ndarray, p: ChainParams, backend: str) -> np.ndarray:
    """Threshold, blur (fast 2x down/up or gaussian), add, clip (crt_filter.py:599-612)."""
    if not (p.bloom_strength > 0.0 and (p.bloom_sigma > 0.0 or p.fast_bloom)):
        return img
    h, w = img.shape[:2]
    src = img
    if p.bloom_threshold > 0.0:
        thr = float(min(0.99, max(0.0, p.bloom_threshold)))
        src = np.clip((img - thr) / max(1e-6, (1.0 - thr)), 0.0, 1.0)
    if p.fast_bloom:
        hw, hh = max(1, w // 2), max(1, h // 2)
        if backend == "cv2":
            blur = cv2.resize(cv2.resize(src, (hw, hh), interpolation=cv2.INTER_LINEAR), (w, h), interpolation=cv2.INTER_LINEAR)
        else:
            blur = R.resize_linear(R.resize_linear(src, hw, hh), w, h)
    else:
        k = bloom_kernel_size(p.bloom_sigma)
        if backend == "cv2":
            blur = cv2.GaussianBlur(src, (k, k), sigmaX=p.bloom_sigma, sigmaY=p.bloom_sigma, borderType=cv2.BORDER_REPLICATE)
        else:
            blur = R.gaussian_blur(src, k, k, p.bloom_sigma, p.bloom_sigma)
    return np.clip(img + p.bloom_strength * blur, 0.0, 1.0)


LUT_SIZE = 1024


def triad_luts(gamma: float) -> Tuple[np.ndarray, np.ndarray]:
    """The two 1025-entry float32 tables of crt_filter.py:246-249, :260."""
    grid = np.linspace(0.0, 1.0, LUT_SIZE + 1, dtype=F32)
    g = float(gamma)
    return np.power(grid, g, dtype=F32), np.power(grid, 1.0 / g, dtype=F32)


def apply_triad(img: np.ndarray, mask: np.ndarray, gamma: float, preserve_luma: bool) -> np.ndarray:
    """Mask applied in LUT-linearised light with floor indexing (crt_filter.py:238-263)."""
    g = float(gamma)
    if ((not preserve_luma) and abs(g - 1.0) < 1e-3) or g <= 0.0:
        return np.clip(img * mask, 0.0, 1.0)
    fwd, inv = triad_luts(g)
    scale = float(LUT_SIZE)
    lin = fwd[np.clip((np.clip(img, 0.0, 1.0) * scale).astype(np.int32), 0, LUT_SIZE)]
    lit = lin * mask
    if preserve_luma:
        wr, wg, wb = 0.2126, 0.7152, 0.0722
        before = wr * lin[:, :, 0] + wg * lin[:, :, 1] + wb * lin[:, :, 2]
        after = wr * lit[:, :, 0] + wg * lit[:, :, 1] + wb * lit[:, :, 2]
        ratio = np.clip(before / np.maximum(after, 1e-6), 0.5, 2.0)
        lit = lit * ratio[:, :, None]
    out = inv[np.clip((np.clip(lit, 0.0, 1.0) * scale).astype(np.int32), 0, LUT_SIZE)]
    return np.clip(out, 0.0, 1.0)


def scanlines(img: np.ndarray, p: ChainParams, phase_px: float) -> np.ndarray:
    """crt_filter.py:617-625."""
    if not p.scanline_strength > 0.0:
        return img
    h, w = img.shape[:2]
    if p.scanline_angle == 0.0 and p.scanline_thickness == 1.0:
        line = scanline_rows(h, p.scanline_strength, p.scanline_period_px, phase_px)
        return np.clip(img * line[:, None, None], 0.0, 1.0)
    plane = scanline_plane(h, w, p.scanline_strength, p.scanline_period_px, phase_px, p.scanline_angle, p.scanline_thickness)
    return np.clip(img * plane[:, :, None], 0.0, 1.0)


def flicker_factor(p: ChainParams, time_sec: float):
    """np.float64 scalar of crt_filter.py:632 (or None when the stage is off)."""
    if p.flicker_strength > 0.0 and p.flicker_hz > 0.0:
        return 1.0 + 0.25 * float(p.flicker_strength) * np.sin(2.0 * np.pi * float(p.flicker_hz) * float(time_sec))
    return None


def noise_plane_shape(h: int, w: int, grain_size: int) -> Tuple[int, int]:
    if grain_size and grain_size > 1:
        return max(1, h // int(grain_size)), max(1, w // int(grain_size))
    return h, w


def draw_noise_plane(h: int, w: int, grain_size: int) -> np.ndarray:
    """The N(0,1) float32 plane cv2.randn fills (crt_filter.py:640-641, :644-645);
    uses the calling thread's OpenCV RNG state (cv2.setRNGSeed)."""
    plane = np.empty(noise_plane_shape(h, w, grain_size), dtype=F32)
    cv2.randn(plane, 0.0, 1.0)
    return plane


def noise(img: np.ndarray, p: ChainParams, plane: Optional[np.ndarray], backend: str) -> np.ndarray:
    """Same draw added to all three channels (crt_filter.py:635-648)."""
    if not p.noise_strength > 0.0:
        return img
    h, w = img.shape[:2]
    if plane is None:
        plane = draw_noise_plane(h, w, p.grain_size)
    if plane.shape != (h, w):
        plane = cv2.resize(plane, (w, h), interpolation=cv2.INTER_LINEAR) if backend == "cv2" else R.resize_linear(plane, w, h)
    plane = plane * (p.noise_strength / 255.0)
    return np.clip(img + plane[:, :, None], 0.0, 1.0)


def barrel_maps(h: int, w: int, strength: float) -> Tuple[np.ndarray, np.ndarray]:
    """float32 source-coordinate maps of crt_filter.py:335-346."""
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    xn = (np.arange(w, dtype=F32) - cx) / max(1.0, cx)
    yn = (np.arange(h, dtype=F32) - cy) / max(1.0, cy)
    xv, yv = np.meshgrid(xn, yn)
    factor = 1.0 + (float(strength) * 0.5) * (xv * xv + yv * yv)
    return (xv * factor * cx + cx).astype(F32), (yv * factor * cy + cy).astype(F32)


def barrel_warp(img: np.ndarray, strength: float, backend: str) -> np.ndarray:
    """crt_filter.py:331-348."""
    if float(strength) == 0.0:
        return img
    mx, my = barrel_maps(img.shape[0], img.shape[1], strength)
    if backend == "cv2":
        return cv2.remap(img, mx, my, interpolation=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    return R.remap_bilinear_const0(img, mx, my)


@dataclass
class GlitchTable:
    """Integer horizontal offsets for the bottom rows: offs[row, x // seg_len]."""
    y0: int
    seg_len: int
    offs: np.ndarray  # int32 [rows, segments]


def glitch_table(variant: str, h: int, w: int, amp_px: int, height_frac: float, phase_px: float) -> Optional[GlitchTable]:
    """Offsets drawn exactly as the reference draws them from numpy's PCG64.

    variant "gui":    one offset per row, seed factor 0.05 (crt_filter.py:664-679)
    variant "export": one offset per row and segment, seed factor 2.0 (crt_filter.py:835-853)
    """
    if not (amp_px > 0 and height_frac > 0.0):
        return None
    y0 = max(0, min(h, h - int(h * height_frac)))
    if y0 >= h:
        return None
    rows = h - y0
    ridx = np.arange(rows, dtype=F32)
    if variant == "gui":
        seed = (int(abs(float(phase_px)) * 0.05) + (w << 10) + (h << 1)) & 0xFFFFFFFF
        rng = np.random.default_rng(seed)
        amp = np.asarray(float(amp_px) * np.exp(-3.0 * (ridx / max(1.0, float(rows)))), dtype=F32)
        base = np.clip(rng.normal(loc=0.0, scale=0.5, size=rows).astype(F32), -1.0, 1.0)
        jump = rng.random(rows).astype(F32) < 0.03
        sign = rng.choice(np.array([-1.0, 1.0], dtype=F32), size=rows)
        base = base + jump * sign
        offs = np.rint(np.clip(base * amp, -amp, amp)).astype(np.int32)[:, None]
        return GlitchTable(y0, w, offs)
    seed = (int(abs(float(phase_px)) * 2.0) + (w << 10) + (h << 1)) & 0xFFFFFFFF
    rng = np.random.default_rng(seed)
    seg_len = max(8, min(32, w // 120 if w >= 120 else 8))
    nseg = (w + seg_len - 1) // seg_len
    amp = float(amp_px) * (1.0 - (ridx / max(1.0, float(rows))))
    seg = rng.standard_normal((rows, nseg)).astype(F32) * (amp[:, None] * 0.7)
    walk = np.cumsum(rng.standard_normal(rows).astype(F32)) * 0.1
    walk = np.clip(walk, -amp * 0.4, amp * 0.4)
    offs = np.rint(walk[:, None] + seg).astype(np.int32)
    return GlitchTable(y0, seg_len, offs)


def glitch(img: np.ndarray, table: Optional[GlitchTable]) -> np.ndarray:
    """Circular horizontal gather on the bottom rows (crt_filter.py:680-685, :852-858)."""
    if table is None:
        return img
    h, w = img.shape[:2]
    seg_of_x = np.arange(w, dtype=np.int64) // int(table.seg_len)
    src_x = (np.arange(w, dtype=np.int64)[None, :] + table.offs[:, seg_of_x]) % w
    bottom = img[table.y0:]
    img[table.y0:] = np.take_along_axis(bottom, np.broadcast_to(src_x[:, :, None], bottom.shape), axis=1)
    return img


# ------------------------------------------------------------------ chain ---
def static_chain(frame: np.ndarray, p: ChainParams, *, phase_px: float, time_sec: float = 0.0,
                 variant: str = "export", triad: Optional[np.ndarray] = None,
                 vignette: Optional[np.ndarray] = None, noise_plane: Optional[np.ndarray] = None,
                 glitch_offsets: Optional[GlitchTable] = None, text_rgba: Optional[np.ndarray] = None,
                 text_after: bool = True, backend: str = "cv2", build_masks: bool = True) -> np.ndarray:
    """Everything up to (not including) persistence: the float image that
    apply_static_effects returns (crt_filter.py:735-861) and that
    apply_crt_effect holds at :686.

    `triad` / `vignette` are the mask arrays the reference's callers pass; when
    omitted (and build_masks) they are built from the strengths in `p` the way
    process_video does (crt_filter.py:919-920)."""
    h, w = frame.shape[:2]
    if build_masks and triad is None and p.triad_strength > 0.0:
        triad = triad_mask(h, w, p.triad_strength, p.triad_softness, backend)
    if build_masks and vignette is None and p.vignette_strength > 0.0:
        vignette = vignette_mask(h, w, p.vignette_strength)

    img = to_unit_float(frame)
    img = aberration(img, int(p.aberration_px))
    img = pixelate(img, int(p.pixel_size), backend)
    img = colour(img, p)
    if text_rgba is not None and not text_after:
        img = text_blend(img, text_rgba)
    img = bloom(img, p, backend)
    if triad is not None:
        img = apply_triad(img, triad, p.triad_gamma, p.triad_preserve_luma)
    img = scanlines(img, p, phase_px)
    if vignette is not None:
        img = np.clip(img * vignette[:, :, None], 0.0, 1.0)          # :626-629 (promotes to float64)
    fl = flicker_factor(p, time_sec)
    if fl is not None:
        img = np.clip(img * fl, 0.0, 1.0)                            # :630-634
    img = noise(img, p, noise_plane, backend)
    img = barrel_warp(img, p.warp_strength, backend)
    if text_rgba is not None and text_after:
        img = text_blend(img, text_rgba)
    if glitch_offsets is None:
        glitch_offsets = glitch_table(variant, h, w, int(p.glitch_amp_px), float(p.glitch_height_frac), phase_px)
    img = glitch(img, glitch_offsets)
    return img


def blend_gui(img: np.ndarray, prev: Optional[np.ndarray], persistence: float, backend: str = "cv2") -> np.ndarray:
    """cv2.addWeighted blend of apply_crt_effect (crt_filter.py:687-694)."""
    if prev is None or not persistence > 0.0:
        return img
    if backend == "cv2":
        return cv2.addWeighted(prev, float(persistence), img, float(1.0 - persistence), 0.0)
    return R.add_weighted(prev, float(persistence), img, float(1.0 - persistence))


def blend_export(img: np.ndarray, prev: Optional[np.ndarray], persistence: float) -> np.ndarray:
    """Clipped blend of process_video (crt_filter.py:1086-1096)."""
    if prev is None or not persistence > 0.0:
        return img
    return np.clip(persistence * prev + (1.0 - persistence) * img, 0.0, 1.0)


def quantise(img: np.ndarray, backend: str = "cv2") -> np.ndarray:
    """cv2.convertScaleAbs(img, alpha=255) (crt_filter.py:696, :1098)."""
    if backend == "cv2":
        return cv2.convertScaleAbs(img, alpha=255.0, beta=0)
    return R.convert_scale_abs_255(img)


def frame_step(frame: np.ndarray, p: ChainParams, prev_state: Optional[np.ndarray], *, phase_px: float,
               time_sec: float = 0.0, variant: str = "export", backend: str = "cv2", **kw):
    """One frame through chain + persistence + quantise.  Returns (uint8, state)."""
    img = static_chain(frame, p, phase_px=phase_px, time_sec=time_sec, variant=variant, backend=backend, **kw)
    state = blend_gui(img, prev_state, p.persistence, backend) if variant == "gui" else blend_export(img, prev_state, p.persistence)
    return quantise(state, backend), state


def run_clip(frames, p: ChainParams, fps: float = 30.0, *, variant: str = "export", backend: str = "cv2",
             first_index: int = 0, prev_state: Optional[np.ndarray] = None, noise_planes=None,
             triad: Optional[np.ndarray] = None, vignette: Optional[np.ndarray] = None):
    """Serial clip loop with the reference's per-frame scalars: phase =
    (i / fps) * speed (crt_filter.py:1043), time_sec = i / fps (:1064); masks
    built once per clip (:919-920).  Returns (list of uint8 frames, final state)."""
    frames = list(frames)
    if not frames:
        return [], prev_state
    h, w = frames[0].shape[:2]
    if triad is None and p.triad_strength > 0.0:
        triad = triad_mask(h, w, p.triad_strength, p.triad_softness, backend)
    if vignette is None and p.vignette_strength > 0.0:
        vignette = vignette_mask(h, w, p.vignette_strength)
    outs = []
    state = prev_state
    for j, frame in enumerate(frames):
        i = first_index + j
        phase = (i / float(fps)) * p.scanline_speed_px_s
        plane = None if noise_planes is None else noise_planes[j]
        out, state = frame_step(frame, p, state, phase_px=phase, time_sec=i / float(fps), variant=variant,
                                backend=backend, triad=triad, vignette=vignette, noise_plane=plane, build_masks=False)
        outs.append(out)
    return outs, state


def synthetic_frame(index: int, h: int, w: int, seed: int = 1234) -> np.ndarray:
    """White-noise uint8 test frame."""
    return np.random.default_rng(seed + index).integers(0, 256, (h, w, 3), dtype=np.uint8)


def structured_frame(index: int, h: int, w: int) -> np.ndarray:
    """Ramps (incl. a 0-24 dark ramp that stresses the triad LUT bins), an 8-px
    checker and a moving bright box."""
    yy, xx = np.mgrid[0:h, 0:w]
    f = np.zeros((h, w, 3), np.uint8)
    f[:, :, 0] = (xx * 255 // max(1, w - 1)).astype(np.uint8)
    f[:, :, 1] = (yy * 255 // max(1, h - 1)).astype(np.uint8)
    f[:, :, 2] = (((xx // 8) + (yy // 8)) % 2 * 255).astype(np.uint8)
    band = slice(h // 4, h // 4 + max(1, h // 8))
    f[band, :, :] = ((xx[band] * 25 // max(1, w)) % 25)[:, :, None].astype(np.uint8)
    bx = (index * 7) % max(1, w - w // 6)
    f[h // 2:h // 2 + h // 6, bx:bx + w // 6, :] = 250
    return f
