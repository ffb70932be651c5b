"""Host-side table builders for the C ABI (`crt_set_table`, include/crt_b200.h).

Small, per-clip (not per-pixel) data the kernels index: the triad column mask,
the two gamma look-up tables, gaussian taps, pixelate index tables, glitch
offset tables.  They are built with numpy on the host because their values must
be bit-identical to what the reference indexes (numpy's float32 `power`, and
OpenCV's separable-filter arithmetic, restated here); the per-pixel work all
happens on the device.

Also provides drop-in equivalents of the reference's mask constructors
(`make_triad_mask` crt_filter.py:220-235, `make_vignette` :266-276) that
remember their parameters, so the device path never has to upload an H x W mask.
"""
from __future__ import annotations

import functools
from typing import Optional, Tuple

import numpy as np

F32 = np.float32
LUT_SIZE = 1024


# ------------------------------------------------------------ gaussian taps --
def bloom_ksize(sigma: float) -> int:
    """Kernel width the reference asks OpenCV for (crt_filter.py:609):
    max(1, int(round(3 sigma)) * 2 + 1), Python round = half-to-even."""
    return max(1, int(round(float(sigma) * 3)) * 2 + 1)


def gaussian_taps(ksize: int, sigma: float) -> np.ndarray:
    """float32 taps equal to cv2.getGaussianKernel(ksize, sigma, CV_32F): left
    half of exp(-x^2 / 2 sigma^2) in double, normalised by the reciprocal of
    (2 * sum + 1), mirrored, narrowed to float32."""
    n, sigma = int(ksize), float(sigma)
    half = (n - 1) // 2
    x = np.arange(1 - n, 1 - n + 2 * half, 2, dtype=np.float64)
    vals = np.exp(x * x * (-0.5 / (sigma * sigma)) * 0.25)
    total = 0.0
    for v in vals:
        total += float(v)
    total = total * 2.0 + 1.0 + (1.0 if n % 2 == 0 else 0.0)
    mul = 1.0 / total
    taps = np.empty(n, np.float64)
    taps[:half] = vals * mul
    taps[n - half:] = (vals * mul)[::-1]
    taps[half] = mul
    if n % 2 == 0:
        taps[half + 1] = mul
    return taps.astype(F32)


def _fma32(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F32)


def _row_blur_replicate(cols: np.ndarray, taps: np.ndarray) -> np.ndarray:
    """Horizontal pass of cv2.GaussianBlur on a [W][C] float32 row with
    BORDER_REPLICATE, in OpenCV's float32 operation order (k = 3 and 5 use the
    small-kernel form, larger kernels accumulate left to right with fma)."""
    k = len(taps)
    r = k // 2
    w = cols.shape[0]
    idx = np.clip(np.arange(-r, w + r), 0, w - 1)
    p = cols[idx]
    tap = lambda i: p[i:i + w]  # noqa: E731
    if k == 3:
        return _fma32(tap(1), taps[1], (tap(0) + tap(2)) * taps[2])
    if k == 5:
        return _fma32(tap(4) + tap(0), taps[4], _fma32(tap(2), taps[2], (tap(1) + tap(3)) * taps[3]))
    acc = tap(0) * taps[0]
    for i in range(1, k):
        acc = _fma32(tap(i), taps[i], acc)
    return acc


# -------------------------------------------------------------- triad mask --
def triad_columns(w: int, strength: float, softness_px: float = 0.0) -> np.ndarray:
    """Row 0 of the reference's triad mask as a [W][3] float32 table
    (crt_filter.py:220-235; the mask is identical on every row).  Memoised: the GUI rebuilds the
    mask on every preview tick (:1813) with unchanged arguments."""
    return _triad_columns_cached(int(w), float(strength), float(softness_px)).copy()


@functools.lru_cache(maxsize=16)
def _triad_columns_cached(w: int, strength: float, softness_px: float) -> np.ndarray:
    col = np.arange(w)
    base = 1.0 - float(strength)
    cols = np.stack([(base + float(strength) * (col % 3 == c).astype(F32)) for c in range(3)], axis=1).astype(F32)
    soft = float(max(0.0, softness_px))
    if soft > 0.0:
        k = max(3, int(round(soft * 3)) * 2 + 1)
        cols = _row_blur_replicate(cols, gaussian_taps(k, soft))
    return np.ascontiguousarray(cols, dtype=F32)


class _MadeMask(np.ndarray):
    """ndarray that remembers the parameters it was built from.  The memory survives only on exact views of the
    whole mask (same shape, dtype and values); anything derived — `mask * 0.5`, `mask.astype(...)`, a slice —
    forgets them (None), so that such an array is inspected numerically / uploaded as a plane instead of being
    taken for the mask it came from."""
    _REMEMBER = ()

    def __array_finalize__(self, obj):
        for name in self._REMEMBER:
            setattr(self, name, None)

    def __array_wrap__(self, out, context=None, return_scalar=False):       # ufunc results are plain arrays
        out = np.asarray(out)
        return out[()] if return_scalar else out


class TriadMask(_MadeMask):
    """H x W x 3 float32 mask (a broadcast view of one row) that remembers how it was made."""
    _REMEMBER = ("strength", "softness")
    strength: Optional[float] = None
    softness: Optional[float] = None


def make_triad_mask(h: int, w: int, strength: float, softness_px: float = 0.0) -> np.ndarray:
    """Drop-in for make_triad_mask (crt_filter.py:220-235)."""
    cols = triad_columns(w, strength, softness_px)
    m = np.broadcast_to(cols[None, :, :], (h, w, 3)).view(TriadMask)
    m.strength, m.softness = float(strength), float(softness_px)
    return m


def triad_luts(gamma: float) -> Tuple[np.ndarray, np.ndarray]:
    """The forward (x^g) and inverse (x^(1/g)) 1025-entry float32 tables of
    _apply_triad_mask (crt_filter.py:246-249, :260), numpy's own float32 power."""
    grid = np.linspace(0.0, 1.0, LUT_SIZE + 1, dtype=F32)
    g = float(gamma)
    return np.power(grid, g, dtype=F32), np.power(grid, 1.0 / g, dtype=F32)


# ---------------------------------------------------------------- vignette --
class VignetteMask(_MadeMask):
    """H x W float64 vignette (crt_filter.py:266-276) that remembers its strength."""
    _REMEMBER = ("strength",)
    strength: Optional[float] = None


def make_vignette(h: int, w: int, strength: float) -> np.ndarray:
    """Drop-in for make_vignette (crt_filter.py:266-276)."""
    yy, xx = np.mgrid[0:h, 0:w]
    nx = (xx - (w - 1) / 2.0) / max(1.0, w / 2.0)
    ny = (yy - (h - 1) / 2.0) / max(1.0, h / 2.0)
    v = (1.0 - strength * np.clip(nx * nx + ny * ny, 0.0, 1.0)).view(VignetteMask)
    v.strength = float(strength)
    return v


class LazyVignette:
    """make_vignette for callers that rebuild the mask on every preview tick (the GUI does, crt_filter.py:1821): carries
    (h, w, strength) and builds the H x W float64 array only if somebody asks for it (np.asarray).  The device path never
    does: it evaluates the vignette analytically from the strength (crt_params.vignette_on = 1)."""

    def __init__(self, h: int, w: int, strength: float):
        self.shape, self.strength, self.dtype, self.ndim = (int(h), int(w)), float(strength), np.dtype(np.float64), 2

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(make_vignette(self.shape[0], self.shape[1], self.strength))
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, idx):
        return np.asarray(self)[idx]


def make_vignette_lazy(h: int, w: int, strength: float) -> LazyVignette:
    """Drop-in for make_vignette (crt_filter.py:266-276) without the per-tick H x W host array."""
    return LazyVignette(h, w, strength)


def infer_vignette_strength(mask: np.ndarray) -> Optional[float]:
    """Recover `strength` from a mask made by the reference's make_vignette, or
    None if the array is not such a mask (then it is uploaded as a plane)."""
    s = getattr(mask, "strength", None)
    if s is not None:
        return float(s)
    h, w = mask.shape[:2]
    nx0 = (0 - (w - 1) / 2.0) / max(1.0, w / 2.0)
    ny0 = (0 - (h - 1) / 2.0) / max(1.0, h / 2.0)
    r2 = min(1.0, nx0 * nx0 + ny0 * ny0)
    if r2 <= 0:
        return None
    s = (1.0 - float(mask[0, 0])) / r2
    if np.asarray(mask).dtype != np.float64:
        s = float(np.round(s, 6))
    ys = np.array([0, h // 3, h // 2, h - 1, h // 5])
    xs = np.array([0, w // 2, w // 3, w - 1, (4 * w) // 5])
    nx = (xs - (w - 1) / 2.0) / max(1.0, w / 2.0)
    ny = (ys - (h - 1) / 2.0) / max(1.0, h / 2.0)
    want = 1.0 - s * np.clip(nx * nx + ny * ny, 0.0, 1.0)
    atol = 1e-12 if np.asarray(mask).dtype == np.float64 else 2e-7          # a float32 copy of the mask is still that mask
    if np.allclose(np.asarray(mask)[ys, xs], want, rtol=0, atol=atol):
        return float(s)
    return None


# ---------------------------------------------------------------- pixelate --
def _nearest_index(n_dst: int, n_src: int) -> np.ndarray:
    """cv2.resize INTER_NEAREST source index: min(floor(d / (n_dst / n_src)), n_src - 1)."""
    inv = 1.0 / (float(n_dst) / float(n_src))
    return np.minimum(np.floor(np.arange(n_dst, dtype=np.float64) * inv).astype(np.int64), n_src - 1)


def pixelate_table(n: int, pixel_size: int) -> np.ndarray:
    """int32 [n]: out[i] = in[table[i]] for the NEAREST down-then-up pixelate (crt_filter.py:580-583)."""
    small = max(1, n // int(pixel_size))
    return _nearest_index(small, n)[_nearest_index(n, small)].astype(np.int32)


# ------------------------------------------------------------------ glitch --
def glitch_geometry(variant: str, h: int, w: int, height_frac: float) -> Tuple[int, int, int, int]:
    """(y0, rows, seg_len, segments) of the glitch band (crt_filter.py:667-669, :843-844)."""
    y0 = max(0, min(h, h - int(h * height_frac)))
    if variant == "gui":
        return y0, h - y0, w, 1
    seg_len = max(8, min(32, w // 120 if w >= 120 else 8))
    return y0, h - y0, seg_len, (w + seg_len - 1) // seg_len


def glitch_offsets(variant: str, h: int, w: int, amp_px: int, height_frac: float, phase_px: float) -> Optional[np.ndarray]:
    """int32 [rows][segments] horizontal offsets drawn from numpy's PCG64 with the
    reference's seed and draw order, so an injected table reproduces the
    reference's glitch exactly (gui: crt_filter.py:670-679, export: :841-853)."""
    if not (amp_px > 0 and height_frac > 0.0):
        return None
    y0, rows, seg_len, nseg = glitch_geometry(variant, h, w, height_frac)
    if rows <= 0:
        return None
    ramp = np.arange(rows, dtype=F32) / max(1.0, float(rows))
    geom = (w << 10) + (h << 1)
    if variant == "gui":
        rng = np.random.default_rng((int(abs(float(phase_px)) * 0.05) + geom) & 0xFFFFFFFF)
        amp = np.asarray(float(amp_px) * np.exp(-3.0 * ramp), dtype=F32)
        jitter = np.clip(rng.normal(loc=0.0, scale=0.5, size=rows).astype(F32), -1.0, 1.0)
        jumps = rng.random(rows).astype(F32) < 0.03
        signs = rng.choice(np.array([-1.0, 1.0], dtype=F32), size=rows)
        jitter = jitter + jumps * signs
        return np.rint(np.clip(jitter * amp, -amp, amp)).astype(np.int32)[:, None]
    rng = np.random.default_rng((int(abs(float(phase_px)) * 2.0) + geom) & 0xFFFFFFFF)
    amp = float(amp_px) * (1.0 - ramp)
    per_segment = rng.standard_normal((rows, nseg)).astype(F32) * (amp[:, None] * 0.7)
    drift = np.clip(np.cumsum(rng.standard_normal(rows).astype(F32)) * 0.1, -amp * 0.4, amp * 0.4)
    return np.ascontiguousarray(np.rint(drift[:, None] + per_segment).astype(np.int32))


def noise_plane_shape(h: int, w: int, grain_size: int) -> Tuple[int, int]:
    """Shape of the N(0,1) plane the reference fills (crt_filter.py:637-645)."""
    if grain_size and grain_size > 1:
        return max(1, h // int(grain_size)), max(1, w // int(grain_size))
    return h, w
