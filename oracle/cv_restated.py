"""Pure-numpy restatements of the OpenCV primitives the reference chain calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference's stencil/gather
arithmetic is not in /root/reference: it lives in opencv-python-headless
(requirements.txt:6 `>=4.8.0`, installed 4.13.0, built with IPP 2022.2 and
AVX512/FMA dispatch).  Each function here restates one cv2 entry point at the
call site the reference uses and is checked bit-for-bit (or to the stated
bound) against the installed cv2 by tests/test_cv_restated.py.  These
restatements are the specification the CUDA kernels are written to.

Facts established empirically against cv2 4.13.0 (probe results, float32 data):
  * cv2.resize(INTER_LINEAR), incl. the exact 2x downscale used by fast bloom
    (crt_filter.py:606-607): source coordinate (dx+0.5)*(src/dst)-0.5 in
    double, floor, edge clamp with weight 0; weights cast to float32;
    horizontal pass first, then vertical, each tap pair combined as
    fma(q - p, w, p) in float32.                                  [bit-exact]
  * cv2.GaussianBlur float32 (crt_filter.py:234, :610): separable, row pass
    then column pass, BORDER_REPLICATE.  Row pass k>=7: s = x[0]*k[0];
    s = fma(x[i], k[i], s) left to right.  Row pass k==3:
    fma(x0, k0, (x-1 + x+1)*k1);  k==5: fma(x-2 + x+2, k2, fma(x0, k0,
    (x-1 + x+1)*k1)).  Column pass (all k): s = c*k0; s = fma(x+i + x-i, ki, s)
    for i = 1..r.  [bit-exact when W*channels is a multiple of the SIMD width,
    i.e. all the frame sizes in BASELINE.json; a scalar tail in the last
    columns of odd widths may differ by 1 ulp]
  * cv2.remap(INTER_LINEAR, BORDER_CONSTANT 0) (crt_filter.py:347): map*32
    rounded half-even to 1/32 px, float32 weight table, 4 taps, out-of-range
    taps contribute 0.                                            [<= 1 ulp-ish]
  * cv2.convertScaleAbs(alpha=255) (crt_filter.py:696, :1098):
    saturate_u8(round-half-even(|255*x|)).                        [bit-exact]
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def _fma32(a, b, c):
    """float32 fused multiply-add emulated through float64 (the product of two
    float32 values is exact in float64; the final double rounding differs from a
    true fma only in ~1e-9 of cases)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F32)


# ---------------------------------------------------------------- resize ----
def linear_coords(n_dst: int, n_src: int):
    """Source index pair and float32 weight for cv2.resize INTER_LINEAR."""
    scale = float(n_src) / float(n_dst)
    d = np.arange(n_dst, dtype=np.float64)
    fx = (d + 0.5) * scale - 0.5
    s0 = np.floor(fx).astype(np.int64)
    w = fx - s0
    lo = s0 < 0
    w[lo] = 0.0
    s0[lo] = 0
    hi = s0 >= n_src - 1
    w[hi] = 0.0
    s0[hi] = n_src - 1
    s1 = np.minimum(s0 + 1, n_src - 1)
    return s0, s1, w.astype(F32)


def resize_linear(src: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """cv2.resize(src, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR), float32."""
    src = np.asarray(src, F32)
    squeeze = src.ndim == 2
    if squeeze:
        src = src[:, :, None]
    h, w = src.shape[:2]
    x0, x1, wx = linear_coords(dst_w, w)
    y0, y1, wy = linear_coords(dst_h, h)
    p, q = src[:, x0], src[:, x1]
    rows = _fma32(q - p, wx[None, :, None], p)
    p, q = rows[y0], rows[y1]
    out = _fma32(q - p, wy[:, None, None], p)
    return out[:, :, 0] if squeeze else out


def nearest_index(n_dst: int, n_src: int) -> np.ndarray:
    """Source index for cv2.resize INTER_NEAREST along one axis:
    min(floor(d * (1 / (n_dst / n_src))), n_src - 1)."""
    inv = 1.0 / (float(n_dst) / float(n_src))
    idx = np.floor(np.arange(n_dst, dtype=np.float64) * inv).astype(np.int64)
    return np.minimum(idx, n_src - 1)


def pixelate_index(n: int, pixel_size: int) -> np.ndarray:
    """Composite index of the reference's NEAREST down-then-up pixelate
    (crt_filter.py:578-584): out[i] = in[table[i]]."""
    small = max(1, n // int(pixel_size))
    down = nearest_index(small, n)   # small index -> full index
    up = nearest_index(n, small)     # full index  -> small index
    return down[up]


# ------------------------------------------------------------- gaussian ----
def gaussian_kernel(ksize: int, sigma: float) -> np.ndarray:
    """cv2.getGaussianKernel(ksize, sigma, CV_32F).ravel() for sigma > 0.

    OpenCV evaluates exp(-x^2 / (2 sigma^2)) for the left half in double
    (softfloat), sums 2*half + centre, multiplies by the reciprocal of the sum
    and casts to float32."""
    n = int(ksize)
    sigma = float(sigma)
    scale2x = -0.5 / (sigma * sigma)
    half = (n - 1) // 2
    x = np.arange(1 - n, 1 - n + 2 * half, 2, dtype=np.float64)  # 2*(i - (n-1)/2)
    vals = np.exp(x * x * scale2x * 0.25)
    total = 0.0
    for v in vals:
        total += float(v)
    total = total * 2.0 + 1.0
    if n % 2 == 0:
        total += 1.0
    mul = 1.0 / total
    out = np.empty(n, np.float64)
    out[:half] = vals * mul
    out[n - half:] = (vals * mul)[::-1]
    out[half] = mul
    if n % 2 == 0:
        out[half + 1] = mul
    return out.astype(F32)


def _replicate(img, r, axis):
    idx = np.clip(np.arange(-r, img.shape[axis] + r), 0, img.shape[axis] - 1)
    return np.take(img, idx, axis=axis)


def _tap(padded, k, n, axis):
    sl = [slice(None)] * padded.ndim
    sl[axis] = slice(k, k + n)
    return padded[tuple(sl)]


def gaussian_row_pass(img: np.ndarray, kern: np.ndarray) -> np.ndarray:
    """Horizontal pass of cv2.GaussianBlur (float32), see module docstring."""
    K = len(kern)
    r = K // 2
    n = img.shape[1]
    p = _replicate(img, r, 1)
    X = lambda k: _tap(p, k, n, 1)  # noqa: E731
    if K == 1:
        return (img * kern[0]).astype(F32)
    if K == 3:
        return _fma32(X(1), kern[1], (X(0) + X(2)) * kern[2])
    if K == 5:
        inner = _fma32(X(2), kern[2], (X(1) + X(3)) * kern[3])
        return _fma32(X(4) + X(0), kern[4], inner)
    s = X(0) * kern[0]
    for k in range(1, K):
        s = _fma32(X(k), kern[k], s)
    return s


def gaussian_col_pass(img: np.ndarray, kern: np.ndarray) -> np.ndarray:
    """Vertical pass of cv2.GaussianBlur (float32), see module docstring."""
    K = len(kern)
    r = K // 2
    n = img.shape[0]
    p = _replicate(img, r, 0)
    X = lambda k: _tap(p, k, n, 0)  # noqa: E731
    s = X(r) * kern[r]
    for k in range(1, r + 1):
        s = _fma32(X(r + k) + X(r - k), kern[r + k], s)
    return s


def gaussian_blur(img: np.ndarray, kx: int, ky: int, sigma_x: float, sigma_y: float) -> np.ndarray:
    """cv2.GaussianBlur(img, (kx, ky), sigmaX, sigmaY, BORDER_REPLICATE), float32."""
    img = np.asarray(img, F32)
    if sigma_y <= 0:
        sigma_y = sigma_x
    out = gaussian_row_pass(img, gaussian_kernel(kx, sigma_x)) if kx > 1 else img
    if ky > 1:
        out = gaussian_col_pass(out, gaussian_kernel(ky, sigma_y))
    return out


# ---------------------------------------------------------------- remap ----
def remap_bilinear_const0(img: np.ndarray, map_x: np.ndarray, map_y: np.ndarray) -> np.ndarray:
    """cv2.remap(img, map_x, map_y, INTER_LINEAR, BORDER_CONSTANT, 0).

    Coordinates are quantised to 1/32 px (round-half-even of map*32), weights
    come from a float32 table (1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy*fx, and
    every tap outside the image contributes 0."""
    h, w = img.shape[:2]
    sx = np.rint(map_x.astype(F32) * F32(32)).astype(np.int64)
    sy = np.rint(map_y.astype(F32) * F32(32)).astype(np.int64)
    ix, iy = sx >> 5, sy >> 5
    fx = ((sx & 31).astype(F32) * F32(1.0 / 32.0))
    fy = ((sy & 31).astype(F32) * F32(1.0 / 32.0))
    one = F32(1)
    w00 = (one - fy) * (one - fx)
    w01 = (one - fy) * fx
    w10 = fy * (one - fx)
    w11 = fy * fx
    wt = img.dtype if img.dtype == np.float64 else F32

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        v = img[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)].astype(wt)
        return v * ok[:, :, None]

    out = (tap(iy, ix) * w00[:, :, None].astype(wt) + tap(iy, ix + 1) * w01[:, :, None].astype(wt)
           + tap(iy + 1, ix) * w10[:, :, None].astype(wt) + tap(iy + 1, ix + 1) * w11[:, :, None].astype(wt))
    return out.astype(img.dtype)


# ------------------------------------------------------- scale / blend ----
def convert_scale_abs_255(img: np.ndarray) -> np.ndarray:
    """cv2.convertScaleAbs(img, alpha=255.0, beta=0): |255*x| in float32,
    round-half-even, saturate to uint8.  float64 inputs are first narrowed to
    float32 by OpenCV's vector path (observed: 0.531372540997836 -> 136, not
    135), which covers every element when W*3 is a multiple of the SIMD width;
    the scalar tail of odd widths stays in double."""
    v = np.abs(np.asarray(img).astype(F32) * F32(255))
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def add_weighted(a: np.ndarray, alpha: float, b: np.ndarray, beta: float) -> np.ndarray:
    """cv2.addWeighted(a, alpha, b, beta, 0) for float images (crt_filter.py:693).

    OpenCV keeps alpha/beta in double and evaluates a*alpha + b*beta in double
    (with an fma), rounding once to the image dtype: this restatement is within
    1 ulp of cv2, not bit-exact (the stage is after the triad LUT, where only
    the final +-1 LSB matters)."""
    out = np.asarray(a, np.float64) * float(alpha) + np.asarray(b, np.float64) * float(beta)
    return out.astype(a.dtype)
