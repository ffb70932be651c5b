"""The pure-numpy restatements of the OpenCV primitives (the specification the
CUDA kernels follow) against the installed cv2.  CPU only."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import cv_restated as R  # noqa: E402


def _rand(shape, seed, power=3):
    return (np.random.default_rng(seed).random(shape, dtype=np.float32) ** power).astype(np.float32)


@pytest.mark.parametrize("hw", [(64, 96), (480, 640), (62, 94), (63, 95), (101, 37)])
def test_resize_linear_down_up_bit_exact(hw):
    h, w = hw
    src = _rand((h, w, 3), 1)
    down = cv2.resize(src, (w // 2, h // 2), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(down, R.resize_linear(src, w // 2, h // 2))
    up = cv2.resize(down, (w, h), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(up, R.resize_linear(down, w, h))


def test_resize_linear_grain_plane():
    small = np.random.default_rng(2).standard_normal((54, 96)).astype(np.float32)
    assert np.array_equal(cv2.resize(small, (193, 109), interpolation=cv2.INTER_LINEAR), R.resize_linear(small, 193, 109))


@pytest.mark.parametrize("n,ps", [(12, 5), (13, 2), (480, 2), (641, 3), (1080, 7), (97, 16), (64, 1)])
def test_pixelate_index(n, ps):
    ramp = np.arange(n, dtype=np.float32)[None, :].repeat(2, 0)
    small = cv2.resize(ramp, (max(1, n // ps), 2), interpolation=cv2.INTER_NEAREST)
    back = cv2.resize(small, (n, 2), interpolation=cv2.INTER_NEAREST)
    assert np.array_equal(back[0].astype(np.int64), R.pixelate_index(n, ps))


def test_gaussian_kernel_bit_exact():
    for k in range(1, 63, 2):
        for s in (0.17, 0.3, 0.5, 0.8, 1.0, 1.2, 1.5, 2.0, 3.3, 4.0, 7.7, 10.0):
            assert np.array_equal(cv2.getGaussianKernel(k, s, cv2.CV_32F).ravel(), R.gaussian_kernel(k, s)), (k, s)


@pytest.mark.parametrize("k,sigma", [(3, 0.4), (5, 0.8), (7, 1.2), (9, 1.5), (25, 4.0), (61, 10.0)])
@pytest.mark.parametrize("hw", [(48, 80), (270, 480)])
def test_gaussian_blur_bit_exact(k, sigma, hw):
    src = _rand((*hw, 3), 3)
    ref = cv2.GaussianBlur(src, (k, k), sigmaX=sigma, sigmaY=sigma, borderType=cv2.BORDER_REPLICATE)
    assert np.array_equal(ref, R.gaussian_blur(src, k, k, sigma, sigma))


@pytest.mark.parametrize("k,sigma", [(3, 0.3), (5, 0.5), (7, 1.0), (25, 4.0)])
def test_triad_row_blur_bit_exact(k, sigma):
    w = 160
    m = np.zeros((4, w, 3), np.float32)
    for c in range(3):
        m[:, :, c] = np.float32(0.65) + np.float32(0.35) * ((np.arange(w) % 3) == c)
    ref = cv2.GaussianBlur(m, (k, 1), sigmaX=sigma, sigmaY=0, borderType=cv2.BORDER_REPLICATE)
    assert np.array_equal(ref, R.gaussian_blur(m, k, 1, sigma, 0))


@pytest.mark.parametrize("strength", [0.15, -0.3, 1.0])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_remap_bit_exact(strength, dtype):
    from oracle.crt_oracle import barrel_maps
    img = _rand((135, 240, 3), 4, power=1).astype(dtype)
    mx, my = barrel_maps(135, 240, strength)
    ref = cv2.remap(img, mx, my, interpolation=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    assert np.array_equal(ref, R.remap_bilinear_const0(img, mx, my))


def test_convert_scale_abs():
    v = _rand((96, 128, 3), 5, power=1)
    v[0, 0] = [0.5 / 255, 1.5 / 255, 2.5 / 255]
    assert np.array_equal(cv2.convertScaleAbs(v, alpha=255.0, beta=0), R.convert_scale_abs_255(v))
    v64 = v.astype(np.float64) * 1.0000001
    v64[0, 1, 0] = 0.531372540997836  # 135.4999979 in double, 135.5 after OpenCV narrows to float32
    assert np.array_equal(cv2.convertScaleAbs(v64, alpha=255.0, beta=0), R.convert_scale_abs_255(v64))


def test_add_weighted_within_one_ulp():
    a, b = _rand((64, 64, 3), 6, 1), _rand((64, 64, 3), 7, 1)
    ref = cv2.addWeighted(a, 0.2, b, 0.8, 0.0)
    got = R.add_weighted(a, 0.2, b, 0.8)
    assert np.max(np.abs(ref.view(np.int32) - got.view(np.int32))) <= 1
