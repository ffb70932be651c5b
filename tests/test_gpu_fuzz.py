"""Randomised parity on the B200: seeded random parameter sets and frame sizes through the CUDA path
(whatever kernel the planner picks) against the oracle.  Sizes and parameters are drawn so that every
kernel family is hit: block kernels (pixel_size 2, even sizes, fast / gaussian bloom with each
instantiated tap count), general tile kernels (other pixel sizes, odd sizes), two-pass (warp, glitch)
and the staged fallback (unsupported tap counts).  Same bar as test_gpu_parity."""
import numpy as np
import pytest

from gpu_util import log_report, run_case_gpu
from oracle import harness
from oracle.cases import BASE, Case

pytestmark = pytest.mark.gpu

SIGMAS = [0.7, 1.0, 1.2, 1.5, 1.7, 2.0, 4.0, 0.4, 3.0]       # k = 5, 7, 9, 9, 11, 13, 25, 3 (staged), 19 (staged)


def _random_case(i: int) -> Case:
    rng = np.random.default_rng(1000 + i)
    pick = lambda *a: a[int(rng.integers(len(a)))]
    u = lambda lo, hi: float(rng.uniform(lo, hi))
    w = int(pick(128, 136, 192, 200, 256, 101, 324))
    h = int(pick(96, 64, 130, 75, 150))
    over = dict(
        pixel_size=pick(2, 2, 2, 1, 3),
        aberration_px=int(pick(1, 0, -1, 2, -3, 5, 8)),
        scanline_strength=pick(0.6, 0.0, 0.9), scanline_period_px=pick(2.0, 3.0, 2.5), scanline_speed_px_s=pick(30.0, 0.0, 47.0),
        triad_strength=pick(0.35, 0.0, 0.8), triad_softness=pick(0.5, 0.0, 1.0), triad_gamma=pick(2.2, 1.0, 1.8),
        triad_preserve_luma=bool(pick(False, False, True)),
        vignette_strength=pick(0.25, 0.0, 0.6), persistence=pick(0.2, 0.0, 0.5),
        bloom_strength=pick(0.25, 0.0, 0.5), bloom_threshold=pick(0.0, 0.0, 0.5, 0.7),
        fast_bloom=bool(pick(True, False, False)), bloom_sigma=pick(*SIGMAS),
    )
    if rng.random() < 0.4:
        over.update(brightness=u(-0.1, 0.1), contrast=u(0.8, 1.3), saturation=u(0.7, 1.4), temperature=u(-0.3, 0.3),
                    gamma=pick(1.0, 1.1, 0.8, 2.2))
    if rng.random() < 0.25:
        over.update(warp_strength=pick(0.15, -0.2, 0.4))
    if rng.random() < 0.25:
        over.update(scanline_angle=pick(3.0, -10.0), scanline_thickness=pick(1.0, 1.2, 0.6))
    if rng.random() < 0.2:
        over.update(noise_strength=u(1.0, 6.0), grain_size=int(pick(1, 2, 3)))
    if rng.random() < 0.2:
        over.update(flicker_strength=u(0.1, 0.8), flicker_hz=pick(60.0, 7.0))
    if rng.random() < 0.2:
        over.update(glitch_amp_px=int(pick(8, 24)), glitch_height_frac=pick(0.25, 0.5))
    return Case(f"fuzz{i}", h, w, BASE.but(**over), frames=3, fps=pick(30.0, 60.0), source=pick("noise", "structured"),
                first_index=int(rng.integers(0, 50)))


@pytest.mark.parametrize("i", range(48))
def test_random_parameters_match_oracle(i):
    case = _random_case(i)
    variant = "export" if i % 2 else "gui"
    got, state, fused = run_case_gpu(case, variant, "auto")
    want, want_state = harness.run_oracle(case, variant, backend="cv2")
    worst = {"max": 0, "frac_gt1": 0.0, "frac_ne": 0.0, "psnr": float("inf")}
    for a, b in zip(want, got):
        st = harness.diff_stats(a, b)
        worst = {"max": max(worst["max"], st["max"]), "frac_gt1": max(worst["frac_gt1"], st["frac_gt1"]),
                 "frac_ne": max(worst["frac_ne"], st["frac_ne"]), "psnr": min(worst["psnr"], st["psnr"])}
    log_report(case=case.name, what=f"{variant}/auto/fused={fused}", w=case.w, h=case.h, **worst)
    assert worst["psnr"] >= 50.0, (worst, vars(case.params))
    assert worst["max"] <= 1, (worst, vars(case.params))
    assert worst["frac_ne"] <= 5e-3, (worst, vars(case.params))
    d = np.abs(want_state.astype(np.float64) - state)
    assert d.max() < 1.0 / 255 and (d > 4e-6).mean() < 4e-3, (d.max(), (d > 4e-6).mean())
