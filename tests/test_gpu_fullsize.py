"""BASELINE.json configurations at their full frame sizes on the B200: direct
comparison with the oracle on a couple of frames (the CPU path needs ~1-4 s per
frame there), plus size-independent properties."""
import numpy as np
import pytest

from oracle import crt_oracle as O
from oracle import harness
from oracle.cases import BASE, GAUSS, GRADE, LIVE, WARP, Case

pytestmark = pytest.mark.gpu

FULL = [
    Case("full_cfg1_vga", 480, 640, BASE, frames=4),
    Case("full_cfg2_1080p", 1080, 1920, BASE.but(**GAUSS, **GRADE), frames=2),
    Case("full_default_720p", 720, 1280, BASE, frames=3),           # smallest frame class that takes the TMA-pipelined kernel
    Case("full_default_4k", 2160, 3840, BASE, frames=3),           # north_star chain: TMA-pipelined block kernel from frame 1 on
    Case("full_default_4k_thr", 2160, 3840, BASE.but(bloom_threshold=0.5, flicker_strength=0.3, flicker_hz=50.0), frames=3, source="structured"),
    Case("full_cfg3_4k", 2160, 3840, BASE.but(**WARP), frames=2),
    Case("full_cfg4_4k", 2160, 3840, BASE.but(**GAUSS, **GRADE, **WARP, **LIVE), frames=2, fps=60.0, first_index=100),
    # BASELINE configs[4] at its real size: K = 25 taps at 7680 x 4320 (tile counts, the < 2^31 offset check of plan_fused,
    # the 398 MB pre-warp scratch of the two-pass path).  One frame: the CPU path needs ~20 s for it.
    Case("full_cfg5_8k", 4320, 7680, BASE.but(**GRADE, **WARP, **LIVE, fast_bloom=False, bloom_sigma=4.0, bloom_strength=0.3), frames=1, first_index=3),
]


@pytest.mark.parametrize("case", FULL, ids=lambda c: c.name)
def test_full_size_matches_oracle(case):
    from gpu_util import log_report, run_case_gpu
    got, _, fused = run_case_gpu(case, "export")
    want, _ = harness.run_oracle(case, "export", backend="cv2")
    for a, b in zip(want, got):
        st = harness.diff_stats(a, b)
        log_report(case=case.name, what=f"fullsize/fused={fused}", **st)
        assert st["psnr"] >= 50.0 and st["max"] <= 1, st


def test_identity_4k_bit_exact():
    import torch
    from pythoncrt_b200 import CrtEngine, CrtParams
    p = CrtParams(scanline_strength=0.0, triad_strength=0.0, aberration_px=0, pixel_size=1, bloom_strength=0.0,
                  vignette_strength=0.0, persistence=0.0, noise_strength=0.0)
    eng = CrtEngine(3840, 2160).configure(p)
    g = torch.Generator(device="cuda").manual_seed(7)
    frames = torch.randint(0, 256, (3, 2160, 3840, 3), dtype=torch.uint8, device="cuda", generator=g)
    out, _ = eng.process(frames)
    assert torch.equal(out, frames)


@pytest.mark.parametrize("hw", [(1080, 1920), (2160, 3840)])
def test_temporal_shards_match_serial(hw):
    """a chunk started `halo` frames early from an empty state equals
    the serial run to within p^halo (<= 1 LSB, in practice identical)."""
    import torch
    from pythoncrt_b200 import CrtEngine, CrtParams, clip
    h, w = hw
    p = CrtParams(noise_strength=0.0, warp_strength=0.15, scanline_angle=3.0, scanline_thickness=1.2)
    eng = CrtEngine(w, h).configure(p)
    n = 24
    g = torch.Generator(device="cuda").manual_seed(11)
    frames = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
    serial, _ = clip.process_clip(eng, frames, fps=30.0)
    for rank in range(3):
        def run_range(first, last, fresh):
            out, _ = clip.process_clip(eng, frames[first:last], fps=30.0, first_index=first)
            return out
        mine, (a, b) = clip.process_clip_sharded(run_range, n, p, rank, 3)
        d = (mine.to(torch.int16) - serial[a:b].to(torch.int16)).abs()
        assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 1e-3
