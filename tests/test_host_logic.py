"""Host-side logic of the product package: parameter surface, table builders,
shard planning, and that the C-ABI library loads and exports every symbol the
header declares.  CPU only (no compute calls)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from pythoncrt_b200 import cabi, clip, tables
from pythoncrt_b200.config import build_config
from pythoncrt_b200.params import PRESET_IGNORED, PRESET_KEYS, CrtParams

cv2 = pytest.importorskip("cv2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ params --
def test_cli_defaults_and_clamps():
    d = CrtParams.from_cli([])
    assert d == CrtParams()                                   # crt_filter.py:1160-1205
    p = CrtParams.from_cli("--scanline-strength 3 --triad-gamma 0.01 --aberration-px 99 --persistence 2 --pixel-size 0 "
                           "--gamma 0 --temperature -7 --warp-strength 5 --glitch-height 9 --scanline-period 0.2 "
                           "--no-fast-bloom --triad-preserve-luma --input x.mp4 --crf 20".split())
    assert (p.scanline_strength, p.triad_gamma, p.aberration_px, p.persistence, p.pixel_size) == (1.0, 0.1, 8, 0.95, 1)
    assert (p.gamma, p.temperature, p.warp_strength, p.glitch_height_frac, p.scanline_period_px) == (1e-3, -1.0, 1.0, 1.0, 1.0)
    assert p.fast_bloom is False and p.triad_preserve_luma is True


def test_preset_round_trip_and_ignored_keys():
    p = CrtParams.gui_defaults().but(warp_strength=0.15, glitch_amp_px=16, fast_bloom=False, grain_size=3)
    blob = p.to_preset()
    assert set(blob) == set(PRESET_KEYS)                      # crt_filter.py:2044-2080
    blob.update({k: 1 for k in PRESET_IGNORED})
    blob["unknown_key"] = "x"
    assert CrtParams.from_preset(json.loads(json.dumps(blob))) == p
    assert CrtParams.from_preset({"scanline": 0.1}).scanline_strength == 0.1
    assert CrtParams.from_preset("not a dict") == CrtParams.gui_defaults()
    assert CrtParams.gui_defaults().triad_preserve_luma and CrtParams.gui_defaults().scanline_speed_px_s == 60.0


def test_frame_scalars():
    p = CrtParams(scanline_speed_px_s=30.0)
    assert p.phase_px(45, 30) == (45 / 30.0) * 30.0 and p.time_sec(45, 60) == 45 / 60.0


# ------------------------------------------------------------------ tables --
@pytest.mark.parametrize("w,strength,soft", [(128, 0.35, 0.5), (640, 0.35, 0.5), (1920, 1.0, 0.0), (160, 0.6, 1.0), (3840, 0.2, 4.0), (101, 0.35, 0.5)])
def test_triad_columns_match_reference_mask(w, strength, soft):
    from oracle.crt_oracle import triad_mask
    ref = triad_mask(3, w, strength, soft, backend="cv2")
    assert np.array_equal(ref[0], ref[2])
    got = tables.triad_columns(w, strength, soft)
    if (w * 3) % 16 == 0:
        assert np.array_equal(got, ref[0])
    else:  # OpenCV's scalar SIMD tail on odd widths may differ in the last bit
        assert np.max(np.abs(got - ref[0])) <= 1.2e-7
    m = tables.make_triad_mask(5, w, strength, soft)
    assert m.shape == (5, w, 3) and m.dtype == np.float32 and np.array_equal(np.asarray(m)[4], got)


def test_gaussian_taps_and_ksize():
    for s in (0.17, 0.4, 0.7, 1.2, 1.5, 4.0, 10.0):
        k = tables.bloom_ksize(s)
        assert np.array_equal(tables.gaussian_taps(k, s), cv2.getGaussianKernel(k, s, cv2.CV_32F).ravel())
    assert [tables.bloom_ksize(s) for s in (1.2, 1.5, 4.0, 10.0, 0.1)] == [9, 9, 25, 61, 1]


def test_luts_are_numpys_own_power():
    fwd, inv = tables.triad_luts(2.2)
    x = np.linspace(0.0, 1.0, 1025, dtype=np.float32)
    assert np.array_equal(fwd, np.power(x, 2.2, dtype=np.float32)) and np.array_equal(inv, np.power(x, 1 / 2.2, dtype=np.float32))


@pytest.mark.parametrize("n,ps", [(12, 5), (480, 2), (641, 3), (1080, 7), (64, 1)])
def test_pixelate_table(n, ps):
    ramp = np.arange(n, dtype=np.float32)[None, :].repeat(2, 0)
    small = cv2.resize(ramp, (max(1, n // ps), 2), interpolation=cv2.INTER_NEAREST)
    assert np.array_equal(cv2.resize(small, (n, 2), interpolation=cv2.INTER_NEAREST)[0].astype(np.int32), tables.pixelate_table(n, ps))


@pytest.mark.parametrize("variant", ["gui", "export"])
@pytest.mark.parametrize("h,w,amp,frac,phase", [(96, 128, 16, 0.25, 17.0), (480, 640, 64, 0.6, 3.3), (75, 101, 8, 1.0, 900.0)])
def test_glitch_offsets_equal_oracle_draws(variant, h, w, amp, frac, phase):
    from oracle.crt_oracle import glitch_table
    ref = glitch_table(variant, h, w, amp, frac, phase)
    got = tables.glitch_offsets(variant, h, w, amp, frac, phase)
    assert np.array_equal(ref.offs, got)
    assert tables.glitch_geometry(variant, h, w, frac) == (ref.y0, h - ref.y0, ref.seg_len, ref.offs.shape[1])
    assert tables.glitch_offsets(variant, h, w, 0, frac, phase) is None


def test_vignette_drop_in_and_inference():
    from oracle.crt_oracle import vignette_mask
    v = tables.make_vignette(48, 64, 0.25)
    assert v.dtype == np.float64 and np.array_equal(np.asarray(v), vignette_mask(48, 64, 0.25))
    assert tables.infer_vignette_strength(v) == 0.25
    assert abs(tables.infer_vignette_strength(vignette_mask(48, 64, 0.4)) - 0.4) < 1e-12     # a foreign (reference-made) mask
    assert tables.infer_vignette_strength(np.random.default_rng(0).random((48, 64))) is None


def test_build_config_tables():
    p = CrtParams(fast_bloom=False, bloom_sigma=1.5, noise_strength=0.0)
    c, tabs = build_config(p, 128, 96)
    assert c.triad_on == 1 and c.vignette_on == 1 and c.vignette_strength == 0.25 and c.pixel_size == 2
    assert set(tabs) == {cabi.TABLE_TRIAD_COLS, cabi.TABLE_LUT_FWD, cabi.TABLE_LUT_INV, cabi.TABLE_GAUSS_TAPS,
                         cabi.TABLE_PIXELATE_X, cabi.TABLE_PIXELATE_Y}
    assert tabs[cabi.TABLE_GAUSS_TAPS].shape == (9,)
    c2, tabs2 = build_config(p.but(triad_gamma=1.0), 128, 96, triad_cols=None, vignette=np.random.default_rng(1).random((96, 128)))
    assert c2.triad_on == 0 and c2.vignette_on == 2 and cabi.TABLE_VIGNETTE_PLANE in tabs2
    with pytest.raises(ValueError):
        build_config(p, 128, 96, text_rgba=np.zeros((4, 4, 4), np.uint8))


# --------------------------------------------------------------- sharding --
def test_halo_and_shard_plan():
    assert [clip.halo_frames(p) for p in (0.0, 0.2, 0.5, 0.8, 0.9, 0.95)] == [0, 5, 11, 35, 73, 149]
    for n, world in [(7200, 8), (600, 4), (10, 3), (3, 8)]:
        spans = [clip.shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert clip.shard_plan(7200, 0, 8, 0.2) == (0, 0, 900) and clip.shard_plan(7200, 3, 8, 0.2) == (2695, 2700, 3600)


# ------------------------------------------------------------------ C ABI --
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "crt_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char\*)\s+(crt_\w+)\s*\(", header, flags=re.M))
    assert declared == set(cabi.EXPORTS), declared ^ set(cabi.EXPORTS)
    if not os.path.isfile(cabi.LIB_PATH):
        from pythoncrt_b200 import build
        build.build()
    lib = C.CDLL(cabi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.crt_abi_version() == cabi.ABI_VERSION


def test_struct_layout_matches_the_c_compiler():
    import host_emu
    L = host_emu.lib()
    assert L.emu_sizeof_params() == C.sizeof(cabi.CrtParamsC)
    assert L.emu_sizeof_frame() == C.sizeof(cabi.CrtFrameC)


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pythoncrt_b200.engine import CrtEngine
    with pytest.raises(cabi.CrtError):
        CrtEngine(64, 48)
    lib = cabi.load_library()
    h = C.c_void_p()
    assert lib.crt_create(0, 64, 48, C.byref(h)) == 3         # CRT_ERR_NO_DEVICE


# ------------------------------------------------------------ tile planner --
def test_fused_tile_planner_on_every_case():
    """Host-side planner of the fused kernel (csrc/crt_fused.cuh plan_fused): must never touch
    device memory, must fit shared memory, and must hand glitch / warp+gaussian to the staged path."""
    import host_emu
    from oracle.cases import CASES
    for case in CASES:
        p = host_emu.oracle_to_product_params(case.params)
        pl = host_emu.plan(p, case.w, case.h)
        glitch = p.glitch_amp_px > 0 and p.glitch_height_frac > 0
        warp_gauss = p.warp_strength != 0 and not p.fast_bloom and p.bloom_strength > 0 and p.bloom_sigma > 0
        assert pl["ok"] == (not glitch and not warp_gauss), (case.name, pl)
        if pl["ok"]:
            assert pl["smem"] <= 200 * 1024 and pl["th"] in (16, 32, 64)
            gauss = (not p.fast_bloom) and p.bloom_strength > 0 and p.bloom_sigma > 0
            if p.warp_strength == 0 and not gauss and pl["cap_px"] > 0:      # without the warp the region is at least the tile itself
                assert pl["cap_px"] >= min(64, case.w) * min(pl["th"], case.h)   # (the packed gaussian and pixel_size-2 block kernels size their own buffers: cap_px == 0)
    big = host_emu.plan(CrtParams(noise_strength=0.0, warp_strength=0.15, scanline_angle=3.0), 3840, 2160)
    assert big["ok"] and big["smem"] <= 110 * 1024          # BASELINE configs[2] keeps two CTAs per SM


def test_rawvideo_streaming_chunks_and_short_read():
    """clip.process_rawvideo: rgb24 framing as the reference's pipe reader (:494-502) — a trailing
    partial frame is dropped, chunks carry consecutive frame indices, every frame is written once."""
    import io
    from pythoncrt_b200 import clip

    class FakeEngine:
        height, width = 4, 6
        def __init__(self): self.calls = []; self.resets = 0
        def reset_state(self): self.resets += 1
        def process_host(self, a, out, fps, first_index):
            self.calls.append((a.shape[0], first_index, fps))
            out[...] = 255 - a
            return out

    fb = 4 * 6 * 3
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256, 7 * fb + 11, dtype=np.uint8).tobytes()        # 7 frames + a partial one
    frames = list(clip.iter_rawvideo(io.BytesIO(data), 6, 4))
    assert len(frames) == 7 and frames[3].shape == (4, 6, 3)
    eng, dst = FakeEngine(), io.BytesIO()
    n = clip.process_rawvideo(eng, io.BytesIO(data), dst, fps=24.0, first_index=10, chunk_frames=3, pinned=False)
    assert n == 7 and eng.resets == 1
    assert eng.calls == [(3, 10, 24.0), (3, 13, 24.0), (1, 16, 24.0)]
    want = (255 - np.frombuffer(data[:7 * fb], np.uint8)).tobytes()
    assert dst.getvalue() == want
