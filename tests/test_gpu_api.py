"""Reference-facing API on the B200: drop-in functions, host-buffer path, generators."""
import numpy as np
import pytest

from oracle import crt_oracle as O
from oracle import harness
from oracle.cases import CASES_BY_NAME, case_frames

pytestmark = pytest.mark.gpu


def _args(p, tri, vig, phase):
    return (p.scanline_strength, tri, float(p.triad_gamma), bool(p.triad_preserve_luma), int(p.aberration_px), p.bloom_sigma,
            p.bloom_strength, float(p.bloom_threshold), p.noise_strength, vig)


def _kw(p, tsec):
    return dict(time_sec=tsec, brightness=p.brightness, contrast=p.contrast, gamma=p.gamma, saturation=p.saturation,
                temperature=p.temperature, flicker_strength=p.flicker_strength, flicker_hz=p.flicker_hz, grain_size=p.grain_size,
                scanline_angle=p.scanline_angle, scanline_thickness=p.scanline_thickness, warp_strength=p.warp_strength)


@pytest.mark.parametrize("name", ["cfg1_cli_default", "cfg3_warp", "glitch_big", "noise_grain5"])
def test_apply_crt_effect_drop_in(name):
    """Same call the GUI makes (crt_filter.py:1810-1852), masks built by the reference-compatible constructors."""
    import pythoncrt_b200 as crt
    case = CASES_BY_NAME[name]
    p = case.params
    tri = crt.make_triad_mask(case.h, case.w, p.triad_strength, p.triad_softness)
    vig = O.vignette_mask(case.h, case.w, p.vignette_strength)      # a foreign (reference-style) float64 mask
    want, _ = harness.run_oracle(case, "gui")
    planes = harness.noise_planes(case)
    state = None
    for j, frame in enumerate(case_frames(case)):
        frame.setflags(write=False)                                   # the reference hands out read-only frames (:501)
        phase, tsec = harness.frame_scalars(case, j)
        out, state = crt.apply_crt_effect(frame, *_args(p, tri, vig, phase), p.persistence, state, p.scanline_period_px, phase,
                                          p.fast_bloom, int(p.pixel_size), int(p.glitch_amp_px), float(p.glitch_height_frac),
                                          noise_plane=None if planes is None else planes[j], **_kw(p, tsec))
        st = harness.diff_stats(want[j], out)
        assert out.dtype == np.uint8 and out.shape == frame.shape and st["max"] <= 1 and st["psnr"] >= 50, st
    assert state.shape == (case.h, case.w, 3) and np.asarray(state).dtype == np.float32


@pytest.mark.parametrize("name", ["cfg3_warp", "cfg1_cli_default", "cfg2_gauss_grade", "glitch_big", "text_after_warp", "odd_size_gauss", "gauss_k61"])
def test_apply_static_effects_drop_in(name):
    """apply_static_effects (:702-861) = crt_process_static; every kernel family (block kernels through q_out, two-pass
    with the gather writing the float image, general tile kernels, staged fallback) against the oracle's float image."""
    import pythoncrt_b200 as crt
    from oracle.cases import case_text_layer
    case = CASES_BY_NAME[name]
    p = case.params
    tri = crt.make_triad_mask(case.h, case.w, p.triad_strength, p.triad_softness)
    vig = crt.make_vignette(case.h, case.w, p.vignette_strength)
    frame = case_frames(case)[0]
    phase, tsec = harness.frame_scalars(case, 0)
    text = case_text_layer(case)
    img = crt.apply_static_effects(frame, *_args(p, tri, vig, phase), p.scanline_period_px, phase, p.fast_bloom, int(p.pixel_size),
                                   int(p.glitch_amp_px), float(p.glitch_height_frac), text_overlay_rgba=text,
                                   text_overlay_after=(case.text != "before"), **_kw(p, tsec))
    ref = O.static_chain(frame, p, phase_px=phase, time_sec=tsec, variant="export", text_rgba=text, text_after=(case.text != "before"))
    d = np.abs(ref - img)
    # float32 against the reference's float64 tail; a sample may sit on the other side of a LUT bin after the colour gamma
    assert img.dtype == np.float32 and d.max() < 1.0 / 255 and (d > 4e-6).mean() < 2e-3, (d.max(), (d > 4e-6).mean())


def test_host_buffer_path_equals_device_path():
    import torch
    from gpu_util import run_case_gpu
    import host_emu
    from pythoncrt_b200 import CrtEngine
    case = CASES_BY_NAME["vga_cfg1"]
    want, _, _ = run_case_gpu(case, "export")
    frames = np.stack(case_frames(case) * 40)                        # 80 frames: several ring chunks
    eng = CrtEngine(case.w, case.h).configure(host_emu.oracle_to_product_params(case.params))
    pinned = torch.from_numpy(frames).pin_memory()
    out = eng.process_host(pinned.numpy(), fps=case.fps)
    assert np.array_equal(out[0], want[0])
    dev, _ = eng.process(torch.from_numpy(frames).cuda(), fps=case.fps)
    assert np.array_equal(out, dev.cpu().numpy())
    eng.reset_state()
    assert np.array_equal(eng.process_host(frames[:2], fps=case.fps), out[:2])   # pageable memory works too


def test_rawvideo_stream_equals_device_path():
    """clip.process_rawvideo (rgb24 pipe format of the reference, crt_filter.py:484-502 / :1101) against
    the device-resident path, chunked so that the state crosses chunk boundaries."""
    import io
    import torch
    import host_emu
    from pythoncrt_b200 import CrtEngine, clip
    case = CASES_BY_NAME["cfg1_cli_default"]
    frames = np.stack(case_frames(case) * 4)                         # 12 frames
    eng = CrtEngine(case.w, case.h).configure(host_emu.oracle_to_product_params(case.params))
    dev, _ = eng.process(torch.from_numpy(frames).cuda(), fps=case.fps)
    dst = io.BytesIO()
    n = clip.process_rawvideo(eng, io.BytesIO(frames.tobytes() + b"\x00" * 100), dst, fps=case.fps, chunk_frames=5)
    assert n == 12
    got = np.frombuffer(dst.getvalue(), np.uint8).reshape(frames.shape)
    assert np.array_equal(got, dev.cpu().numpy())


def test_device_generators():
    import torch
    from pythoncrt_b200 import CrtEngine, CrtParams
    p = CrtParams(noise_strength=3.0, grain_size=1, glitch_amp_px=16, glitch_height_frac=0.25)
    eng = CrtEngine(640, 480).configure(p, noise_mode="generate", glitch_mode="generate", seed=123)
    a, b, a2 = eng.generate_noise(5), eng.generate_noise(6), eng.generate_noise(5)
    assert torch.equal(a, a2) and not torch.equal(a, b)
    assert abs(float(a.mean())) < 0.01 and abs(float(a.std()) - 1.0) < 0.01 and float(a.abs().max()) < 6.5
    offs = eng.generate_glitch(9)
    assert offs.shape == (120, 80) and int(offs.abs().max()) <= 16 * 3 and float(offs.float().std()) > 1.0
    g = torch.Generator(device="cuda").manual_seed(3)
    frames = torch.randint(0, 256, (3, 480, 640, 3), dtype=torch.uint8, device="cuda", generator=g)
    noisy, _ = eng.process(frames)
    clean, _ = CrtEngine(640, 480).configure(p.but(noise_strength=0.0, glitch_amp_px=0)).process(frames)
    top = (noisy[:, :300].float() - clean[:, :300].float())
    assert 0.5 < float(top.std()) < 4.0 and abs(float(top.mean())) < 0.5    # ~N(0, 3/255) before masks and persistence
    assert not torch.equal(noisy[:, 400:], clean[:, 400:])


def test_unaligned_buffers_are_rejected_not_faulted():
    """crt_process validates buffer alignment (16-byte state, 4-byte output) instead of faulting in a kernel."""
    import torch
    from pythoncrt_b200 import CrtEngine, CrtParams
    from pythoncrt_b200.cabi import CrtError
    eng = CrtEngine(128, 96).configure(CrtParams(noise_strength=0.0))
    frames = torch.randint(0, 256, (2, 96, 128, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty_like(frames)
    raw = torch.zeros(96 * 128 * 3 + 4, dtype=torch.float32, device="cuda")
    bad_state = raw[1:1 + 96 * 128 * 3].view(96, 128, 3)             # 4-byte offset: not 16-byte aligned
    assert bad_state.data_ptr() % 16 != 0
    with pytest.raises(CrtError, match="aligned"):
        eng.process(frames, out, state=bad_state, state_valid=False)
    good, _ = eng.process(frames)                                     # the engine is still usable
    assert good.shape == frames.shape
