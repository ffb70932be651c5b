"""Boundary behaviour on the B200 beyond per-case parity: channel order, thread safety of the C ABI,
distribution of the device generators against the reference's own draws, long persistence halos,
and the patched export pipeline (effects.install) end to end."""
import threading
import types

import numpy as np
import pytest

from gpu_util import run_case_gpu
from oracle import crt_oracle as O
from oracle import harness
from oracle.cases import BASE, CASES_BY_NAME, GAUSS, GRADE, LIVE, WARP, Case, case_frames

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------------------------------- channel order --
@pytest.mark.parametrize("name", ["cfg1_cli_default", "cfg2_gauss_grade", "cfg4_full", "cfg1_gui_default", "text_after_warp", "odd_size"])
def test_bgr_frames_equal_swapped_rgb(name):
    """crt_params.channel_order = BGR: same bytes as the RGB run on the channel-swapped clip, swapped back
    (the reference's index rules :289, :296-297, :573-575 follow the colour, not the index)."""
    import torch
    from pythoncrt_b200.engine import CrtEngine
    import host_emu
    from oracle.cases import case_text_layer
    case = CASES_BY_NAME[name]
    p = host_emu.oracle_to_product_params(case.params)
    frames = torch.from_numpy(np.ascontiguousarray(np.stack(case_frames(case)))).cuda()
    planes = harness.noise_planes(case)
    nz = None if planes is None else torch.from_numpy(np.stack(planes)).cuda()
    sc = [harness.frame_scalars(case, j) for j in range(case.frames)]
    kw = dict(phases=[s[0] for s in sc], times=[s[1] for s in sc], first_index=case.first_index, noise_planes=nz)
    outs = {}
    for order in ("rgb", "bgr"):
        eng = CrtEngine(case.w, case.h).configure(p, variant="export", text_rgba=case_text_layer(case), text_after=(case.text != "before"),
                                                  channel_order=order)
        src = frames if order == "rgb" else frames.flip(-1).contiguous()
        out, state = eng.process(src, **kw)
        outs[order] = (out if order == "rgb" else out.flip(-1), state if order == "rgb" else state.flip(-1))
        eng.close()
    assert torch.equal(outs["rgb"][0], outs["bgr"][0]) and torch.equal(outs["rgb"][1], outs["bgr"][1])


# ----------------------------------------------------------------------------------------- thread safety --
def test_two_threads_two_contexts_match_the_serial_run():
    """The reference calls the chain from two pool threads (crt_filter.py:1015-1017); one context per thread.  Two
    engines with DIFFERENT parameter sets (different kernels, different shared-memory opt-ins) process their clips
    concurrently, several times over; every output must equal the same engine's single-threaded result."""
    import torch
    from pythoncrt_b200.engine import CrtEngine
    import host_emu
    jobs = [("cfg2_gauss_grade", 1080, 1920), ("cfg3_warp", 720, 1280), ("cfg1_cli_default", 1080, 1920), ("cfg4_full_fastbloom", 480, 640)]
    clips, want, params = [], [], []
    for k, (name, h, w) in enumerate(jobs):
        p = host_emu.oracle_to_product_params(CASES_BY_NAME[name].params).but(noise_strength=0.0, glitch_amp_px=0)
        g = torch.Generator(device="cuda").manual_seed(50 + k)
        fr = torch.randint(0, 256, (6, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
        eng = CrtEngine(w, h).configure(p)
        out, _ = eng.process(fr, fps=30.0)
        torch.cuda.synchronize()
        eng.close()
        clips.append(fr); want.append(out.clone()); params.append(p)
    errors = []

    def worker(k):
        try:
            name, h, w = jobs[k]
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for rep in range(4):
                    eng = CrtEngine(w, h).configure(params[k])          # fresh context: first-use configuration races are exercised
                    out, _ = eng.process(clips[k], fps=30.0)
                    stream.synchronize()
                    if not torch.equal(out, want[k]):
                        errors.append((name, rep, int((out != want[k]).sum())))
                    eng.close()
        except Exception as e:  # noqa: BLE001
            errors.append((jobs[k][0], repr(e)))

    for pair in ((0, 1), (2, 3), (0, 2)):
        ts = [threading.Thread(target=worker, args=(k,)) for k in pair]
        [t.start() for t in ts]
        [t.join() for t in ts]
    assert not errors, errors


def test_drop_in_noise_differs_between_frames_of_two_workers():
    """after install(), frames handled by the two export workers must not share their grain."""
    import pythoncrt_b200 as crt
    case = CASES_BY_NAME["noise_grain1"]
    p = case.params
    frame = np.full((case.h, case.w, 3), 128, np.uint8)
    tri = crt.make_triad_mask(case.h, case.w, p.triad_strength, p.triad_softness)
    vig = crt.make_vignette(case.h, case.w, p.vignette_strength)
    imgs = {}

    def worker(start):
        for i in range(start, 6, 2):
            t = i / 30.0
            imgs[i] = crt.apply_static_effects(frame, p.scanline_strength, tri, float(p.triad_gamma), False, 1, p.bloom_sigma, p.bloom_strength,
                                               0.0, p.noise_strength, vig, p.scanline_period_px, t * 30.0, True, 2, 0, 0.0, time_sec=t)
    ts = [threading.Thread(target=worker, args=(s,)) for s in (0, 1)]
    [t.start() for t in ts]; [t.join() for t in ts]
    flat = [imgs[i] for i in range(6)]
    for a in range(6):
        for b in range(a + 1, 6):
            assert not np.array_equal(flat[a], flat[b]), (a, b)


# ----------------------------------------------------------------------------- generators vs the reference --
def _ks(a, b):
    from scipy import stats
    return stats.ks_2samp(np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel())


def test_generated_noise_has_the_distribution_of_cv2_randn():
    """noise_mode = generate (what bench.py times) against the reference's own generator (cv2.randn, :641): two-sample
    Kolmogorov-Smirnov on 2 x 300k draws, moments, tail mass, and independence across frames and cells."""
    import cv2
    import torch
    from pythoncrt_b200 import CrtEngine, CrtParams
    h, w = 480, 640
    eng = CrtEngine(w, h).configure(CrtParams(noise_strength=2.0, grain_size=1), noise_mode="generate", seed=77)
    a = eng.generate_noise(5).cpu().numpy().ravel()
    b = eng.generate_noise(6).cpu().numpy().ravel()
    ref = np.empty((h, w), np.float32)
    cv2.setRNGSeed(123)
    cv2.randn(ref, 0.0, 1.0)
    ks = _ks(a, ref)
    assert ks.pvalue > 1e-3, ks
    for x in (a, b):
        assert abs(x.mean()) < 6e-3 and abs(x.std() - 1.0) < 6e-3
        assert abs(float((x ** 3).mean())) < 0.03 and abs(float((x ** 4).mean()) - 3.0) < 0.08        # skewness 0, kurtosis 3
        assert abs(float((np.abs(x) > 3.0).mean()) - 0.0027) < 6e-4                                      # tail mass of N(0,1)
    assert abs(float(np.corrcoef(a, b)[0, 1])) < 6e-3                       # frames independent
    assert abs(float(np.corrcoef(a[:-1], a[1:])[0, 1])) < 6e-3              # neighbouring cells independent
    assert np.array_equal(a, eng.generate_noise(5).cpu().numpy().ravel())   # counter-based: same (seed, frame) -> same plane


@pytest.mark.parametrize("variant", ["gui", "export"])
def test_generated_glitch_has_the_distribution_of_the_reference_draws(variant):
    """glitch_mode = generate against numpy PCG64 draws made exactly as the reference makes them (:672-679, :845-853;
    tables.glitch_offsets is bit-identical to the reference, test_host_logic), 48 patterns from each generator.
    gui: one independent offset per row -> two-sample KS per row band (the amplitude decays down the band).
    export: offset = rint(row drift + per-segment noise); the drift is a random walk shared by the 80 segments of a row,
    so the raw offsets are not independent samples: KS on the per-segment part (offset minus the row mean), and the
    spread of the row means (drift + noise / 80) compared between the generators."""
    from pythoncrt_b200 import CrtEngine, CrtParams, tables
    h, w, amp, frac = 480, 640, 48, 0.5
    eng = CrtEngine(w, h).configure(CrtParams(glitch_amp_px=amp, glitch_height_frac=frac, noise_strength=0.0), variant=variant,
                                    glitch_mode="generate", seed=9)
    k = 0.05 if variant == "gui" else 2.0
    phases = [(j * 3 + 1) / k for j in range(48)]          # 48 distinct reference seeds
    gen = np.stack([eng.generate_glitch(phase_px=ph).cpu().numpy() for ph in phases]).astype(np.float64)
    ref = np.stack([tables.glitch_offsets(variant, h, w, amp, frac, ph) for ph in phases]).astype(np.float64)
    assert gen.shape == ref.shape
    rows = gen.shape[1]
    for lo, hi in ((0, rows // 4), (rows // 4, rows // 2), (rows // 2, rows)):
        a, b = gen[:, lo:hi], ref[:, lo:hi]
        if variant == "export":
            ra, rb = a.mean(axis=2), b.mean(axis=2)                      # per (pattern, row): drift + noise / segments
            sa, sb = ra.std(), rb.std()
            assert 0.7 * sb <= sa <= 1.4 * sb, (lo, hi, sa, sb)
            assert abs(ra.mean()) < 4 * sb / np.sqrt(48) + 0.05 and abs(ra.mean() - rb.mean()) < 6 * sb / np.sqrt(48) + 0.05
            a, b = a - ra[:, :, None], b - rb[:, :, None]
        ks = _ks(a, b)
        assert ks.pvalue > 1e-3, (variant, lo, hi, ks)
        ma, mb = np.abs(a).mean(), np.abs(b).mean()
        assert abs(ma - mb) <= 0.08 * max(mb, 0.5), (variant, lo, hi, ma, mb)
    assert np.abs(gen).max() <= amp * (1.0 if variant == "gui" else 4.0)


def test_generated_glitch_is_keyed_like_the_reference():
    """Temporal coherence: the reference reseeds on int(|phase| * k) (+ geometry) — the GUI holds a pattern for 20 phase-px
    (:670), the export for half a pixel (:841).  The device generator uses the same key."""
    import torch
    from pythoncrt_b200 import CrtEngine, CrtParams
    for variant, same, other in (("gui", (40.0, 59.9), 60.0), ("export", (3.0, 3.49), 3.5)):
        eng = CrtEngine(640, 480).configure(CrtParams(glitch_amp_px=32, glitch_height_frac=0.5), variant=variant, glitch_mode="generate", seed=1)
        a, b, c = (eng.generate_glitch(phase_px=ph) for ph in (same[0], same[1], other))
        assert torch.equal(a, b) and not torch.equal(a, c), variant
        assert torch.equal(a, eng.generate_glitch(phase_px=-same[0]))      # |phase|
        eng.close()


# ------------------------------------------------------------------------------------- long persistence halo --
def test_temporal_shards_with_persistence_095():
    """SURVEY.md §8e at the CLI's maximum persistence: 149 warm-up frames (0.95^149 <= 1/2040).  Two shards of a
    400-frame clip against the serial run."""
    import torch
    from pythoncrt_b200 import CrtEngine, CrtParams, clip
    h, w, n = 96, 160, 400
    p = CrtParams(noise_strength=0.0, persistence=0.95)
    assert clip.halo_frames(p.persistence) == 149
    eng = CrtEngine(w, h).configure(p)
    g = torch.Generator(device="cuda").manual_seed(3)
    frames = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
    # slowly varying content: persistence 0.95 of white noise is a grey field, a drifting ramp is the harder case
    ramp = (torch.arange(n, device="cuda").view(n, 1, 1, 1) * 2 + torch.arange(w, device="cuda").view(1, 1, w, 1)) % 256
    frames = ((frames.to(torch.int32) // 4 + ramp.to(torch.int32)) % 256).to(torch.uint8).contiguous()
    serial, _ = clip.process_clip(eng, frames, fps=30.0)
    for rank in range(2):
        def run_range(first, last, fresh):
            out, _ = clip.process_clip(eng, frames[first:last], fps=30.0, first_index=first)
            return out
        mine, (a, b) = clip.process_clip_sharded(run_range, n, p, rank, 2)
        warm, _, _ = clip.shard_plan(n, rank, 2, p.persistence)
        assert (rank == 0 and warm == 0) or a - warm == 149
        d = (mine.to(torch.int16) - serial[a:b].to(torch.int16)).abs()
        assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 2e-3, (rank, int(d.max()), float((d > 0).float().mean()))
    # a halo that is too short is visibly wrong: the test above is not vacuous
    short, _ = clip.process_clip(eng, frames[200 - 10:260], fps=30.0, first_index=190)
    assert int((short[10:].to(torch.int16) - serial[200:260].to(torch.int16)).abs().max()) > 1


# ------------------------------------------------------------------------ patched export pipeline end to end --
_DRAIN_SRC = '''
def export_like(frames, fps, speed, persistence, chain_args, chain_kw, workers=2):
    """The frame loop of process_video (crt_filter.py:1015-1131) without decode / encode: 2-worker pool running
    apply_static_effects, ordered drain with the persistence blend (:1092) and convertScaleAbs (:1098)."""
    from concurrent.futures import ThreadPoolExecutor
    outs, futures, next_write, prev_state = [], {}, 0, None
    def drain_one():
        nonlocal next_write, prev_state
        static_img = futures.pop(next_write).result()
        if prev_state is not None and persistence > 0.0:
            blended = np.clip(persistence * prev_state + (1.0 - persistence) * static_img, 0.0, 1.0)
        else:
            blended = static_img
        prev_state = blended
        outs.append(cv2.convertScaleAbs(blended, alpha=255.0, beta=0))
        next_write += 1
    with ThreadPoolExecutor(max_workers=workers) as executor:
        for i, frame in enumerate(frames):
            a = list(chain_args)
            a[11] = (i / float(fps)) * speed
            futures[i] = executor.submit(apply_static_effects, frame, *a, time_sec=(i / float(fps)), **chain_kw)
            while len(futures) >= workers * 4 or next_write in futures:
                if next_write in futures:
                    drain_one()
                else:
                    break
        while next_write in futures:
            drain_one()
    return outs
'''


@pytest.mark.parametrize("name,persistence", [("cfg1_cli_default", 0.2), ("cfg4_full", 0.35), ("cfg3_warp", 0.0)])
def test_install_runs_the_export_drain_on_the_device(name, persistence):
    """effects.install on a stand-in module whose frame loop restates process_video's: the patched pipeline must give the
    oracle's export frames (+-1 LSB), identical bytes to the device-resident engine path, and move exactly 3 bytes per
    pixel each way per frame (no float image crosses the host link)."""
    import cv2
    import pythoncrt_b200 as crt
    from pythoncrt_b200 import effects
    case = CASES_BY_NAME[name]
    p = case.params.but(persistence=persistence, noise_strength=0.0)
    case = Case(case.name, case.h, case.w, p, frames=7, fps=case.fps, source=case.source)
    mod = types.ModuleType("crt_filter_standin")
    mod.np, mod.cv2 = np, cv2
    mod.apply_static_effects = lambda *a, **k: (_ for _ in ()).throw(AssertionError("not patched"))
    exec(_DRAIN_SRC, mod.__dict__)
    effects.install(mod)
    tri = mod.make_triad_mask(case.h, case.w, p.triad_strength, p.triad_softness)
    vig = mod.make_vignette(case.h, case.w, p.vignette_strength)
    args = (p.scanline_strength, tri, float(p.triad_gamma), bool(p.triad_preserve_luma), int(p.aberration_px), p.bloom_sigma, p.bloom_strength,
            float(p.bloom_threshold), p.noise_strength, vig, p.scanline_period_px, 0.0, p.fast_bloom, int(p.pixel_size), int(p.glitch_amp_px),
            float(p.glitch_height_frac))
    kw = dict(brightness=p.brightness, contrast=p.contrast, gamma=p.gamma, saturation=p.saturation, temperature=p.temperature,
              flicker_strength=p.flicker_strength, flicker_hz=p.flicker_hz, grain_size=p.grain_size, scanline_angle=p.scanline_angle,
              scanline_thickness=p.scanline_thickness, warp_strength=p.warp_strength)
    frames = case_frames(case)
    before = dict(effects.TRANSFER_LOG)
    outs = mod.export_like(frames, case.fps, p.scanline_speed_px_s, p.persistence, args, kw)
    moved = {k: effects.TRANSFER_LOG[k] - before[k] for k in before}
    fbytes = case.h * case.w * 3
    assert moved == {"h2d_bytes": 7 * fbytes, "d2h_bytes": 7 * fbytes, "frames": 7}, moved
    want, _ = harness.run_oracle(case, "export")
    for a, b in zip(want, outs):
        st = harness.diff_stats(a, b)
        assert b.dtype == np.uint8 and st["max"] <= 1 and st["psnr"] >= 50, st
    got, _, _ = run_case_gpu(case, "export")
    assert all(np.array_equal(a, b) for a, b in zip(got, outs))


def test_export_frame_still_converts_to_the_float_image():
    """An ExportFrame handed to any other consumer behaves like apply_static_effects' return value (:861)."""
    import pythoncrt_b200 as crt
    from pythoncrt_b200 import effects
    case = CASES_BY_NAME["cfg3_warp"]
    p = case.params
    tri = crt.make_triad_mask(case.h, case.w, p.triad_strength, p.triad_softness)
    vig = crt.make_vignette(case.h, case.w, p.vignette_strength)
    frame = case_frames(case)[0]
    phase, tsec = harness.frame_scalars(case, 0)
    a = (frame, p.scanline_strength, tri, float(p.triad_gamma), False, int(p.aberration_px), p.bloom_sigma, p.bloom_strength, 0.0,
         p.noise_strength, vig, p.scanline_period_px, phase, p.fast_bloom, int(p.pixel_size), 0, 0.0)
    kw = dict(time_sec=tsec, scanline_angle=p.scanline_angle, scanline_thickness=p.scanline_thickness, warp_strength=p.warp_strength)
    lazy = effects.apply_static_effects_lazy(*a, **kw)
    img = np.asarray(lazy)
    ref = O.static_chain(frame, p, phase_px=phase, time_sec=tsec, variant="export")
    assert img.dtype == np.float32 and img.shape == ref.shape and np.max(np.abs(ref - img)) < 4e-6
    assert np.array_equal(img, crt.apply_static_effects(*a, **kw))


def test_stale_state_is_resized_like_the_reference():
    """State of another frame size: the GUI chain resizes it with cv2.resize INTER_LINEAR (:689-690), the export drain
    through uint8 and PIL bilinear (:1088-1091)."""
    import cv2
    from PIL import Image
    import pythoncrt_b200 as crt
    from pythoncrt_b200 import effects
    case = CASES_BY_NAME["cfg1_cli_default"]
    p = case.params
    h, w = case.h, case.w
    frame = case_frames(case)[0]
    rng = np.random.default_rng(4)
    small = rng.random((h // 2, w // 2 + 8, 3))          # float64, like the state the reference's GUI keeps once the vignette promoted it
    tri = crt.make_triad_mask(h, w, p.triad_strength, p.triad_softness)
    vig = crt.make_vignette(h, w, p.vignette_strength)
    args = (p.scanline_strength, tri, float(p.triad_gamma), False, 1, p.bloom_sigma, p.bloom_strength, 0.0, 0.0, vig)
    # GUI: oracle fed the state resized the reference's way
    out, state = crt.apply_crt_effect(frame, *args, 0.5, small, p.scanline_period_px, 7.0, True, 2)
    want, _ = O.frame_step(frame, p.but(persistence=0.5), cv2.resize(small, (w, h), interpolation=cv2.INTER_LINEAR), phase_px=7.0, time_sec=0.0,
                           variant="gui")
    assert harness.diff_stats(want, out)["max"] <= 1
    # export drain: blend against a foreign float state of another size
    cur = effects.apply_static_effects_lazy(frame, *args, p.scanline_period_px, 7.0, True, 2, 0, 0.0)
    small = small.astype(np.float32)
    blended = np.clip(0.5 * _Foreign(small) + (1.0 - 0.5) * cur, 0.0, 1.0)
    got = blended.quantised()
    prev = np.asarray(Image.fromarray(np.clip(small * 255.0, 0, 255).astype(np.uint8)).resize((w, h), Image.BILINEAR)).astype(np.float32) / 255.0
    img = O.static_chain(frame, p, phase_px=7.0, time_sec=0.0, variant="export")
    want = cv2.convertScaleAbs(np.clip(0.5 * prev + 0.5 * img, 0.0, 1.0), alpha=255.0, beta=0)
    assert harness.diff_stats(want, got)["max"] <= 1


class _Foreign:
    """A float state that did not come from this library (e.g. kept by the caller across an install())."""
    def __init__(self, arr):
        self.arr, self.shape, self.state = arr, arr.shape, None

    def __rmul__(self, k):
        from pythoncrt_b200 import effects
        return effects._Scaled(self.arr, k)


# ------------------------------------------------------------------------------- intra-GPU temporal shards --
@pytest.mark.parametrize("name,hw,n", [("cfg1_cli_default", (480, 640), 260), ("cfg2_gauss_grade", (270, 480), 230), ("cfg3_warp", (360, 640), 200),
                                       ("cfg4_full", (240, 320), 210), ("no_persistence", (480, 640), 200)])
def test_concurrent_temporal_shards_match_the_serial_run(name, hw, n):
    """crt_set_shards: a long clip processed as concurrent temporal shards on ONE GPU (own stream + context each, persistence
    warm-up per shard) against the strictly serial call: at most 1 LSB apart, in a tiny fraction of samples (identical with
    persistence off), same final state; the generated noise / glitch draws are keyed by global frame index / phase."""
    import torch
    from pythoncrt_b200 import CrtEngine
    import host_emu
    h, w = hw
    p = host_emu.oracle_to_product_params(CASES_BY_NAME[name].params)
    g = torch.Generator(device="cuda").manual_seed(21)
    frames = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
    kw = dict(variant="export", noise_mode="generate", glitch_mode="generate", seed=5)
    serial, s_state = CrtEngine(w, h).configure(p, shards=1, **kw).process(frames, fps=30.0, first_index=7)
    for shards in ("auto", 3):
        eng = CrtEngine(w, h).configure(p, shards=shards, **kw)
        out, state = eng.process(frames, fps=30.0, first_index=7)
        torch.cuda.synchronize()
        k = int(eng.last_info.reserved[0])
        assert k == (4 if shards == "auto" else 3), k                     # the clips are long enough to be cut
        d = (out.to(torch.int16) - serial.to(torch.int16)).abs()
        if p.persistence == 0.0:
            assert int(d.max()) == 0
        else:
            assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 1e-3, (int(d.max()), float((d > 0).float().mean()))
            assert int(eng.last_info.reserved[1]) == 5                    # 0.2^5 <= 1/2040
        assert float((state - s_state).abs().max()) < 1e-3
        # a second call continues from the returned state like the serial engine does
        out2, _ = eng.process(frames[:8], state=state, state_valid=True, fps=30.0, first_index=7 + n)
        ref2, _ = CrtEngine(w, h).configure(p, shards=1, **kw).process(frames[:8], state=s_state.clone(), state_valid=True, fps=30.0, first_index=7 + n)
        assert int((out2.to(torch.int16) - ref2.to(torch.int16)).abs().max()) <= 1
        eng.close()


def test_short_clips_are_never_sharded():
    import torch
    from pythoncrt_b200 import CrtEngine, CrtParams
    eng = CrtEngine(128, 96).configure(CrtParams(noise_strength=0.0), shards="auto")
    frames = torch.zeros((40, 96, 128, 3), dtype=torch.uint8, device="cuda")
    eng.process(frames)
    assert int(eng.last_info.reserved[0]) == 1
    eng2 = CrtEngine(128, 96).configure(CrtParams(noise_strength=0.0, persistence=0.95), shards="auto")
    eng2.process(torch.zeros((600, 96, 128, 3), dtype=torch.uint8, device="cuda"))
    assert int(eng2.last_info.reserved[0]) == 1          # 149 warm-up frames per shard: not worth it at 600 frames


def test_state_resize_is_cv2_resize():
    """crt_resize_state (GUI chain, crt_filter.py:689-690) against cv2.resize(INTER_LINEAR) itself: bit for bit on float32."""
    import cv2
    import torch
    from pythoncrt_b200 import CrtEngine
    eng = CrtEngine(128, 96)
    rng = np.random.default_rng(8)
    for h0, w0 in ((48, 72), (200, 300), (97, 131), (96, 128), (30, 500)):
        src = rng.random((h0, w0, 3)).astype(np.float32)
        got = eng.resize_state(torch.from_numpy(src).cuda()).cpu().numpy()
        want = cv2.resize(src, (128, 96), interpolation=cv2.INTER_LINEAR)
        assert got.shape == (96, 128, 3) and np.array_equal(got, want), (h0, w0, float(np.abs(got - want).max()))


# ------------------------------------------------------------------------- opt-in single-pass warp kernels --
@pytest.mark.parametrize("env", ["CRT_WARP_SRC", "CRT_WARP_PS2"])
@pytest.mark.parametrize("name", ["cfg3_warp", "cfg3_warp_structured", "warp_negative", "text_after_warp"])
def test_opt_in_single_pass_warp_kernels_match_the_oracle(env, name, monkeypatch):
    """csrc/crt_fused_warp_src.cuh (source-driven) and csrc/crt_fused_warp_ps2.cuh (output-driven) are measured slower than
    the two-pass path and therefore opt-in (the variable is read when a context is configured); they stay under the same
    parity bar as the default path: +-1 LSB against the oracle of crt_filter.py:678-690, and against the default path."""
    case = CASES_BY_NAME[name]
    base_out, base_state, _ = run_case_gpu(case, "export")
    monkeypatch.setenv(env, "1")
    outs, state, fused = run_case_gpu(case, "export")
    want, _ = harness.run_oracle(case, "export", backend="cv2")
    worst = max(int(np.abs(o.astype(np.int16) - w.astype(np.int16)).max()) for o, w in zip(outs, want))
    assert worst <= 1, (env, name, worst)
    drift = max(int(np.abs(o.astype(np.int16) - b.astype(np.int16)).max()) for o, b in zip(outs, base_out))
    assert drift <= 1, (env, name, drift)
    assert np.abs(state - base_state).max() <= 1.0 / 255 + 1e-6


@pytest.mark.parametrize("env", ["CRT_WARP_SRC", "CRT_WARP_PS2"])
def test_opt_in_single_pass_warp_kernels_at_4k(env, monkeypatch):
    """At BASELINE configs[2]'s size the opt-in kernels take their tiled fast paths (interior tiles, border items): same
    bytes as the default two-pass path within 1 LSB over a persistence chain."""
    import torch
    from pythoncrt_b200.engine import CrtEngine
    import host_emu
    p = host_emu.oracle_to_product_params(CASES_BY_NAME["cfg3_warp"].params).but(noise_strength=0.0, glitch_amp_px=0)
    g = torch.Generator(device="cuda").manual_seed(77)
    fr = torch.randint(0, 256, (4, 2160, 3840, 3), dtype=torch.uint8, device="cuda", generator=g)
    res = []
    for on in (False, True):
        if on:
            monkeypatch.setenv(env, "1")
        eng = CrtEngine(3840, 2160).configure(p)
        out, state = eng.process(fr, fps=30.0)
        res.append((out.clone(), state.clone(), int(eng.last_info.kernels_launched)))
        eng.close()
    assert res[1][2] < res[0][2]                                     # one kernel per frame instead of two
    assert int((res[0][0].to(torch.int16) - res[1][0].to(torch.int16)).abs().max()) <= 1
    assert float((res[0][1] - res[1][1]).abs().max()) <= 1.0 / 255 + 1e-6


# ------------------------------------------------------------------------------------------- clip mode --
def _clip_params(kind):
    import host_emu
    base = host_emu.oracle_to_product_params(CASES_BY_NAME["cfg1_cli_default"].params).but(noise_strength=0.0, glitch_amp_px=0)
    if kind == "default":
        return base
    if kind == "slanted_flicker":      # per-frame scalars: scanline phase and flicker gain differ from frame to frame
        return base.but(scanline_angle=3.0, scanline_thickness=1.2, flicker_strength=0.08, flicker_hz=50.0)
    if kind == "gauss_grade":          # BASELINE configs[1]: gaussian bloom (K = 9) + colour grade
        return host_emu.oracle_to_product_params(CASES_BY_NAME["cfg2_gauss_grade"].params).but(noise_strength=0.0, glitch_amp_px=0)
    if kind == "gauss_wide":           # K = 7, per-frame scalars
        return base.but(fast_bloom=False, bloom_sigma=1.0, scanline_angle=2.0, flicker_strength=0.05, flicker_hz=50.0)
    if kind == "threshold":
        return base.but(bloom_threshold=0.6)
    if kind == "no_bloom":
        return base.but(bloom_strength=0.0)
    if kind == "no_triad":             # the general per-pixel tail (FAST = false instantiations)
        return base.but(triad_strength=0.0, flicker_strength=0.05, flicker_hz=50.0)
    if kind == "no_triad_no_bloom":
        return base.but(triad_strength=0.0, bloom_strength=0.0)
    if kind == "preserve_luma":        # triad with luma preservation: not the composite-LUT tail
        return base.but(triad_preserve_luma=True)
    raise KeyError(kind)


@pytest.mark.parametrize("kind,hw,n", [("default", (1080, 1920), 40), ("default", (720, 1280), 70), ("default", (2160, 3840), 6),
                                       ("slanted_flicker", (1080, 1920), 24), ("threshold", (1080, 1920), 12), ("no_bloom", (720, 1280), 12),
                                       ("default", (480, 640), 130), ("gauss_grade", (1080, 1920), 40), ("gauss_grade", (2160, 3840), 5),
                                       ("gauss_wide", (1080, 1920), 10), ("no_triad", (1080, 1920), 9), ("no_triad_no_bloom", (1080, 1920), 9),
                                       ("preserve_luma", (1080, 1920), 9), ("default", (4320, 7680), 4), ("gauss_grade", (4320, 7680), 4)])
def test_clip_mode_is_the_serial_run_bit_for_bit(kind, hw, n, monkeypatch):
    """Clip mode (csrc/crt_fused_ps2.cuh ClipArgs): a run of frames in ONE launch, chained tile by tile through the persistence
    state (crt_filter.py:1092).  Same kernel arithmetic in the same order per pixel -> identical bytes and identical state to
    one launch per frame, for runs longer than a launch holds (64), frames smaller than the resident grid, per-frame scalars,
    and a second call continuing from the first call's state."""
    import torch
    from pythoncrt_b200.engine import CrtEngine
    h, w = hw
    p = _clip_params(kind)
    g = torch.Generator(device="cuda").manual_seed(1234 + n)
    fr = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
    phases = [0.37 * j for j in range(n)]
    times = [j / 30.0 for j in range(n)]
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("CRT_CLIP", mode)
        eng = CrtEngine(w, h).configure(p)
        eng.set_shards(1)
        cut = n // 3
        out = torch.empty_like(fr)
        state = eng.new_state()
        eng.process(fr[:cut], out[:cut], state=state, state_valid=False, phases=phases[:cut], times=times[:cut])
        first = int(eng.last_info.reserved[2])
        eng.process(fr[cut:], out[cut:], state=state, state_valid=True, phases=phases[cut:], times=times[cut:], first_index=cut)
        res[mode] = (out, state.clone(), first, int(eng.last_info.reserved[2]), int(eng.last_info.kernels_launched))
        eng.close()
    assert res["0"][2] == 0 and res["0"][3] == 0
    if h * w >= 1920 * 1080:            # (smaller frames: fewer tiles than resident CTAs -> one launch per frame, see clip_wanted)
        assert res["1"][3] == n - n // 3 - ((n - n // 3) % 64 == 1), res["1"][2:]          # the second call went through clip-mode launches
        assert res["1"][4] <= (n - n // 3 + 63) // 64 + 1
    else:
        monkeypatch.setenv("CRT_CLIP_MIN_TILES", "1")         # force clip mode on the small frame too
        monkeypatch.setenv("CRT_CLIP", "1")
        eng = CrtEngine(w, h).configure(p)
        eng.set_shards(1)
        out2, state2 = eng.process(fr, phases=phases, times=times)
        assert int(eng.last_info.reserved[2]) == n - 1 - ((n - 1) % 64 == 1)      # runs of <= 64 frames; a single left-over frame goes alone
        eng.close()
        assert torch.equal(out2, res["0"][0]) and torch.equal(state2, res["0"][1])
    assert torch.equal(res["0"][0], res["1"][0])
    assert torch.equal(res["0"][1], res["1"][1])


@pytest.mark.parametrize("env", [{}, {"CRT_CLIP_ITEMS": "0"}, {"CRT_CLIP_ITEMS": "1", "CRT_CLIP_COOP_GAUSS": "1"}, {"CRT_CLIP_ITEMS": "2"},
                                 {"CRT_CLIP_ITEMS": "3"}, {"CRT_CLIP_ITEMS": "0", "CRT_CLIP_RELEASE": "2"}, {"CRT_CLIP_ITEMS": "0", "CRT_CLIP_RELEASE": "23"}])
@pytest.mark.parametrize("kind", ["default", "gauss_grade"])
def test_clip_mode_variants_are_bit_exact_too(kind, env, monkeypatch):
    """The two flag protocols — fixed stride under a cooperative launch (default of the fast-bloom kernel), atomic item counter
    (default of the gaussian kernel) — with every publication mode that carries a gpu-scope membar, and owned tiles (CTA i keeps
    tiles i, i + G, ... through the run, no flags; measured slower, opt-in) (csrc/crt_fused_ps2.cuh clip_publish; the mode without one is measurably racy and not offered): same bytes as one
    launch per frame, twice in a row (a race would show as a difference between runs)."""
    import torch
    from pythoncrt_b200.engine import CrtEngine
    p = _clip_params(kind)
    g = torch.Generator(device="cuda").manual_seed(99)
    fr = torch.randint(0, 256, (30, 1080, 1920, 3), dtype=torch.uint8, device="cuda", generator=g)
    res = []
    for clip in ("0", "1", "1"):
        monkeypatch.setenv("CRT_CLIP", clip)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng = CrtEngine(1920, 1080).configure(p)
        eng.set_shards(1)
        out, state = eng.process(fr, fps=30.0)
        res.append((out, state, int(eng.last_info.reserved[2])))
        eng.close()
    assert res[0][2] == 0 and res[1][2] == 29 and res[2][2] == 29
    for k in (1, 2):
        assert torch.equal(res[0][0], res[k][0]) and torch.equal(res[0][1], res[k][1])


def test_clip_mode_matches_the_oracle():
    """... and against the oracle of the chain, with the reference's own state hand-off."""
    case = CASES_BY_NAME["cfg1_cli_default"]
    outs, state, fused = run_case_gpu(case, "export")
    want, want_state = harness.run_oracle(case, "export", backend="cv2")
    assert max(int(np.abs(o.astype(np.int16) - w.astype(np.int16)).max()) for o, w in zip(outs, want)) <= 1


@pytest.mark.parametrize("kind", ["default", "gauss_grade", "threshold"])
@pytest.mark.parametrize("hw", [(1080, 1920), (1440, 2560), (900, 1600), (1088, 2048)])
def test_clip_mode_is_deterministic_across_sizes(kind, hw, monkeypatch):
    """Race hunt in the suite: the flag protocol's margins depend on tiles per resident CTA (1080p: 1.7 rounds for the fast-bloom
    kernel, 2.3 for the gaussian one).  Three clip-mode runs of 48 frames against one launch per frame, with host <-> device
    copies running on another stream as in crt_process_host: every byte equal every time."""
    import torch
    from pythoncrt_b200.engine import CrtEngine
    h, w = hw
    p = _clip_params(kind)
    g = torch.Generator(device="cuda").manual_seed(h + w)
    fr = torch.randint(0, 256, (48, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
    monkeypatch.setenv("CRT_CLIP", "0")
    eng = CrtEngine(w, h).configure(p)
    eng.set_shards(1)
    ref, ref_state = eng.process(fr, fps=30.0)
    eng.close()
    monkeypatch.setenv("CRT_CLIP", "1")
    side = torch.cuda.Stream()
    hbuf = torch.empty((64 << 20,), dtype=torch.uint8).pin_memory()
    dbuf = torch.empty((64 << 20,), dtype=torch.uint8, device="cuda")
    for run in range(3):
        eng = CrtEngine(w, h).configure(p)
        eng.set_shards(1)
        with torch.cuda.stream(side):
            for _ in range(4):
                dbuf.copy_(hbuf, non_blocking=True)
                hbuf.copy_(dbuf, non_blocking=True)
        out, state = eng.process(fr, fps=30.0)
        torch.cuda.synchronize()
        assert int(eng.last_info.reserved[2]) == 47, (kind, hw)
        eng.close()
        assert torch.equal(out, ref) and torch.equal(state, ref_state), (kind, hw, run, int((out != ref).sum()))
