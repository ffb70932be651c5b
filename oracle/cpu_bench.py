"""Timing harness for the reference's CPU path (TEST/BENCH INFRASTRUCTURE).

Drives the oracle restatement (same numpy passes and cv2 calls as the reference)
the way process_video drives the reference chain (crt_filter.py:1015-1131):
a ThreadPoolExecutor with max(1, min(2, cpu // 2)) workers runs the stateless
chain, the main thread drains futures in frame order, blends the persistence
state and quantises.  Decode/encode are left out (out of scope by north_star).
"""
from __future__ import annotations

import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import crt_oracle as O


def export_like_loop(frames, p: O.ChainParams, fps: float, first_index: int = 0):
    """Returns (list of uint8 frames, seconds, worker threads)."""
    h, w = frames[0].shape[:2]
    tri = O.triad_mask(h, w, p.triad_strength, p.triad_softness) if p.triad_strength > 0.0 else None   # once per clip, :919
    vig = O.vignette_mask(h, w, p.vignette_strength) if p.vignette_strength > 0.0 else None            # :920
    workers = max(1, min(2, (os.cpu_count() or 4) // 2))                                               # :1015
    cap = workers * 4                                                                                  # :1016
    outs, futures, nxt, state = [], {}, 0, None
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=workers) as pool:
        for j, frame in enumerate(frames):
            i = first_index + j
            futures[j] = pool.submit(O.static_chain, frame, p, phase_px=(i / float(fps)) * p.scanline_speed_px_s,
                                     time_sec=i / float(fps), variant="export", triad=tri, vignette=vig, build_masks=False)
            while len(futures) >= cap or nxt in futures and futures[nxt].done():
                if nxt in futures:
                    img = futures.pop(nxt).result()
                    state = O.blend_export(img, state, p.persistence)
                    outs.append(O.quantise(state))
                    nxt += 1
                else:
                    break
        while nxt in futures:
            img = futures.pop(nxt).result()
            state = O.blend_export(img, state, p.persistence)
            outs.append(O.quantise(state))
            nxt += 1
    return outs, time.perf_counter() - t0, workers


def reference_export_loop(ref, frames, p: O.ChainParams, fps: float, first_index: int = 0):
    """Same loop driving the UNMODIFIED reference module `ref` (oracle/ref_loader.py): its own make_triad_mask /
    make_vignette once per clip (:919-920), its apply_static_effects on the pool (:1045-1078), and the drain's two
    statements verbatim in meaning (:1092, :1098).  Returns (list of uint8 frames, seconds, worker threads)."""
    import cv2
    h, w = frames[0].shape[:2]
    tri = ref.make_triad_mask(h, w, p.triad_strength, p.triad_softness) if p.triad_strength > 0.0 else None
    vig = ref.make_vignette(h, w, p.vignette_strength) if p.vignette_strength > 0.0 else None
    workers = max(1, min(2, (os.cpu_count() or 4) // 2))
    cap = workers * 4
    persistence = float(p.persistence)
    outs, futures, nxt, prev_state = [], {}, 0, None

    def drain_one():
        nonlocal nxt, prev_state
        static_img = futures.pop(nxt).result()
        if prev_state is not None and persistence > 0.0:
            blended = np.clip(persistence * prev_state + (1.0 - persistence) * static_img, 0.0, 1.0)
        else:
            blended = static_img
        prev_state = blended
        outs.append(cv2.convertScaleAbs(blended, alpha=255.0, beta=0))
        nxt += 1

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=workers) as pool:
        for j, frame in enumerate(frames):
            i = first_index + j
            futures[j] = pool.submit(
                ref.apply_static_effects, frame, p.scanline_strength, tri, float(p.triad_gamma), bool(p.triad_preserve_luma),
                int(p.aberration_px), p.bloom_sigma, p.bloom_strength, float(p.bloom_threshold), p.noise_strength, vig,
                p.scanline_period_px, (i / float(fps)) * p.scanline_speed_px_s, bool(p.fast_bloom), int(p.pixel_size),
                int(p.glitch_amp_px), float(p.glitch_height_frac), time_sec=(i / float(fps)), brightness=float(p.brightness),
                contrast=float(p.contrast), gamma=float(p.gamma), saturation=float(p.saturation), temperature=float(p.temperature),
                flicker_strength=float(p.flicker_strength), flicker_hz=float(p.flicker_hz), grain_size=int(p.grain_size),
                scanline_angle=float(p.scanline_angle), scanline_thickness=float(p.scanline_thickness),
                warp_strength=float(p.warp_strength), text_overlay_rgba=None, text_overlay_after=True)
            while len(futures) >= cap or nxt in futures:
                if nxt in futures:
                    drain_one()
                else:
                    break
        while nxt in futures:
            drain_one()
    return outs, time.perf_counter() - t0, workers


def reference_module():
    """The unmodified reference if it can be imported here (/root/reference or the staged baseline/_ref copy), else None."""
    try:
        from . import ref_loader
        return ref_loader.load() if ref_loader.available() else None
    except Exception:  # noqa: BLE001
        return None


def time_cpu_path(h: int, w: int, p: O.ChainParams, fps: float, n_frames: int, repeats: int = 1, prefer_reference: bool = True):
    """Frames/s of the export-like CPU loop on `n_frames` synthetic frames (best of `repeats`): through the unmodified
    reference when it is importable (kind "reference"), else through the oracle port (kind "port")."""
    frames = [O.synthetic_frame(i, h, w) for i in range(n_frames)]
    ref = reference_module() if prefer_reference else None
    best = None
    for _ in range(repeats):
        if ref is not None:
            _, sec, workers = reference_export_loop(ref, frames, p, fps)
        else:
            _, sec, workers = export_like_loop(frames, p, fps)
        best = sec if best is None else min(best, sec)
    import cv2
    return {"fps": n_frames / best, "seconds": best, "frames": n_frames, "workers": workers, "kind": "reference" if ref is not None else "port",
            "cv2_threads": cv2.getNumThreads(), "cpu_count": os.cpu_count(), "numpy": np.__version__, "cv2": cv2.__version__}
