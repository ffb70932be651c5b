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


def time_cpu_path(h: int, w: int, p: O.ChainParams, fps: float, n_frames: int, repeats: int = 1):
    """Frames/s of the export-like CPU loop on `n_frames` synthetic frames (best of `repeats`)."""
    frames = [O.synthetic_frame(i, h, w) for i in range(n_frames)]
    best = None
    for _ in range(repeats):
        _, sec, workers = export_like_loop(frames, p, fps)
        best = sec if best is None else min(best, sec)
    import cv2
    return {"fps": n_frames / best, "seconds": best, "frames": n_frames, "workers": workers,
            "cv2_threads": cv2.getNumThreads(), "cpu_count": os.cpu_count(), "numpy": np.__version__, "cv2": cv2.__version__}
