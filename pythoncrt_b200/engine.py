"""CrtEngine — host-side owner of one C-ABI context (one device, one frame size).

PyTorch is used for what it is good at here: device memory, streams, pinned
host memory.  All compute goes through `libcrt_b200.so` (include/crt_b200.h).
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Optional, Sequence

import numpy as np

from . import cabi, tables
from .config import build_config
from .params import CrtParams


def _torch():
    import torch
    return torch


class CrtEngine:
    """One CRT chain context for frames of `height` x `width` on CUDA device `device`.

    configure() takes the reference's parameter surface (a `CrtParams`) plus the
    two mask inputs the reference passes as arrays (crt_filter.py:531-565):
    `triad_cols` ([W][3] table = any row of the triad mask, or None for "mask is
    None") and `vignette` (strength, an H x W array, or None).
    """

    def __init__(self, width: int, height: int, device: int = 0):
        torch = _torch()
        if not torch.cuda.is_available():
            raise cabi.CrtError("no CUDA device: pythoncrt_b200 has no CPU path")
        self.lib = cabi.load_library()
        self.width, self.height, self.device = int(width), int(height), int(device)
        handle = C.c_void_p()
        rc = self.lib.crt_create(self.device, self.width, self.height, C.byref(handle))
        if rc != 0:
            raise cabi.CrtError(f"crt_create failed (status {rc}): {self.lib.crt_last_error(None).decode()}")
        self.ctx = handle
        self.params: Optional[CrtParams] = None
        self.variant = "export"
        self.noise_mode = "inject"
        self.glitch_mode = "inject"
        self._lock = threading.Lock()
        self.last_info = cabi.CrtLaunchInfoC()
        self.kernels_launched = 0

    # ------------------------------------------------------------ lifetime --
    def close(self) -> None:
        if getattr(self, "ctx", None):
            self.lib.crt_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str) -> None:
        cabi.check(self.lib, self.ctx, rc, what)

    def _table(self, which: int, arr: Optional[np.ndarray]) -> None:
        if arr is None:
            self._check(self.lib.crt_set_table(self.ctx, which, None, 0), "crt_set_table")
            return
        a = np.ascontiguousarray(arr)
        self._check(self.lib.crt_set_table(self.ctx, which, a.ctypes.data_as(C.c_void_p), a.nbytes), "crt_set_table")

    # ----------------------------------------------------------- configure --
    def configure(self, params: CrtParams, *, variant: str = "export", triad_cols="auto", vignette="auto",
                  text_rgba: Optional[np.ndarray] = None, text_after: bool = True, noise_mode: str = "inject",
                  glitch_mode: str = "inject", seed: int = 0, policy: str = "auto", channel_order: str = "rgb", shards=1) -> "CrtEngine":
        """Upload parameters and host-built tables.

        triad_cols: "auto" builds the table from params.triad_strength/softness the
            way process_video does (crt_filter.py:919); an array [W][3] (or a full
            H x W x 3 mask, whose row 0 is used) is taken as is; None = no triad.
        vignette: "auto" uses params.vignette_strength (:920); a float is a strength;
            an H x W array is inspected — a make_vignette() mask is recognised and
            evaluated analytically on the device, anything else is uploaded; None = off.
        noise_mode / glitch_mode: "inject" = draws supplied per frame (the reference's
            own draws, for verification), "generate" = counter-based RNG on the device.
        shards: intra-GPU temporal shards for long clips given to process() (crt_set_shards): 1 = strictly serial
            (default), "auto" = up to 4 concurrent shards with a persistence warm-up each, k = at most k.
        channel_order: "rgb" = channel index 0 is R, as the reference assumes (crt_filter.py:289,
            :296-297, :573-575); "bgr" = frames (and state) are B,G,R: the R rules go to index 2.
        """
        c, tabs = build_config(params, self.width, self.height, variant=variant, triad_cols=triad_cols, vignette=vignette,
                               text_rgba=text_rgba, text_after=text_after, noise_mode=noise_mode, glitch_mode=glitch_mode, seed=seed,
                               channel_order=channel_order)
        for which, arr in tabs.items():
            self._table(which, arr)
        p = params
        self._check(self.lib.crt_set_params(self.ctx, C.byref(c)), "crt_set_params")
        self._check(self.lib.crt_set_policy(self.ctx, {"auto": 0, "staged": 1, "fused": 2}[policy]), "crt_set_policy")
        self._check(self.lib.crt_set_shards(self.ctx, 0 if shards == "auto" else int(shards)), "crt_set_shards")
        self.params, self.variant, self.noise_mode, self.glitch_mode = p, variant, noise_mode, glitch_mode
        self._c_params = c
        return self

    def set_shards(self, shards) -> None:
        """Intra-GPU temporal shards for process() on long clips (crt_set_shards): 1 = serial, "auto", or at most k."""
        self._check(self.lib.crt_set_shards(self.ctx, 0 if shards == "auto" else int(shards)), "crt_set_shards")

    # -------------------------------------------------------- frame records --
    def frame_records(self, n: int, *, first_index: int = 0, fps: float = 30.0, phases: Optional[Sequence[float]] = None,
                      times: Optional[Sequence[float]] = None, noise_planes=None, glitch_tables=None):
        """Build the crt_frame array.  By default phase/time follow process_video
        (crt_filter.py:1043, :1064).  `noise_planes`: CUDA float32 tensor [n][gh][gw];
        `glitch_tables`: list of CUDA int32 tensors [rows][segments] (or None entries)."""
        p = self.params
        recs = (cabi.CrtFrameC * n)()
        keep = []
        geom = tables.glitch_geometry(self.variant, self.height, self.width, p.glitch_height_frac) \
            if (p.glitch_amp_px > 0 and p.glitch_height_frac > 0.0) else None
        for j in range(n):
            i = first_index + j
            r = recs[j]
            r.phase_px = float(phases[j]) if phases is not None else p.phase_px(i, fps)
            r.time_sec = float(times[j]) if times is not None else p.time_sec(i, fps)
            r.frame_index = i
            if noise_planes is not None:
                r.d_noise = noise_planes[j].data_ptr()
            if glitch_tables is not None and glitch_tables[j] is not None and geom is not None:
                t = glitch_tables[j]
                keep.append(t)
                r.d_glitch_offs = t.data_ptr()
                r.glitch_y0, r.glitch_rows, r.glitch_seg_len, r.glitch_segments = geom
        return recs, keep

    def host_glitch_tables(self, n: int, *, first_index: int = 0, fps: float = 30.0, phases=None):
        """The reference's own glitch draws (numpy PCG64) for frames first_index.., as CUDA tensors."""
        torch = _torch()
        p = self.params
        out = []
        for j in range(n):
            ph = float(phases[j]) if phases is not None else p.phase_px(first_index + j, fps)
            t = tables.glitch_offsets(self.variant, self.height, self.width, int(p.glitch_amp_px), float(p.glitch_height_frac), ph)
            out.append(None if t is None else torch.from_numpy(t).to(f"cuda:{self.device}"))
        return out

    # -------------------------------------------------------------- process --
    def new_state(self):
        torch = _torch()
        return torch.empty((self.height, self.width, 3), dtype=torch.float32, device=f"cuda:{self.device}")

    def process(self, frames, out=None, *, state=None, state_valid: bool = False, **frame_kw):
        """Run [N][H][W][3] uint8 CUDA frames through the chain (crt_process).
        Returns (out, state).  `state` is updated in place."""
        torch = _torch()
        if frames.dim() == 3:
            frames = frames[None]
        n = frames.shape[0]
        assert frames.is_cuda and frames.dtype == torch.uint8 and frames.is_contiguous()
        assert tuple(frames.shape[1:]) == (self.height, self.width, 3), frames.shape
        if out is None:
            out = torch.empty_like(frames)
        if state is None:
            state = self.new_state()
            state_valid = False
        if self.params.glitch_amp_px > 0 and self.params.glitch_height_frac > 0 and self.glitch_mode == "inject" \
                and frame_kw.get("glitch_tables") is None:
            frame_kw["glitch_tables"] = self.host_glitch_tables(n, first_index=frame_kw.get("first_index", 0),
                                                                fps=frame_kw.get("fps", 30.0), phases=frame_kw.get("phases"))
        recs, keep = self.frame_records(n, **frame_kw)
        stream = torch.cuda.current_stream(frames.device).cuda_stream
        with self._lock:
            rc = self.lib.crt_process(self.ctx, frames.data_ptr(), out.data_ptr(), state.data_ptr(), int(bool(state_valid)),
                                      recs, n, C.c_void_p(stream), C.byref(self.last_info))
        self._check(rc, "crt_process")
        self.kernels_launched += self.last_info.kernels_launched
        return out, state

    def process_static(self, frames, **frame_kw):
        """Float image before persistence/quantise (crt_process_static = apply_static_effects)."""
        torch = _torch()
        if frames.dim() == 3:
            frames = frames[None]
        n = frames.shape[0]
        img = torch.empty(frames.shape, dtype=torch.float32, device=frames.device)
        if self.params.glitch_amp_px > 0 and self.params.glitch_height_frac > 0 and self.glitch_mode == "inject" \
                and frame_kw.get("glitch_tables") is None:
            frame_kw["glitch_tables"] = self.host_glitch_tables(n, first_index=frame_kw.get("first_index", 0),
                                                                fps=frame_kw.get("fps", 30.0), phases=frame_kw.get("phases"))
        recs, keep = self.frame_records(n, **frame_kw)
        stream = torch.cuda.current_stream(frames.device).cuda_stream
        with self._lock:
            rc = self.lib.crt_process_static(self.ctx, frames.data_ptr(), img.data_ptr(), recs, n, C.c_void_p(stream),
                                             C.byref(self.last_info))
        self._check(rc, "crt_process_static")
        self.kernels_launched += self.last_info.kernels_launched
        return img

    def process_host(self, frames: np.ndarray, out: Optional[np.ndarray] = None, **frame_kw) -> np.ndarray:
        """Host-buffer path (crt_process_host): numpy (ideally pinned) in, numpy out;
        the persistence state stays on the device inside the context."""
        a = np.ascontiguousarray(frames)
        if a.ndim == 3:
            a = a[None]
        n = a.shape[0]
        assert a.dtype == np.uint8 and a.shape[1:] == (self.height, self.width, 3), a.shape
        if out is None:
            out = np.empty_like(a)
        if self.params.glitch_amp_px > 0 and self.params.glitch_height_frac > 0 and self.glitch_mode == "inject" \
                and frame_kw.get("glitch_tables") is None:
            frame_kw["glitch_tables"] = self.host_glitch_tables(n, first_index=frame_kw.get("first_index", 0),
                                                                fps=frame_kw.get("fps", 30.0), phases=frame_kw.get("phases"))
        recs, keep = self.frame_records(n, **frame_kw)
        with self._lock:
            rc = self.lib.crt_process_host(self.ctx, a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), recs, n,
                                           C.byref(self.last_info))
        self._check(rc, "crt_process_host")
        self.kernels_launched += self.last_info.kernels_launched
        return out

    def resize_state(self, state):
        """A float32 CUDA state [h0][w0][3] of another frame size -> this engine's size, as the reference's GUI chain does
        with cv2.resize(INTER_LINEAR) (crt_filter.py:689-690); on the device (crt_resize_state)."""
        torch = _torch()
        src = state.contiguous().to(torch.float32)
        dst = self.new_state()
        stream = torch.cuda.current_stream(dst.device).cuda_stream
        self._check(self.lib.crt_resize_state(self.ctx, src.data_ptr(), int(src.shape[1]), int(src.shape[0]), dst.data_ptr(),
                                              C.c_void_p(stream)), "crt_resize_state")
        return dst

    def reset_state(self) -> None:
        """'state_prev = None' for the host-buffer path (crt_filter.py:1765)."""
        self._check(self.lib.crt_reset_state(self.ctx), "crt_reset_state")

    def profile_begin(self, max_samples: int = 8192, every: int = 1) -> None:
        """Start recording CUDA events around the dominant kernel of one frame in `every` (bench.py)."""
        self._check(self.lib.crt_profile_sample_every(self.ctx, int(every)), "crt_profile_sample_every")
        self._check(self.lib.crt_profile_begin(self.ctx, int(max_samples)), "crt_profile_begin")

    def profile_end(self):
        """Returns (summed kernel time in ms, number of launches timed)."""
        ms, n = C.c_double(), C.c_int()
        self._check(self.lib.crt_profile_end(self.ctx, C.byref(ms), C.byref(n)), "crt_profile_end")
        return ms.value, n.value

    def generate_noise(self, frame_index: int):
        torch = _torch()
        gh, gw = tables.noise_plane_shape(self.height, self.width, self.params.grain_size)
        plane = torch.empty((gh, gw), dtype=torch.float32, device=f"cuda:{self.device}")
        stream = torch.cuda.current_stream(plane.device).cuda_stream
        self._check(self.lib.crt_generate_noise(self.ctx, int(frame_index), plane.data_ptr(), C.c_void_p(stream)), "crt_generate_noise")
        return plane

    def generate_glitch(self, frame_index: int = 0, phase_px: Optional[float] = None, fps: float = 30.0):
        """Offsets [rows][segments] of the device glitch generator for one frame.  The pattern is keyed on the frame's
        scanline phase exactly as the reference seeds its generator (crt_filter.py:670, :841); without `phase_px` the
        phase of frame `frame_index` at `fps` is used (process_video's rule, :1043)."""
        torch = _torch()
        p = self.params
        y0, rows, seg_len, nseg = tables.glitch_geometry(self.variant, self.height, self.width, p.glitch_height_frac)
        offs = torch.empty((max(rows, 1), nseg), dtype=torch.int32, device=f"cuda:{self.device}")
        rec = cabi.CrtFrameC()
        rec.frame_index = int(frame_index)
        rec.phase_px = float(phase_px) if phase_px is not None else p.phase_px(int(frame_index), fps)
        stream = torch.cuda.current_stream(offs.device).cuda_stream
        self._check(self.lib.crt_generate_glitch(self.ctx, C.byref(rec), offs.data_ptr(), C.c_void_p(stream)), "crt_generate_glitch")
        return offs
