"""Drop-in replacements for the reference's two per-frame entry points.

    apply_crt_effect      /root/reference/crt_filter.py:531-699  (GUI tick, :1810 / :1972)
    apply_static_effects  /root/reference/crt_filter.py:702-861  (export worker, :1045)

Same positional/keyword arguments, same meaning, same return shapes.  Frames are
numpy uint8 H x W x 3 on the host (possibly read-only, :501); they are copied to
the GPU, run through the C-ABI chain and copied back.  Differences a caller can
observe, all deliberate:

  * the persistence state returned by apply_crt_effect is a `DeviceState` (the
    float32 state kept in HBM) instead of a float ndarray; it has `.shape`,
    converts with `np.asarray(state)` and is accepted back as `state_prev`, which
    is all the reference's callers do with it (:1810-1823, :689);
  * noise uses the device's counter-based generator (the reference's cv2.randn
    stream is thread-local and not reproducible) unless draws
    are injected with the extra keyword `noise_plane=`;
  * glitch offsets are drawn on the host from numpy's PCG64 exactly like the
    reference, so glitch output is identical.

One engine (C-ABI context) is cached per (thread, device, frame size), which is
what makes the reference's two-worker export pool (:1015-1017) safe.
"""
from __future__ import annotations

import threading
from typing import Optional, Tuple

import numpy as np

from . import tables
from .params import CrtParams

_local = threading.local()


class DeviceState:
    """float32 H x W x 3 persistence state resident on the GPU."""

    def __init__(self, tensor):
        self.tensor = tensor

    @property
    def shape(self) -> Tuple[int, ...]:
        return tuple(self.tensor.shape)

    dtype = np.dtype(np.float32)

    def numpy(self) -> np.ndarray:
        return self.tensor.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a


def _engine_for(h: int, w: int, device: int):
    from .engine import CrtEngine
    cache = getattr(_local, "engines", None)
    if cache is None:
        cache = _local.engines = {}
    key = (device, h, w)
    if key not in cache:
        import torch
        eng = CrtEngine(w, h, device)
        eng._cfg_key = None
        eng._pin_in = torch.empty((h, w, 3), dtype=torch.uint8).pin_memory()
        eng._pin_out = torch.empty((h, w, 3), dtype=torch.uint8).pin_memory()
        eng._dev_in = torch.empty((1, h, w, 3), dtype=torch.uint8, device=f"cuda:{device}")
        eng._dev_out = torch.empty((1, h, w, 3), dtype=torch.uint8, device=f"cuda:{device}")
        cache[key] = eng
    return cache[key]


def _mask_key(mask) -> Optional[tuple]:
    if mask is None:
        return None
    s = getattr(mask, "strength", None)
    if s is not None:
        return ("made", float(s), float(getattr(mask, "softness", 0.0)))
    a = np.asarray(mask)
    row = np.ascontiguousarray(a[0] if a.ndim == 3 else a[:: max(1, a.shape[0] // 16)])
    return ("raw", a.shape, hash(row.tobytes()))


def _configure(eng, p: CrtParams, variant: str, triad_mask, vignette_mask, text, text_after, inject_noise: bool):
    key = (p, variant, _mask_key(triad_mask), _mask_key(vignette_mask), None if text is None else hash(np.asarray(text).tobytes()),
           bool(text_after), inject_noise)
    if eng._cfg_key == key:
        return
    triad_cols = None if triad_mask is None else np.asarray(triad_mask)[0]
    eng.configure(p, variant=variant, triad_cols=triad_cols, vignette=vignette_mask, text_rgba=text, text_after=text_after,
                  noise_mode="inject" if inject_noise else "generate", glitch_mode="inject", seed=0x5EED)
    eng._cfg_key = key


def _params(scanline_strength, triad_gamma, triad_preserve_luma, aberration_px, bloom_sigma, bloom_strength, bloom_threshold,
            noise_strength, persistence, scanline_period_px, fast_bloom, pixel_size, glitch_amp_px, glitch_height_frac,
            brightness, contrast, gamma, saturation, temperature, flicker_strength, flicker_hz, grain_size, scanline_angle,
            scanline_thickness, warp_strength) -> CrtParams:
    return CrtParams(
        scanline_strength=float(scanline_strength), triad_gamma=float(triad_gamma), triad_preserve_luma=bool(triad_preserve_luma),
        aberration_px=int(aberration_px), bloom_sigma=float(bloom_sigma), bloom_strength=float(bloom_strength),
        bloom_threshold=float(bloom_threshold), noise_strength=float(noise_strength), persistence=float(persistence),
        scanline_period_px=float(scanline_period_px), fast_bloom=bool(fast_bloom), pixel_size=int(pixel_size),
        glitch_amp_px=int(glitch_amp_px), glitch_height_frac=float(glitch_height_frac), brightness=float(brightness),
        contrast=float(contrast), gamma=float(gamma), saturation=float(saturation), temperature=float(temperature),
        flicker_strength=float(flicker_strength), flicker_hz=float(flicker_hz), grain_size=int(grain_size),
        scanline_angle=float(scanline_angle), scanline_thickness=float(scanline_thickness), warp_strength=float(warp_strength),
        # strengths of the masks travel with the mask arrays, not with the scalars
        triad_strength=0.0, triad_softness=0.0, vignette_strength=0.0, scanline_speed_px_s=0.0)


def _frame_index(time_sec: float, phase_px: float) -> int:
    """Counter of the device noise generator for one drop-in call.  Derived from the arguments that differ from frame
    to frame (time_sec :1064, phase :1043), NOT from a per-thread call counter: the reference's export pool runs this
    function on two workers (:1015-1017), each of which would count 0, 1, 2, ... and hand frames 2k and 2k+1 the same
    grain.  A 63-bit hash of the two doubles' bit patterns: deterministic, and distinct for distinct frames."""
    import hashlib
    import struct
    return int.from_bytes(hashlib.blake2b(struct.pack("<dd", float(time_sec), float(phase_px)), digest_size=8).digest(), "little") >> 1


def _check_text(text_overlay_rgba, h, w):
    if text_overlay_rgba is None:
        return None
    ov = np.asarray(text_overlay_rgba)
    if ov.shape[0] != h or ov.shape[1] != w:
        from PIL import Image   # same host-side resize as the reference (:594)
        if ov.dtype != np.uint8:
            ov = np.clip(ov, 0, 255).astype(np.uint8)
        ov = np.asarray(Image.fromarray(ov, mode="RGBA").resize((w, h), Image.BILINEAR))
    return ov


def apply_crt_effect(frame, scanline_strength, triad_mask, triad_gamma, triad_preserve_luma, aberration_px, bloom_sigma,
                     bloom_strength, bloom_threshold, noise_strength, vignette_mask, persistence, state_prev,
                     scanline_period_px, scanline_phase_px, fast_bloom, pixel_size, glitch_amp_px=0, glitch_height_frac=0.0,
                     time_sec=0.0, brightness=0.0, contrast=1.0, gamma=1.0, saturation=1.0, temperature=0.0,
                     flicker_strength=0.0, flicker_hz=0.0, grain_size=1, scanline_angle=0.0, scanline_thickness=1.0,
                     warp_strength=0.0, text_overlay_rgba=None, text_overlay_after=True, *, noise_plane=None, device: int = 0):
    """GUI chain incl. persistence and uint8 quantise; returns (uint8 H x W x 3, DeviceState)."""
    import torch
    frame = np.asarray(frame)
    h, w = frame.shape[0], frame.shape[1]
    eng = _engine_for(h, w, device)
    p = _params(scanline_strength, triad_gamma, triad_preserve_luma, aberration_px, bloom_sigma, bloom_strength, bloom_threshold,
                noise_strength, persistence, scanline_period_px, fast_bloom, pixel_size, glitch_amp_px, glitch_height_frac,
                brightness, contrast, gamma, saturation, temperature, flicker_strength, flicker_hz, grain_size, scanline_angle,
                scanline_thickness, warp_strength)
    text = _check_text(text_overlay_rgba, h, w)
    _configure(eng, p, "gui", triad_mask, vignette_mask, text, text_overlay_after, noise_plane is not None)
    # state handling (:687-694): blend only when state_prev is given and persistence > 0
    if isinstance(state_prev, DeviceState) and state_prev.shape == (h, w, 3):
        state, valid = state_prev.tensor, True
    elif state_prev is not None:
        prev = state_prev.tensor if isinstance(state_prev, DeviceState) else \
            torch.from_numpy(np.ascontiguousarray(np.asarray(state_prev, dtype=np.float32))).to(eng._dev_in.device)
        if tuple(prev.shape) != (h, w, 3):       # a stale state of another preview size: cv2.resize(INTER_LINEAR) (:689-690), on the device
            prev = eng.resize_state(prev)
        state, valid = prev, True
    else:
        state, valid = eng.new_state(), False
    eng._pin_in.numpy()[...] = frame
    eng._dev_in[0].copy_(eng._pin_in, non_blocking=True)
    kw = dict(phases=[float(scanline_phase_px)], times=[float(time_sec)], first_index=_frame_index(time_sec, scanline_phase_px))
    if noise_plane is not None:
        kw["noise_planes"] = torch.from_numpy(np.ascontiguousarray(noise_plane, np.float32)).to(eng._dev_in.device)[None]
    eng.process(eng._dev_in, eng._dev_out, state=state, state_valid=valid, **kw)
    eng._pin_out.copy_(eng._dev_out[0], non_blocking=True)
    torch.cuda.current_stream(eng._dev_in.device).synchronize()
    return eng._pin_out.numpy().copy(), DeviceState(state)


def _static_setup(hw, args, kw, noise_plane, device, persistence=0.0):
    """This thread's engine, configured for one export-chain call on frames of hw = (height, width)."""
    (scanline_strength, triad_mask, triad_gamma, triad_preserve_luma, aberration_px, bloom_sigma, bloom_strength, bloom_threshold,
     noise_strength, vignette_mask, scanline_period_px, scanline_phase_px, fast_bloom, pixel_size, glitch_amp_px, glitch_height_frac) = args
    h, w = int(hw[0]), int(hw[1])
    eng = _engine_for(h, w, device)
    p = _params(scanline_strength, triad_gamma, triad_preserve_luma, aberration_px, bloom_sigma, bloom_strength, bloom_threshold,
                noise_strength, persistence, scanline_period_px, fast_bloom, pixel_size, glitch_amp_px, glitch_height_frac,
                kw["brightness"], kw["contrast"], kw["gamma"], kw["saturation"], kw["temperature"], kw["flicker_strength"],
                kw["flicker_hz"], kw["grain_size"], kw["scanline_angle"], kw["scanline_thickness"], kw["warp_strength"])
    text = _check_text(kw["text_overlay_rgba"], h, w)
    _configure(eng, p, "export", triad_mask, vignette_mask, text, kw["text_overlay_after"], noise_plane is not None)
    return eng


def apply_static_effects(frame, scanline_strength, triad_mask, triad_gamma, triad_preserve_luma, aberration_px, bloom_sigma,
                         bloom_strength, bloom_threshold, noise_strength, vignette_mask, scanline_period_px, scanline_phase_px,
                         fast_bloom, pixel_size, glitch_amp_px, glitch_height_frac, time_sec=0.0, brightness=0.0, contrast=1.0,
                         gamma=1.0, saturation=1.0, temperature=0.0, flicker_strength=0.0, flicker_hz=0.0, grain_size=1,
                         scanline_angle=0.0, scanline_thickness=1.0, warp_strength=0.0, text_overlay_rgba=None,
                         text_overlay_after=True, *, noise_plane=None, device: int = 0) -> np.ndarray:
    """Export chain, stateless; returns the float32 image that process_video blends (:1084-1098)."""
    import torch
    args = (scanline_strength, triad_mask, triad_gamma, triad_preserve_luma, aberration_px, bloom_sigma, bloom_strength, bloom_threshold,
            noise_strength, vignette_mask, scanline_period_px, scanline_phase_px, fast_bloom, pixel_size, glitch_amp_px, glitch_height_frac)
    kw = dict(brightness=brightness, contrast=contrast, gamma=gamma, saturation=saturation, temperature=temperature,
              flicker_strength=flicker_strength, flicker_hz=flicker_hz, grain_size=grain_size, scanline_angle=scanline_angle,
              scanline_thickness=scanline_thickness, warp_strength=warp_strength, text_overlay_rgba=text_overlay_rgba,
              text_overlay_after=text_overlay_after)
    frame = np.asarray(frame)
    eng = _static_setup(frame.shape, args, kw, noise_plane, device)
    eng._pin_in.numpy()[...] = frame
    eng._dev_in[0].copy_(eng._pin_in, non_blocking=True)
    fkw = dict(phases=[float(scanline_phase_px)], times=[float(time_sec)], first_index=_frame_index(time_sec, scanline_phase_px))
    if noise_plane is not None:
        fkw["noise_planes"] = torch.from_numpy(np.ascontiguousarray(noise_plane, np.float32)).to(eng._dev_in.device)[None]
    img = eng.process_static(eng._dev_in, **fkw)
    return img[0].cpu().numpy()


# ---- export pipeline on the device, behind the UNMODIFIED process_video (:864-1150) ----------------------------------
# process_video calls apply_static_effects on pool threads (:1045) and then, in frame order on its main thread, does
#     blended = np.clip(persistence * prev_state + (1.0 - persistence) * static_img, 0.0, 1.0)        (:1092 / :1118)
#     prev_state = blended ; out_frame = cv2.convertScaleAbs(blended, alpha=255.0, beta=0)            (:1096-1098)
# That block is inline code, not a function, so it cannot be replaced; but every operation in it dispatches on its
# operands.  install(lazy_export=True) makes apply_static_effects return an ExportFrame — the uploaded uint8 frame plus its
# per-frame scalars, nothing computed yet.  `p * prev + (1 - p) * frame` builds a _Blend expression, np.clip(...) (through
# numpy's __array_function__ protocol) returns the frame with the blend recorded, and the cv2 proxy's convertScaleAbs runs
# ONE crt_process call: chain + persistence against the state kept in HBM + quantise, and reads back 3 bytes per pixel.
# Per frame the host link carries 3 B/px up and 3 B/px down instead of 3 up and 12 (float32 image) down, and no pixel
# arithmetic is left on the CPU.  An ExportFrame still converts to the float image (np.asarray) for any other consumer.
TRANSFER_LOG = {"h2d_bytes": 0, "d2h_bytes": 0, "frames": 0}        # counted by the lazy export path (tests read it)


class _Scaled:
    def __init__(self, obj, k):
        self.obj, self.k = obj, float(k)

    def __add__(self, other):
        if isinstance(other, _Scaled):
            a, b = (self, other) if isinstance(other.obj, ExportFrame) and other.obj.state is None else (other, self)
            return _Blend(prev=a.obj, p=a.k, frame=b.obj, q=b.k)
        return NotImplemented

    __radd__ = __add__


class _Blend:
    """persistence * prev_state + (1 - persistence) * static_img, not yet evaluated."""
    def __init__(self, prev, p, frame, q):
        self.prev, self.p, self.frame, self.q = prev, p, frame, q

    def __array_function__(self, func, types, args, kwargs):
        if func is np.clip and len(args) >= 3 and args[0] is self and float(args[1]) == 0.0 and float(args[2]) == 1.0:
            self.frame.blend_with = (self.prev, self.p)
            return self.frame
        return NotImplemented


class ExportFrame:
    """What the lazy apply_static_effects returns: one uploaded frame of the export chain, evaluated on demand."""

    def __init__(self, eng, dev_in, args, kw, noise_plane, device, phase, time_sec):
        self.eng, self.dev_in, self.args, self.kw, self.noise_plane, self.device = eng, dev_in, args, kw, noise_plane, device
        self.phase, self.time_sec = float(phase), float(time_sec)
        self.shape = (eng.height, eng.width, 3)
        self.dtype = np.dtype(np.float32)
        self.blend_with = None          # (previous ExportFrame | float ndarray, persistence)
        self.state = None               # device float32 state after this frame (what `prev_state = blended` keeps)
        self.out_u8 = None

    # persistence * prev_state  /  (1 - persistence) * static_img
    def __rmul__(self, k):
        return _Scaled(self, k)

    __mul__ = __rmul__
    __array_ufunc__ = None              # numpy scalars defer to __rmul__ instead of converting this to an array

    def _frame_kw(self):
        import torch
        fkw = dict(phases=[self.phase], times=[self.time_sec], first_index=_frame_index(self.time_sec, self.phase))
        if self.noise_plane is not None:
            fkw["noise_planes"] = torch.from_numpy(np.ascontiguousarray(self.noise_plane, np.float32)).to(self.dev_in.device)[None]
        return fkw

    def __array__(self, dtype=None, copy=None):
        """The float32 image apply_static_effects returns (:861), for consumers other than process_video's drain."""
        eng = _static_setup(self.shape, self.args, self.kw, self.noise_plane, self.device)
        img = eng.process_static(self.dev_in, **self._frame_kw())[0].cpu().numpy()
        return img.astype(dtype) if dtype is not None else img

    def quantised(self) -> np.ndarray:
        """blend (if one was recorded) + convertScaleAbs(alpha=255) on the device; returns uint8 H x W x 3."""
        import torch
        if self.out_u8 is not None:
            return self.out_u8
        h, w, _ = self.shape
        prev, p = self.blend_with if self.blend_with is not None else (None, 0.0)
        eng = _static_setup(self.shape, self.args, self.kw, self.noise_plane, self.device, persistence=p)
        state, valid = eng.new_state(), False
        if prev is not None:
            if isinstance(prev, ExportFrame) and prev.state is not None and tuple(prev.state.shape) == self.shape:
                state, valid = prev.state, True           # in place: the previous frame's state becomes this frame's
            else:
                arr = np.asarray(prev.state.cpu().numpy() if isinstance(prev, ExportFrame) and prev.state is not None else prev, np.float32)
                if arr.shape != self.shape:               # the reference's stale-state rule (:1088-1091): through uint8, PIL bilinear
                    from PIL import Image
                    im = Image.fromarray(np.clip(arr * 255.0, 0, 255).astype(np.uint8)).resize((w, h), Image.BILINEAR)
                    arr = np.asarray(im).astype(np.float32) / 255.0
                state, valid = torch.from_numpy(np.ascontiguousarray(arr)).to(self.dev_in.device), True
        eng.process(self.dev_in, eng._dev_out, state=state, state_valid=valid, **self._frame_kw())
        eng._pin_out.copy_(eng._dev_out[0], non_blocking=True)
        torch.cuda.current_stream(self.dev_in.device).synchronize()
        TRANSFER_LOG["d2h_bytes"] += eng._pin_out.numel()
        TRANSFER_LOG["frames"] += 1
        self.state, self.out_u8 = state, eng._pin_out.numpy().copy()
        self.dev_in = None                              # the uploaded frame is no longer needed
        return self.out_u8


def apply_static_effects_lazy(frame, scanline_strength, triad_mask, triad_gamma, triad_preserve_luma, aberration_px, bloom_sigma,
                              bloom_strength, bloom_threshold, noise_strength, vignette_mask, scanline_period_px, scanline_phase_px,
                              fast_bloom, pixel_size, glitch_amp_px, glitch_height_frac, time_sec=0.0, brightness=0.0, contrast=1.0,
                              gamma=1.0, saturation=1.0, temperature=0.0, flicker_strength=0.0, flicker_hz=0.0, grain_size=1,
                              scanline_angle=0.0, scanline_thickness=1.0, warp_strength=0.0, text_overlay_rgba=None,
                              text_overlay_after=True, *, noise_plane=None, device: int = 0) -> "ExportFrame":
    """apply_static_effects for the export pool (:1045) when the drain runs on the device too: uploads the frame and returns
    an ExportFrame; nothing else happens on the worker thread."""
    import torch
    args = (scanline_strength, triad_mask, triad_gamma, triad_preserve_luma, aberration_px, bloom_sigma, bloom_strength, bloom_threshold,
            noise_strength, vignette_mask, scanline_period_px, scanline_phase_px, fast_bloom, pixel_size, glitch_amp_px, glitch_height_frac)
    kw = dict(brightness=brightness, contrast=contrast, gamma=gamma, saturation=saturation, temperature=temperature,
              flicker_strength=flicker_strength, flicker_hz=flicker_hz, grain_size=grain_size, scanline_angle=scanline_angle,
              scanline_thickness=scanline_thickness, warp_strength=warp_strength, text_overlay_rgba=text_overlay_rgba,
              text_overlay_after=text_overlay_after)
    frame = np.asarray(frame)
    h, w = frame.shape[0], frame.shape[1]
    eng = _engine_for(h, w, device)                     # this worker thread's engine: its pinned staging buffer
    eng._pin_in.numpy()[...] = frame
    dev_in = torch.empty((1, h, w, 3), dtype=torch.uint8, device=f"cuda:{device}")
    dev_in[0].copy_(eng._pin_in, non_blocking=True)
    torch.cuda.current_stream(dev_in.device).synchronize()      # the pinned buffer is reused by this thread's next frame
    TRANSFER_LOG["h2d_bytes"] += frame.size
    return ExportFrame(eng, dev_in, args, kw, noise_plane, device, scanline_phase_px, time_sec)


class _Cv2Proxy:
    """Stands in for the `cv2` module global of the patched reference module: convertScaleAbs(ExportFrame, alpha=255)
    (:1098 / :1124) finishes the frame on the device; everything else is the real cv2."""

    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        return getattr(self._real, name)

    def convertScaleAbs(self, src, *a, **k):
        if isinstance(src, ExportFrame):
            alpha = k.get("alpha", a[1] if len(a) > 1 else 1.0)      # positional order: (src, dst, alpha, beta)
            beta = k.get("beta", a[2] if len(a) > 2 else 0.0)
            if float(alpha) == 255.0 and float(beta) == 0.0:
                return src.quantised()
            src = np.asarray(src)
        return self._real.convertScaleAbs(src, *a, **k)


def install(reference_module, lazy_export: bool = True) -> None:
    """Monkey-patch an imported reference module (crt_filter) so that its unmodified GUI (:1810, :1972) and export
    pipeline (:1045, :1081-1124) run the chain on the GPU.  With lazy_export the export drain's blend and quantise
    (:1086-1098) also run on the device (see ExportFrame); without it apply_static_effects returns the float image
    and the reference's own numpy/cv2 code blends it on the host."""
    reference_module.apply_crt_effect = apply_crt_effect
    reference_module.apply_static_effects = apply_static_effects_lazy if lazy_export else apply_static_effects
    reference_module.make_triad_mask = tables.make_triad_mask
    reference_module.make_vignette = tables.make_vignette_lazy       # no H x W host array per preview tick (:1821)
    if lazy_export and hasattr(reference_module, "cv2") and not isinstance(reference_module.cv2, _Cv2Proxy):
        reference_module.cv2 = _Cv2Proxy(reference_module.cv2)
