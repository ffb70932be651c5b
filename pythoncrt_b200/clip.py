"""Clip driver: frames resident in HBM, optional temporal sharding over GPUs.

Replaces the frame loop of process_video (/root/reference/crt_filter.py:1037-1131)
for frames that are already on the device.  The only cross-frame dependency of
the whole chain is the first-order persistence recurrence
    s_t = clip(p * s_{t-1} + (1 - p) * x_t)                       (:1092)
whose memory of the past decays as p^k.  A clip can therefore be cut into
contiguous temporal chunks, one per GPU: chunk g > 0 starts `halo` frames early
with an empty state, throws those outputs away and then matches the serial run
to within eps = p^halo.  No collective is needed; finished
frames are optionally gathered (torch.distributed) for the host encoder.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Tuple

from .params import CrtParams

EIGHTH_LSB = 1.0 / 2040.0


def iter_rawvideo(stream, width: int, height: int):
    """Frames of an `rgb24` rawvideo byte stream, as FFmpegPipeReader.iter_frames yields them
    (/root/reference/crt_filter.py:494-502): W*H*3 bytes per frame, a short read ends the clip."""
    import numpy as np
    frame_size = width * height * 3
    while True:
        buf = stream.read(frame_size)
        if not buf or len(buf) < frame_size:
            break
        yield np.frombuffer(buf, dtype=np.uint8).reshape((height, width, 3))


def process_rawvideo(engine, src, dst, *, fps: float = 30.0, first_index: int = 0, chunk_frames: int = 0, pinned: bool = True) -> int:
    """Stream an `rgb24` rawvideo clip (the byte format of the reference's ffmpeg pipes: reader
    /root/reference/crt_filter.py:484-502, writer write_frame :1101) through a configured
    CrtEngine: `src.read` -> pinned host chunk -> crt_process_host (copies and kernels overlap
    inside the library) -> `dst.write`.  The persistence state is carried across chunks on the
    device, frame indices continue from `first_index`.  Returns the number of frames written."""
    import numpy as np
    h, w = engine.height, engine.width
    fbytes = h * w * 3
    if chunk_frames <= 0:
        chunk_frames = max(1, min(256, (192 << 20) // fbytes))
    if pinned:
        import torch
        h_in = torch.empty((chunk_frames, h, w, 3), dtype=torch.uint8).pin_memory().numpy()
        h_out = torch.empty((chunk_frames, h, w, 3), dtype=torch.uint8).pin_memory().numpy()
    else:
        h_in = np.empty((chunk_frames, h, w, 3), np.uint8)
        h_out = np.empty_like(h_in)
    engine.reset_state()
    done = 0
    frames = iter_rawvideo(src, w, h)
    while True:
        n = 0
        for fr in frames:
            h_in[n] = fr
            n += 1
            if n == chunk_frames:
                break
        if n == 0:
            break
        engine.process_host(h_in[:n], h_out[:n], fps=fps, first_index=first_index + done)
        dst.write(h_out[:n].tobytes() if not hasattr(dst, "write_frames") else h_out[:n])
        done += n
        if n < chunk_frames:
            break
    return done


def halo_frames(persistence: float, eps: float = EIGHTH_LSB) -> int:
    """Warm-up frames K with persistence^K <= eps (0 when persistence is off)."""
    p = float(persistence)
    if p <= 0.0:
        return 0
    if p >= 1.0:
        raise ValueError("persistence must be < 1")
    return int(math.ceil(math.log(eps) / math.log(p)))


def shard_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of `rank`; the first n % world ranks get one extra frame."""
    base, extra = divmod(int(n_frames), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_plan(n_frames: int, rank: int, world: int, persistence: float, eps: float = EIGHTH_LSB) -> Tuple[int, int, int]:
    """(warm_start, start, stop): process [warm_start, stop), keep [start, stop)."""
    start, stop = shard_range(n_frames, rank, world)
    return max(0, start - halo_frames(persistence, eps)), start, stop


def process_clip(engine, frames, *, fps: float = 30.0, first_index: int = 0, out=None, state=None, state_valid: bool = False,
                 batch: int = 0, **frame_kw):
    """Run device-resident frames [N][H][W][3] through a configured CrtEngine in
    frame order, carrying the persistence state.  Returns (out, state)."""
    import torch
    n = frames.shape[0]
    if out is None:
        out = torch.empty_like(frames)
    if state is None:
        state, state_valid = engine.new_state(), False
    step = n if batch <= 0 else batch
    for s in range(0, n, step):
        e = min(n, s + step)
        kw = {k: (v[s:e] if v is not None and hasattr(v, "__getitem__") else v) for k, v in frame_kw.items()}
        engine.process(frames[s:e], out[s:e], state=state, state_valid=state_valid, fps=fps, first_index=first_index + s, **kw)
        state_valid = engine.params.persistence > 0.0
    return out, state


def process_clip_sharded(run_range: Callable[[int, int, bool], object], n_frames: int, params: CrtParams, rank: int, world: int,
                         eps: float = EIGHTH_LSB):
    """Temporal sharding driver.  `run_range(first, last, fresh_state)` must process
    global frames [first, last) in order starting from an empty persistence state
    and return their outputs indexable as [k] (k = 0 is frame `first`).  Returns
    this rank's outputs for its own frames [start, stop) and (start, stop)."""
    warm, start, stop = shard_plan(n_frames, rank, world, params.persistence, eps)
    if stop <= start:
        return None, (start, stop)
    outs = run_range(warm, stop, True)
    return outs[start - warm:], (start, stop)


def gather_frames(local_out, start_stop: Tuple[int, int], n_frames: int, dst: int = 0):
    """Optional final gather of finished uint8 frames to rank `dst` (the host
    encoder's rank) with torch.distributed; returns the full clip on dst, None elsewhere."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return local_out
    counts = [shard_range(n_frames, r, world) for r in range(world)]
    longest = max(b - a for a, b in counts)
    pad = torch.zeros((longest,) + tuple(local_out.shape[1:]), dtype=local_out.dtype, device=local_out.device)
    pad[: local_out.shape[0]] = local_out
    bucket = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bucket, dst=dst)
    if rank != dst:
        return None
    return torch.cat([bucket[r][: counts[r][1] - counts[r][0]] for r in range(world)], dim=0)
