"""Temporal sharding across ranks (world_size 2, gloo, CPU): the host-side driver
(pythoncrt_b200.clip) with the oracle standing in for the device chain.  Checks
that a rank started `halo` frames early from an empty state reproduces the serial
clip, and that the optional final gather reassembles it in order."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, persistence, q):
    import torch.distributed as dist

    from oracle import crt_oracle as O
    from oracle.cases import BASE
    from pythoncrt_b200 import CrtParams, clip
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    h, w, fps = 48, 64, 30.0
    p = BASE.but(persistence=persistence, warp_strength=0.1)
    prod = CrtParams(persistence=persistence)

    def run_range(first, last, fresh):
        frames = [O.synthetic_frame(i, h, w) for i in range(first, last)]
        outs, _ = O.run_clip(frames, p, fps, first_index=first)
        return torch.from_numpy(np.stack(outs))

    mine, span = clip.process_clip_sharded(run_range, n_frames, prod, rank, world)
    full = clip.gather_frames(mine, span, n_frames, dst=0)
    if rank == 0:
        serial, _ = O.run_clip([O.synthetic_frame(i, h, w) for i in range(n_frames)], p, fps)
        d = np.abs(full.numpy().astype(np.int16) - np.stack(serial).astype(np.int16))
        q.put((int(d.max()), float((d > 0).mean()), tuple(full.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("persistence,n_frames", [(0.2, 21), (0.8, 90)])
def test_two_rank_shards_match_serial(persistence, n_frames):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, persistence, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=240)
        assert pr.exitcode == 0
    worst, frac, shape = q.get(timeout=5)
    assert shape == (n_frames, 48, 64, 3)
    assert worst <= 1 and frac < 1e-3, (worst, frac)
