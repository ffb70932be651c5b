#!/usr/bin/env python
"""How much does one B200 gain from running K temporal shards of a clip CONCURRENTLY (one engine + one CUDA stream per
shard) instead of one after the other?  Prints aggregate frames/s for K = 1, 2, 3, 4 per workload.

    python profiles/tools/concurrency_probe.py [workload ...]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import bench  # noqa: E402
from pythoncrt_b200 import CrtEngine  # noqa: E402


def run(name, k, frames_total, reps=3):
    wl = bench.WORKLOADS[name]
    W, H, fps = wl["w"], wl["h"], wl["fps"]
    p = bench.product_params(wl["over"])
    n = frames_total // k
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    shards = []
    for s in range(k):
        eng = CrtEngine(W, H, 0).configure(p, variant="export", noise_mode="generate", glitch_mode="generate", seed=1234)
        fr = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device=dev, generator=g)
        shards.append((eng, fr, torch.empty_like(fr), eng.new_state(), torch.cuda.Stream()))
    chunk = 8            # frames per crt_process call: the host alternates between the streams

    def step():
        for c0 in range(0, n, chunk):
            for s, (eng, fr, out, st, stream) in enumerate(shards):
                with torch.cuda.stream(stream):
                    eng.process(fr[c0:c0 + chunk], out[c0:c0 + chunk], state=st, state_valid=c0 > 0, fps=fps, first_index=s * n + c0)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    for eng, *_ in shards:
        eng.close()
    return k * n * reps / dt


if __name__ == "__main__":
    names = sys.argv[1:] or ["cfg3", "default4k", "cfg2", "cfg1"]
    for name in names:
        total = {"cfg1": 1024, "cfg2": 480, "default4k": 192, "cfg3": 192, "cfg4": 96, "cfg5": 32}.get(name, 192)
        res = {k: round(run(name, k, total), 1) for k in (1, 2, 3, 4)}
        print(json.dumps({"workload": name, "frames_per_s_by_concurrent_shards": res}))
