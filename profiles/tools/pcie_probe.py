#!/usr/bin/env python
"""Host-link ceiling of the box for the e2e leg of bench.py: pinned host <-> device copies, one direction at a time and
both at once (two streams), in 48 MB chunks as crt_process_host moves them — for ONE process, or for N concurrent
ranks (one per GPU) when launched with torchrun:

    python profiles/tools/pcie_probe.py                                  # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        profiles/tools/pcie_probe.py [--json profiles/pcie_ceiling.json]

With N ranks every rank copies at the same time between barriers; the aggregate is N x chunk bytes over the slowest rank's
time.  The e2e leg moves W*H*3 bytes per frame each way, so its ceiling in frames/s is (aggregate GB/s each way) / frame
bytes.  --json merges {"<N>": {...}} into a file bench.py reads (e2e.ceiling_fps)."""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default="")
    ap.add_argument("--chunks", type=int, default=48)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    n, chunk = args.chunks, 48 << 20
    h_in = torch.empty((n, chunk), dtype=torch.uint8).pin_memory()
    h_out = torch.empty((n, chunk), dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d_in = torch.empty((3, chunk), dtype=torch.uint8, device=dev)
    d_out = torch.empty((3, chunk), dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def run(h2d: bool, d2h: bool) -> float:
        sync_all()
        t0 = time.perf_counter()
        for i in range(n):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in[i % 3].copy_(h_in[i], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out[i].copy_(d_out[i % 3], non_blocking=True)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return world * n * chunk / float(dt.item()) / 1e9          # aggregate GB/s in each active direction

    for _ in range(2):
        a, b, c = run(True, False), run(False, True), run(True, True)
    if rank == 0:
        res = {"ranks": world, "h2d_alone_gbs": a, "d2h_alone_gbs": b, "gbs_each_way": c, "chunk_mb": 48, "host_cpus": os.cpu_count(),
               "fps_1080p": c * 1e9 / (1920 * 1080 * 3), "fps_4k": c * 1e9 / (3840 * 2160 * 3)}
        print(json.dumps(res))
        if args.json:
            try:
                with open(args.json) as f:
                    allres = json.load(f)
            except Exception:
                allres = {}
            allres[str(world)] = res
            with open(args.json, "w") as f:
                json.dump(allres, f, indent=1, sort_keys=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
