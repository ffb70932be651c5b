#!/usr/bin/env python
"""A/B of the single-pass block kernels' tile height (CRT_TILE_H: -1 = always 32 rows, 0 = choose_tile_h, else the height) on one
GPU: bench.py in a fresh process per arm; prints sharded value / single-stream value / kernel-alone ms / roofline fraction."""
import json, os, subprocess, sys
arms = sys.argv[1:] or ["-1", "0", "24", "26", "28", "30"]
for wl in ("cfg2", "default1080", "default720"):
    for arm in arms:
        env = dict(os.environ, CRT_TILE_H=arm)
        r = subprocess.run([sys.executable, "bench.py", "--workload", wl, "--steps", "5", "--warmup", "3", "--no-also", "--no-cpu", "--no-e2e"], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(wl, "arm", arm, "failed", r.stderr[-1500:]); continue
        j = json.loads(line[-1])
        print(wl, "tile_h=" + arm, round(j["value"]), round(j.get("single_stream", {}).get("value", 0)), round(j["roofline"].get("kernel_avg_ms", 0) * 1e3, 2), "us", round(j["roofline"].get("frac", 0), 3))
