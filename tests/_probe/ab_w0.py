#!/usr/bin/env python
"""Clip mode of the fast-bloom kernel with warp 0 kept out of phase 1 (CRT_CLIP_W0=1) against the default."""
import json, os, subprocess, sys
for wl in ("default4k", "default1080"):
    for w0 in ("0", "1", "0", "1"):
        env = dict(os.environ, CRT_CLIP_W0=w0)
        r = subprocess.run([sys.executable, "bench.py", "--workload", wl, "--steps", "4", "--warmup", "3", "--no-also", "--no-cpu", "--no-e2e", "--shards", "1"], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(wl, w0, "failed", r.stderr[-600:]); continue
        j = json.loads(line[-1])
        print(wl, "w0", w0, "value", round(j["value"]), "kernel us", round(j["roofline"]["kernel_avg_ms"] * 1e3, 2))
