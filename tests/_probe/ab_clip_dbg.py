#!/usr/bin/env python
"""Timing of the clip-mode protocol's parts (CRT_CLIP_DBG: 8 = no flags at all [results wrong], 16 = static items instead of the atomic
counter [only safe alone on the GPU]), default chain at 4K and 1080p."""
import json, os, subprocess, sys
for wl in ("default4k", "default1080", "cfg2"):
    for items, rel in (("2", "1"), ("3", "1")):
        env = dict(os.environ, CRT_CLIP_ITEMS=items, CRT_CLIP_RELEASE=rel)
        r = subprocess.run([sys.executable, "bench.py", "--workload", wl, "--steps", "4", "--warmup", "3", "--no-also", "--no-cpu", "--no-e2e", "--shards", "1"], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(wl, items, rel, "failed", r.stderr[-800:]); continue
        j = json.loads(line[-1])
        print(wl, "items", items, "release", rel, "value", round(j["value"]), "kernel us", round(j["roofline"]["kernel_avg_ms"] * 1e3, 2), "frac", round(j["roofline"]["frac"], 3))
