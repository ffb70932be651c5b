#!/usr/bin/env python
"""A/B of clip mode (CRT_CLIP=1: one launch per run of up to 64 frames) against one launch per frame (0) with and without
temporal shards, one GPU, fresh process per arm."""
import json, os, subprocess, sys
wls = sys.argv[1:] or ["default1080", "default4k", "default720", "cfg1", "cfg2"]
for wl in wls:
    for clip, shards in (("1", "auto"), ("0", "auto"), ("0", "1")):
        env = dict(os.environ, CRT_CLIP=clip)
        r = subprocess.run([sys.executable, "bench.py", "--workload", wl, "--steps", "5", "--warmup", "3", "--no-also", "--no-cpu", "--no-e2e", "--shards", shards],
                           env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(wl, "clip", clip, "failed", r.stderr[-1500:]); continue
        j = json.loads(line[-1])
        rf = j["roofline"]
        print(wl, f"clip={clip} shards={shards}", "value", round(j["value"]), "single", round(j.get("single_stream", {}).get("value", 0)),
              "kernel", round(rf.get("kernel_avg_ms", 0) * 1e3, 2), "us frac", round(rf.get("frac", 0), 3), "sustained", round(rf.get("frac_sustained", 0), 3),
              "clip_frames", rf.get("clip_mode_frames_per_step"), "launches", j.get("gpu_launches"))
