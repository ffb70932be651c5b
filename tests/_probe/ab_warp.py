#!/usr/bin/env python
"""A/B of the two cfg3 (warp, 4K) paths on one GPU: CRT_WARP_SRC=1 (source-driven single pass) against 0 (two-pass).  Each arm
runs bench.py in a fresh process (the choice is read when the context is created) and prints value / single-stream / kernel ms."""
import json, os, subprocess, sys
for arm in ("1", "0"):
    env = dict(os.environ, CRT_WARP_SRC=arm)
    r = subprocess.run([sys.executable, "bench.py", "--workload", "cfg3", "--steps", "5", "--warmup", "3", "--no-also", "--no-cpu"], env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        print("arm", arm, "failed", r.stderr[-2000:]); continue
    j = json.loads(line[-1])
    print("cfg3 warp_src=" + arm, round(j["value"]), round(j.get("single_stream", {}).get("value", 0)), j["roofline"].get("kernel_avg_ms"), j["roofline"].get("frac"), j["config"].get("path"), j.get("gpu_launches"))
