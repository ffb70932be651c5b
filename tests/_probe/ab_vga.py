#!/usr/bin/env python
"""VGA (BASELINE configs[0]): TMA-pipelined kernel in low tiles (CRT_TILE_H, CRT_PIPE_MIN_TILES) against the plain block kernel."""
import json, os, subprocess, sys
for th, mt in (("-1", "256"), ("-1", "100"), ("16", "100"), ("8", "100"), ("12", "100")):
    env = dict(os.environ, CRT_TILE_H=th, CRT_PIPE_MIN_TILES=mt, CRT_CLIP="0")
    r = subprocess.run([sys.executable, "bench.py", "--workload", "cfg1", "--steps", "5", "--warmup", "3", "--no-also", "--no-cpu", "--no-e2e"], env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(th, mt, "failed", r.stderr[-600:]); continue
    j = json.loads(line[-1])
    print("cfg1 tile_h", th, "pipe_min_tiles", mt, "value", round(j["value"]), "single", round(j["single_stream"]["value"]), "kernel us", round(j["roofline"]["kernel_avg_ms"] * 1e3, 2))
