#!/usr/bin/env python
"""Markdown table of a set of bench.py JSON lines:  python profiles/tools/make_table.py profiles/r02/bench_*_final.json"""
import json
import sys

rows = []
for path in sys.argv[1:]:
    try:
        d = json.load(open(path))
    except Exception as e:  # noqa: BLE001
        print(f"<!-- {path}: {e} -->")
        continue
    r, c = d["roofline"], d["config"]
    e2e = d.get("e2e") or {}
    cpu = d.get("cpu_baseline") or {}
    rows.append((c["workload"], f"{c['width']}x{c['height']}", d.get("path"), d["value"], d.get("intra_gpu_shards"), d["single_stream"]["value"],
                 r["kernel_avg_ms"] * 1e3, r["frac"], r["frac_sustained"], e2e.get("value"), e2e.get("ceiling_fps"), cpu.get("value"), cpu.get("kind"),
                 r.get("clip_mode_frames_per_step")))
    for name, a in (d.get("also") or {}).items():
        ar = a["roofline"]
        rows.append((name + " (also)", f"{a['config']['width']}x{a['config']['height']}", a.get("path"), a["value"], a.get("intra_gpu_shards"),
                     a["single_stream"]["value"], ar["kernel_avg_ms"] * 1e3, ar["frac"], ar["frac_sustained"], None, None, None, None,
                     ar.get("clip_mode_frames_per_step")))
print("| workload | frame | path | frames/s (clip in HBM, automatic mode) | shards | frames/s one stream | one stream is | kernels of a frame, alone (µs) | roofline frac (alone) | roofline frac (sustained) | end to end frames/s | host-link ceiling | CPU reference frames/s |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
f = lambda v, p=0: "–" if v is None else (f"{v:,.{p}f}".replace(",", " "))
for w, fr, path, v, k, s1, us, fa, fs, e, ce, cp, kind, clipf in rows:
    print(f"| {w} | {fr} | {path} | {f(v)} | {k} | {f(s1)} | {'clip mode' if clipf else 'launch per frame'} | {f(us, 1)} | {f(fa, 3)} | {f(fs, 3)} | {f(e)} | {f(ce)} | {f(cp, 2)}{' (' + kind + ')' if kind else ''} |")
