#!/usr/bin/env python
"""Where do clip mode and one launch per frame differ (if they do)?"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..")); sys.path.insert(0, os.path.join(HERE, "..", ".."))
import torch
from test_gpu_boundary import _clip_params
from pythoncrt_b200.engine import CrtEngine
kind = sys.argv[1] if len(sys.argv) > 1 else "default"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
p = _clip_params(kind)
g = torch.Generator(device="cuda").manual_seed(99)
fr = torch.randint(0, 256, (n, 1080, 1920, 3), dtype=torch.uint8, device="cuda", generator=g)
res = []
for clip in ("0", "1", "1"):
    os.environ["CRT_CLIP"] = clip
    eng = CrtEngine(1920, 1080).configure(p)
    eng.set_shards(1)
    out, state = eng.process(fr, fps=30.0)
    torch.cuda.synchronize()
    res.append((out.clone(), state.clone(), int(eng.last_info.reserved[2])))
    eng.close()
for a, b, name in ((0, 1, "frame-vs-clip"), (1, 2, "clip-vs-clip")):
    d = (res[a][0].to(torch.int16) - res[b][0].to(torch.int16)).abs()
    per = d.flatten(1).amax(1).tolist()
    cnt = (d > 0).flatten(1).sum(1).tolist()
    print(name, "clip frames", res[a][2], res[b][2], "max per frame", per[:12], "count", cnt[:12], "state max", float((res[a][1] - res[b][1]).abs().max()))
    bad = (d > 0).nonzero()
    if len(bad):
        print(" first bad", bad[:5].tolist(), "rows", sorted(set(bad[:, 1].tolist()))[:20])
print([eng.params.phase_px(i, 30.0) for i in range(4)])
