#!/usr/bin/env python
"""Race hunt: clip mode R times against one launch per frame, for several publication modes; prints mismatching pixels per run."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..")); sys.path.insert(0, os.path.join(HERE, "..", ".."))
import torch
from test_gpu_boundary import _clip_params
from pythoncrt_b200.engine import CrtEngine
n, R = 60, 5
for kind, hw in (("default", (1080, 1920)), ("gauss_grade", (1080, 1920)), ("default", (2160, 3840))):
    h, w = hw
    nn = n if h == 1080 else 16
    p = _clip_params(kind)
    g = torch.Generator(device="cuda").manual_seed(7)
    fr = torch.randint(0, 256, (nn, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
    os.environ["CRT_CLIP"] = "0"
    eng = CrtEngine(w, h).configure(p); eng.set_shards(1)
    ref, ref_state = eng.process(fr, fps=30.0); ref = ref.clone(); ref_state = ref_state.clone(); eng.close()
    os.environ["CRT_CLIP"] = "1"
    for mode in sys.argv[1:] or ["2:1", "2:8", "0:-1", "0:1", "0:2", "1:1"]:      # CRT_CLIP_ITEMS : CRT_CLIP_RELEASE
        os.environ["CRT_CLIP_ITEMS"], os.environ["CRT_CLIP_RELEASE"] = mode.split(":")
        bad = []
        for r in range(R):
            eng = CrtEngine(w, h).configure(p); eng.set_shards(1)
            out, state = eng.process(fr, fps=30.0)
            bad.append(int((out != ref).sum()) + int((state != ref_state).sum()))
            eng.close()
        print(kind, hw, "mode", mode, "mismatching values per run", bad, flush=True)
