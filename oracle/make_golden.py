"""Generate tests/golden/chain_golden.npz from the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Run in the build container, where /root/reference exists:

    python -m oracle.make_golden

For every case in oracle/cases.py and both caller variants ("gui" =
apply_crt_effect loop, "export" = apply_static_effects + process_video blend)
the reference's uint8 output of the LAST frame (it carries the accumulated
persistence state) is stored in full, together with a sha256 of every output
frame.  Inputs are not stored: they are regenerated from seeds
(`oracle.cases.case_frames`, `case_noise_seed`, `case_text_layer`).
The file records the numpy / OpenCV versions that produced it, because the
reference pins only lower bounds (requirements.txt).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

from . import harness, ref_loader
from .cases import CASES

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "chain_golden.npz")


def main() -> int:
    if not ref_loader.available():
        print("reference not present; golden vectors can only be generated in the build container", file=sys.stderr)
        return 1
    import cv2
    arrays, meta = {}, {"numpy": np.__version__, "cv2": cv2.__version__, "cases": {}}
    for case in CASES:
        for variant in ("gui", "export"):
            outs, _ = harness.run_reference(case, variant)
            key = f"{case.name}/{variant}"
            twin = arrays.get(f"{case.name}/gui")
            if variant == "export" and twin is not None and np.array_equal(twin, outs[-1]):
                arrays[key] = np.zeros(0, np.uint8)      # identical to the gui entry (glitch off): stored once
            else:
                arrays[key] = outs[-1]
            meta["cases"][key] = {"shape": list(outs[-1].shape), "frames": len(outs),
                                  "sha256": [hashlib.sha256(np.ascontiguousarray(o).tobytes()).hexdigest() for o in outs]}
    arrays["__meta__"] = np.frombuffer(json.dumps(meta, sort_keys=True).encode(), dtype=np.uint8)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **arrays)
    print(f"wrote {OUT}: {len(meta['cases'])} entries, {os.path.getsize(OUT) / 1e6:.2f} MB")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
