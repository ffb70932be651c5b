"""Helpers for the `gpu` tests: run an oracle Case through the CUDA path (C ABI)."""
import json
import os

import numpy as np

import host_emu  # noqa: F401  (only for oracle_to_product_params)
from oracle import harness
from oracle.cases import case_frames, case_text_layer

REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_report.jsonl")


def run_case_gpu(case, variant, policy="auto", batch=0):
    import torch
    from pythoncrt_b200.engine import CrtEngine
    p = host_emu.oracle_to_product_params(case.params)
    eng = CrtEngine(case.w, case.h)
    eng.configure(p, variant=variant, text_rgba=case_text_layer(case), text_after=(case.text != "before"),
                  noise_mode="inject", glitch_mode="inject", policy=policy)
    frames = torch.from_numpy(np.ascontiguousarray(np.stack(case_frames(case)))).cuda()
    planes = harness.noise_planes(case)
    nz = None if planes is None else torch.from_numpy(np.stack(planes)).cuda()
    sc = [harness.frame_scalars(case, j) for j in range(case.frames)]
    phases, times = [s[0] for s in sc], [s[1] for s in sc]
    if batch:
        out = torch.empty_like(frames)
        state, valid = eng.new_state(), False
        for s in range(0, case.frames, batch):
            e = min(case.frames, s + batch)
            eng.process(frames[s:e], out[s:e], state=state, state_valid=valid, phases=phases[s:e], times=times[s:e],
                        first_index=case.first_index + s, noise_planes=None if nz is None else nz[s:e])
            valid = p.persistence > 0
    else:
        out, state = eng.process(frames, phases=phases, times=times, first_index=case.first_index, noise_planes=nz)
    torch.cuda.synchronize()
    fused = int(eng.last_info.fused)
    res = list(out.cpu().numpy()), state.cpu().numpy(), fused
    eng.close()
    return res


def log_report(**kw):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(json.dumps(kw) + "\n")
