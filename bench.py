#!/usr/bin/env python
"""bench.py — frames/s and HBM roofline of the per-frame CRT effect chain.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl ours|reference]

A "step" is one pass of the hot path over one clip of synthetic frames resident
in HBM (BASELINE.json configs; default = configs[1], "1080p full chain with
gaussian bloom + colour grading, 600 frames, 1 B200").  Prints ONE JSON line
(contract in the task brief): value = whole-job frames/s with inputs in HBM,
e2e = the same through the host-buffer C-ABI call (pinned host memory in and
out, copies inside the timed region), roofline = achieved algorithmic GB/s of
all kernels of a frame against the measured HBM peak, cpu_baseline = the
reference's CPU path (the unmodified module staged under baseline/_ref when it
is importable, else the oracle port) timed on this box's host cores.  `also` =
short runs of the other BASELINE configurations (default chain at 4K = the
north_star target, cfg3; cfg4 at N > 1; cfg5 at N = 8) with whole-frame kernel
times.

Multi-GPU (torchrun, one rank per GPU): the clip is sharded temporally, every
rank owns a chunk of `frames` frames preceded by its persistence warm-up halo
(weak scaling); no data-path collective.  Times are CUDA events, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs.  Parameter overrides on the CLI defaults with noise off.
GRADE = dict(brightness=0.05, contrast=1.15, gamma=1.1, saturation=1.1, temperature=0.1)
GAUSS = dict(fast_bloom=False, bloom_sigma=1.5, bloom_threshold=0.7, bloom_strength=0.35)
WARP = dict(warp_strength=0.15, scanline_angle=3.0, scanline_thickness=1.2)
LIVE = dict(noise_strength=1.5, grain_size=2, flicker_strength=0.25, flicker_hz=60.0, glitch_amp_px=16, glitch_height_frac=0.25)
WORKLOADS = {
    "cfg1": dict(desc="CLI default chain, noise off, 640x480", w=640, h=480, frames=64, fps=30.0, over={}, cpu_frames=640),
    "default4k": dict(desc="CLI default chain (scanline+triad+aberration+fast bloom+vignette+persistence, noise off) at 4K — the north_star target chain",
                      w=3840, h=2160, frames=300, fps=30.0, over={}, cpu_frames=24),
    "default720": dict(desc="CLI default chain (as default4k) at 720p", w=1280, h=720, frames=600, fps=30.0, over={}, cpu_frames=96),
    "default1080": dict(desc="CLI default chain (as default4k) at 1080p", w=1920, h=1080, frames=600, fps=30.0, over={}, cpu_frames=96),
    "cfg2": dict(desc="1080p full chain, gaussian bloom sigma 1.5 thr 0.7 + colour grading", w=1920, h=1080, frames=600, fps=30.0,
                 over={**GAUSS, **GRADE}, cpu_frames=96),
    "cfg3": dict(desc="4K, warp 0.15, scanline angle 3.0, aberration, persistence", w=3840, h=2160, frames=600, fps=30.0,
                 over=WARP, cpu_frames=16),
    "cfg4": dict(desc="4K60 full chain incl. noise/grain/flicker/glitch", w=3840, h=2160, frames=600, fps=60.0,
                 over={**GAUSS, **GRADE, **WARP, **LIVE}, cpu_frames=8),
    "cfg5": dict(desc="8K full chain, gaussian bloom sigma 4", w=7680, h=4320, frames=150, fps=30.0,
                 over={**GRADE, **WARP, **LIVE, **dict(fast_bloom=False, bloom_sigma=4.0, bloom_strength=0.3)}, cpu_frames=3),
}
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def alg_bytes_per_px(persistence: float) -> int:
    """uint8 in + uint8 out (+ float32 x3 state read and write when persistence is on)."""
    return 30 if persistence > 0.0 else 6


def bind_near_gpu(index: int):
    """Pin this process to the CPUs NVML reports as local to the GPU (same NUMA node / PCIe root), so that the pinned
    host buffers of the e2e leg are first-touched on that node.  Best effort: any failure leaves the affinity alone."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(index).uuid)).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        near = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = sorted(near & allowed)
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
        return {"near": len(near), "allowed": len(allowed), "bound": len(cpus) if cpus and len(cpus) < len(allowed) else 0}
    except Exception as e:  # noqa: BLE001
        return {"error": type(e).__name__}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4) if r[3 + k].lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def product_params(over):
    from pythoncrt_b200 import CrtParams
    return CrtParams(noise_strength=0.0).but(**over)


def oracle_params(over):
    from oracle.crt_oracle import ChainParams
    return ChainParams(noise_strength=0.0).but(**over)


def make_config(workload: str, wl: dict, world: int) -> dict:
    """`config` of the JSON line — identical for our arm and the reference arm (the latter adds `sampled_frames`)."""
    persistence = wl["over"].get("persistence", 0.2)          # CLI default 0.2 (crt_filter.py:1171); no product import on the reference arm
    return {"workload": workload, "desc": wl["desc"], "width": wl["w"], "height": wl["h"], "frames_per_gpu": wl["frames"], "fps": wl["fps"],
            "parallelism": f"temporal-shards x{world}", "rng": "device counter-based (noise/glitch generated)",
            "l2": "clip (in+out) is larger than L2; no flush needed", "alg_bytes_per_px": alg_bytes_per_px(persistence)}


def run_reference_arm(args, wl):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores — the UNMODIFIED
    reference module (staged under baseline/_ref by oracle/install_ref.py, imported by oracle/ref_loader.py) driven through
    process_video's export loop (2-worker pool + ordered drain, crt_filter.py:1015-1131) when it is importable, else the
    oracle port of the same loop.  Each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import cpu_bench
    p = oracle_params(wl["over"])
    n = wl["cpu_frames"]
    for _ in range(max(0, args.warmup)):
        cpu_bench.time_cpu_path(wl["h"], wl["w"], p, wl["fps"], max(1, n // 4))
    secs, info = 0.0, None
    for _ in range(args.steps):
        info = cpu_bench.time_cpu_path(wl["h"], wl["w"], p, wl["fps"], n)
        secs += info["seconds"]
    fps = n * args.steps / secs
    sample = f"{n} frames of {wl['w']}x{wl['h']} per step ({wl['frames']}-frame workload sampled; CPU path ~{fps:.2f} fps)"
    cfg = make_config(args.workload, wl, max(1, args.gpus))
    cfg["sampled_frames"] = n
    line = {"impl": "reference", "metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": info["cpu_count"], "kind": info["kind"], "sample": sample,
                             "worker_threads": info["workers"], "cv2_threads": info["cv2_threads"], "numpy": info["numpy"], "cv2": info["cv2"]},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


PATHS = {0: "staged", 1: "fused", 2: "two-pass"}
KERNELS = {0: "staged kernels (bloom + pre-warp + output)", 1: "fused tile kernel", 2: "fused first pass + gather (+ generators)"}


def measure(workload: str, wl: dict, steps: int, warmup: int, policy: str, time_every: int, ctx: dict, keep: bool = False, shards="auto"):
    """One device-resident measurement of `workload` on this rank's GPU: W warm-up steps, K timed steps between CUDA
    events, max over ranks.  The kernel time is the duration of ALL kernels of a frame (generators, first pass, gather),
    from CUDA events on the launch stream around one frame in `time_every` (an event pair around every frame would remove
    the frame-to-frame overlap of programmatic dependent launch from `value`).  Returns a dict."""
    import torch
    import torch.distributed as dist
    from pythoncrt_b200 import CrtEngine, clip
    rank, world, local, dev = ctx["rank"], ctx["world"], ctx["local"], ctx["dev"]
    W, H, N, fps = wl["w"], wl["h"], wl["frames"], wl["fps"]
    p = product_params(wl["over"])
    eng = CrtEngine(W, H, local).configure(p, variant="export", policy=policy, noise_mode="generate", glitch_mode="generate", seed=1234, shards=1)
    # this rank's chunk of the global clip: frames [rank*N, (rank+1)*N) preceded by the persistence halo
    halo = clip.halo_frames(p.persistence) if rank > 0 else 0
    first = rank * N - halo
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    frames = torch.randint(0, 256, (N + halo, H, W, 3), dtype=torch.uint8, device=dev, generator=g)
    out = torch.empty_like(frames)
    state = eng.new_state()

    def step():
        eng.process(frames, out, state=state, state_valid=False, fps=fps, first_index=first)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            step()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    bpp = alg_bytes_per_px(p.persistence)
    peak, peak_src = hbm_peak()
    # ---- pass 1: strictly serial (one stream): the per-frame kernel time for the roofline, and the single-stream throughput ----
    for _ in range(max(3, warmup)):
        step()
    sync_all()
    eng.profile_begin(min(16384, (N + halo) * steps), every=time_every)
    serial_ms = timed(steps)
    kern_ms, kern_n = eng.profile_end()
    fused = int(eng.last_info.fused)
    clip_frames = int(eng.last_info.reserved[2])          # frames of a step that went through clip-mode launches (<= 64 frames each)
    serial_value = world * N * steps / (serial_ms * 1e-3)
    kern_avg_ms = kern_ms / max(1, kern_n)
    achieved = (W * H * bpp) / (kern_avg_ms * 1e-3) / 1e9 if kern_n else None
    # ---- pass 2 (the headline `value`): the public API's clip mode — concurrent temporal shards on this GPU (crt_set_shards) ----
    eng.set_shards(shards)
    for _ in range(max(3, warmup)):
        step()
    sync_all()
    launches0 = eng.kernels_launched
    sampler = ClockSampler(local)
    sampler.start()
    total_ms = timed(steps)
    clocks = sampler.stop()
    n_shards, shard_halo = int(eng.last_info.reserved[0]), int(eng.last_info.reserved[1])
    value = world * N * steps / (total_ms * 1e-3)
    sustained = value / world * W * H * bpp / 1e9
    res = {"value": value, "ms_per_step": total_ms / steps, "total_ms": total_ms, "halo": halo, "fused": fused, "bpp": bpp,
           "launches": eng.kernels_launched - launches0, "clocks": clocks, "effective_gbs": value * W * H * bpp / 1e9,
           "intra_gpu_shards": n_shards, "shard_halo_frames": shard_halo,
           "single_stream": {"value": serial_value, "unit": "frames/s", "ms_per_step": serial_ms / steps,
                             "frac_sustained": serial_value / world * W * H * bpp / 1e9 / peak},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                        "traffic": None, "peak_source": peak_src, "kernel": KERNELS.get(fused),
                        "scope": ("clip-mode launches (one launch = a run of up to 64 frames chained tile by tile), timed alone on one stream: "
                                  "launch duration / frames of the launch; algorithmic bytes per launch = bytes per frame x frames")
                        if clip_frames else "all kernels of one frame, timed alone on one stream (event-fenced, no other shard running)",
                        "clip_mode_frames_per_step": clip_frames,
                        "kernel_avg_ms": kern_avg_ms, "kernel_launches_timed": kern_n, "kernel_timed_every": 1 if clip_frames else time_every,
                        "kernel_share_of_step": (kern_ms * (1 if clip_frames else time_every) / serial_ms) if serial_ms else None,
                        "frac_sustained": sustained / peak,
                        "sustained_note": "per-GPU frames/s x algorithmic bytes / peak, with the intra-GPU temporal shards running concurrently"}}
    if keep:
        res.update(eng=eng, frames=frames, out=out)
    else:
        eng.close()
        del frames, out, state
        torch.cuda.empty_cache()
    return res


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=0, help="override frames per clip (debug)")
    ap.add_argument("--policy", default="auto", choices=["auto", "staged", "fused"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the short runs of the other BASELINE configurations")
    ap.add_argument("--shards", default="auto", help="intra-GPU temporal shards of the clip: auto (default), 1 = one stream, k")
    ap.add_argument("--time-every", type=int, default=8, help="CUDA events around the kernels of one frame in this many (1 = every frame)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.frames:
        wl["frames"] = args.frames
    if args.impl == "reference":
        return run_reference_arm(args, wl)

    import numpy as np  # noqa: F401
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    numa = bind_near_gpu(local) if os.environ.get("CRT_BENCH_BIND", "1") != "0" else None
    dev = torch.device(f"cuda:{local}")
    ctx = dict(rank=rank, world=world, local=local, dev=dev)
    W, H, N, fps = wl["w"], wl["h"], wl["frames"], wl["fps"]

    shards = "auto" if args.shards == "auto" else int(args.shards)
    m = measure(args.workload, wl, args.steps, args.warmup, args.policy, args.time_every, ctx, keep=True, shards=shards)
    eng, frames, out, halo = m.pop("eng"), m.pop("frames"), m.pop("out"), m["halo"]
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            m["roofline"]["traffic"] = json.load(f).get(args.workload)
    except Exception:
        pass

    # ---- e2e: host buffers through crt_process_host (pinned in, pinned out) ----
    e2e = None
    if not args.no_e2e:
        import psutil
        need = 2 * N * H * W * 3
        n_e2e = N if psutil.virtual_memory().available > need * 1.5 * world + (8 << 30) else max(8, N // 8)
        if world > 1:                   # every rank pins its own host buffers: keep the total bounded on a shared host
            n_e2e = min(n_e2e, max(150, N // world))
        h_in = torch.empty((n_e2e, H, W, 3), dtype=torch.uint8).pin_memory()
        h_out = torch.empty((n_e2e, H, W, 3), dtype=torch.uint8).pin_memory()
        h_in.copy_(frames[halo:halo + n_e2e].cpu())
        e2e_steps = max(1, min(args.steps, 3))
        eng.set_shards(1)
        eng.reset_state()
        eng.process_host(h_in.numpy(), h_out.numpy(), fps=fps, first_index=rank * N)       # warm-up (allocates the ring)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.reset_state()
            eng.process_host(h_in.numpy(), h_out.numpy(), fps=fps, first_index=rank * N)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(h_out[:4].to(dev), out[halo:halo + 4])) if rank == 0 else True
        e2e = {"value": world * n_e2e * e2e_steps / float(dt.item()), "unit": "frames/s", "h2d_bytes_per_step": n_e2e * H * W * 3,
               "d2h_bytes_per_step": n_e2e * H * W * 3, "frames_per_step": n_e2e, "steps": e2e_steps, "api": "crt_process_host",
               "matches_device_path": same, "cpu_binding": numa}
        ceiling = pcie_ceiling(world)
        if ceiling:
            e2e["ceiling_fps"] = ceiling["gbs_each_way"] * 1e9 / (H * W * 3)
            e2e["ceiling"] = ceiling
        del h_in, h_out
    eng.close()
    del frames, out
    torch.cuda.empty_cache()

    # ---- the other BASELINE configurations, short runs on the same box ----
    also = None
    if not args.no_also and args.workload == "cfg2":
        also = {}
        names = ["default4k", "cfg3"] + (["cfg4"] if world > 1 else []) + (["cfg5"] if world >= 8 else [])
        for name in names:
            w2 = dict(WORKLOADS[name])
            w2["frames"] = min(w2["frames"], {"default4k": 200, "cfg3": 200, "cfg4": 120, "cfg5": 40}[name])
            r = measure(name, w2, 3, 3, "auto", args.time_every, ctx, shards=shards)
            if rank == 0:
                also[name] = {"value": r["value"], "unit": "frames/s", "ms_per_step": r["ms_per_step"], "steps": 3, "warmup": 3,
                              "intra_gpu_shards": r["intra_gpu_shards"], "shard_halo_frames": r["shard_halo_frames"], "single_stream": r["single_stream"],
                              "config": make_config(name, w2, world), "path": PATHS.get(r["fused"]), "kernel_avg_ms": r["roofline"]["kernel_avg_ms"],
                              "frac": r["roofline"]["frac"], "roofline": r["roofline"], "gpu_launches": r["launches"], "clocks": r["clocks"]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import cpu_bench
        info = cpu_bench.time_cpu_path(H, W, oracle_params(wl["over"]), fps, wl["cpu_frames"])
        cpu = {"value": info["fps"], "unit": "frames/s", "cores": info["cpu_count"], "kind": info["kind"],
               "sample": f"{info['frames']} frames of {W}x{H} through the export-like 2-worker loop, {info['seconds']:.1f} s",
               "worker_threads": info["workers"], "cv2_threads": info["cv2_threads"], "numpy": info["numpy"], "cv2": info["cv2"]}

    if rank == 0:
        line = {
            "metric": "frames_per_sec", "value": m["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": make_config(args.workload, wl, world), "path": PATHS.get(m["fused"]), "halo_frames_rank_gt0": None,
            "intra_gpu_shards": m["intra_gpu_shards"], "shard_halo_frames": m["shard_halo_frames"], "single_stream": m["single_stream"],
            "effective_gbs": m["effective_gbs"], "roofline": m["roofline"],
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": m["launches"], "clocks": m["clocks"], "also": also,
        }
        from pythoncrt_b200 import clip as _clip
        line["halo_frames_rank_gt0"] = _clip.halo_frames(product_params(wl["over"]).persistence)
        print(json.dumps(line), flush=True)       # flushed before any teardown: the line must survive whatever the exit path does
    if world > 1:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass
    return 0


def pcie_ceiling(world: int):
    """Aggregate pinned host<->device copy ceiling measured for `world` concurrent ranks (profiles/pcie_ceiling.json, written
    by profiles/tools/pcie_probe.py on the same pool); None when no measurement for this rank count is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "pcie_ceiling.json")) as f:
            return json.load(f).get(str(world))
    except Exception:
        return None


if __name__ == "__main__":
    raise SystemExit(main())
