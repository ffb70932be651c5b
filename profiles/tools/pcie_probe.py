#!/usr/bin/env python
"""PCIe ceiling of the box: pinned host <-> device copies, one direction at a time and both at once
(two streams), 48 MB chunks as crt_process_host uses.  Prints GB/s; the e2e leg of bench.py moves
W*H*3 bytes per frame each way, so its ceiling in frames/s is bidirectional GB/s / frame bytes."""
import time

import torch


def main():
    n, chunk = 64, 48 << 20
    h_in = torch.empty((n, chunk), dtype=torch.uint8).pin_memory()
    h_out = torch.empty((n, chunk), dtype=torch.uint8).pin_memory()
    d_in = torch.empty((3, chunk), dtype=torch.uint8, device="cuda")
    d_out = torch.empty((3, chunk), dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d: bool, d2h: bool) -> float:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in[i % 3].copy_(h_in[i], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out[i].copy_(d_out[i % 3], non_blocking=True)
        torch.cuda.synchronize()
        return n * chunk / (time.perf_counter() - t0) / 1e9

    for _ in range(2):
        a, b, c = run(True, False), run(False, True), run(True, True)
    print(f"H2D alone {a:.1f} GB/s, D2H alone {b:.1f} GB/s, both at once {c:.1f} GB/s each way")
    for name, fb in (("1080p", 1920 * 1080 * 3), ("4K", 3840 * 2160 * 3)):
        print(f"  ceiling for {name} frames through host buffers: {c * 1e9 / fb:.0f} frames/s")


if __name__ == "__main__":
    main()
