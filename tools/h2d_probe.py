#!/usr/bin/env python3
"""Probe: pinned H2D / D2H throughput as a function of copy size (many small copies vs one large one)."""
import torch, time
assert torch.cuda.is_available()
tot = 30 * 1024 * 1024
h = torch.empty(tot, dtype=torch.uint8).pin_memory()
d = torch.empty(tot, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
for chunk in (466616, 1 << 20, 4 << 20, tot):
    n = tot // chunk
    for direction in ("h2d", "d2h"):
        best = 1e9
        for rep in range(5):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(st):
                e0.record()
                for i in range(n):
                    a, b = (d, h) if direction == "h2d" else (h, d)
                    a[i * chunk:(i + 1) * chunk].copy_(b[i * chunk:(i + 1) * chunk], non_blocking=True)
                e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print("%s chunk %8d x %3d: %.3f ms  %.1f GB/s" % (direction, chunk, n, best, n * chunk / best / 1e6))
