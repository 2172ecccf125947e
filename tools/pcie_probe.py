#!/usr/bin/env python3
"""tools/pcie_probe.py -- what the host link of this box can do with pinned memory: H2D alone, D2H alone, and both at
once on two streams (the e2e figure of bench.py moves 4 B/sample in and 8 B/sample out, concurrently)."""
import torch

n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(2 * n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        s1.synchronize(); s2.synchronize()
        e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


t = timed(h2d); print("H2D alone   %.1f GB/s" % (n / t / 1e9))
t = timed(d2h); print("D2H alone   %.1f GB/s" % (2 * n / t / 1e9))
t = timed(both); print("both        %.1f GB/s in + %.1f GB/s out  (=> %.2f Gsamples/s at 4 B in + 8 B out)" % (n / t / 1e9, 2 * n / t / 1e9, (n / 4) / t / 1e9))
