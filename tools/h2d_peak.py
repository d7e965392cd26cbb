#!/usr/bin/env python
"""Bare host-to-device copy rate of this box (pinned memory, CUDA events): the ceiling of bench.py's e2e leg."""
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.fill_(7)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunk in (n, 1 << 25):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        e0.record()
        for off in range(0, n, chunk):
            d[off:off + chunk].copy_(h[off:off + chunk], non_blocking=True)
        e1.record(); torch.cuda.synchronize()
    print(f"H2D 1 GiB in chunks of {chunk >> 20} MiB: {n / e0.elapsed_time(e1) / 1e6:.2f} GB/s")
