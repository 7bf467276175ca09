#!/usr/bin/env python
"""How much of the 126 MB L2 does a buffer read by ALL SMs get?  Re-reads a buffer of N MB many times (torch sum over
float32, every SM touches every part of it over time) and prints the achieved read bandwidth: far above the HBM rate while
the buffer stays L2-resident, the HBM rate beyond.  The knee is the effective capacity for data shared by both dies."""
import torch

torch.cuda.set_device(0)
for mb in (16, 32, 48, 56, 64, 72, 80, 96, 112, 128, 160, 256, 1024):
    n = mb * (1 << 20) // 4
    x = torch.ones(n, dtype=torch.float32, device="cuda")
    for _ in range(5):
        x.sum()
    torch.cuda.synchronize()
    reps = 40
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        x.sum()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{mb:5d} MB: {ms*1e3:8.1f} us per pass, {mb * (1 << 20) / ms / 1e6:8.1f} GB/s", flush=True)
