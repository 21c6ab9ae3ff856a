#!/usr/bin/env python
"""Host->device copy bandwidth from pinned memory (what bounds the e2e figure): one big copy and 512-board chunks."""
import time, torch
n = 4096 * 256 * 256 * 3
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for label, chunks in (("1 copy", 1), ("8 chunks", 8), ("32 chunks", 32)):
    step = n // chunks
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for c in range(chunks):
            d[c * step:(c + 1) * step].copy_(h[c * step:(c + 1) * step], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"H2D {label}: {n / dt / 1e9:.1f} GB/s ({dt * 1e3:.1f} ms for {n / 1e6:.0f} MB)")
s2 = torch.cuda.Stream()
half = n // 2
torch.cuda.synchronize(); t0 = time.perf_counter()
d[:half].copy_(h[:half], non_blocking=True)
with torch.cuda.stream(s2):
    d[half:].copy_(h[half:], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"H2D 2 streams: {n / dt / 1e9:.1f} GB/s")
o = torch.empty(n, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter(); o.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"D2H 1 copy: {n / dt / 1e9:.1f} GB/s")
