#!/usr/bin/env python
"""Resize kernel timing (CUDA events): python tools/gpu_resize_bench.py [boards=2048] [in=400] [out=256]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chess_vision_b200.preprocess import resize_boards
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
a = int(sys.argv[2]) if len(sys.argv) > 2 else 400
b = int(sys.argv[3]) if len(sys.argv) > 3 else 256
src = torch.randint(0, 256, (n, a, a, 3), dtype=torch.uint8, device="cuda")
dst = torch.empty((n, b, b, 3), dtype=torch.uint8, device="cuda")
for _ in range(3): resize_boards(src, b, out=dst)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): resize_boards(src, b, out=dst)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
gb = n * (a * a * 3 + b * b * 3) / 1e9
print(f"{n} boards {a}x{a} -> {b}x{b}: {ms:.3f} ms, {n / ms * 1e3:.0f} boards/s, {gb / ms * 1e3:.0f} GB/s algorithmic")
