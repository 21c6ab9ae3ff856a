#!/usr/bin/env python
"""The reference-surface call `model(images)` (float NCHW input, logits out) against the uint8 fast path: python tools/gpu_float_path.py [boards=1024]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_vision_b200 as cv
from chess_vision_b200 import _native
from chess_vision_b200.dataset import NORM_MEAN, NORM_STD
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
model = cv.build_model({"model": {"arch": "square", "pretrained": False}})
model.load_state_dict(bench.make_state(model.state_dict()))
model = model.cuda().eval()
u8 = torch.empty((n, 256, 256, 3), dtype=torch.uint8, device="cuda")
_native.check(_native.lib().cv_synth_boards(_native.ptr(u8), 0, 0, n, 256, 1, 1, None, _native.stream_ptr(u8.device)))
mean = torch.tensor(NORM_MEAN, device="cuda").view(1, 3, 1, 1); std = torch.tensor(NORM_STD, device="cuda").view(1, 3, 1, 1)
x = ((u8.permute(0, 3, 1, 2).float() / 255.0 - mean) / std).contiguous()
def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
with torch.no_grad():
    t_f = timeit(lambda: model(x))
    t_u = timeit(lambda: model.forward_u8(u8))
    model.profile(True); model(x); ms, cnt = model.profile_read(); model.profile(False)
import numpy as np
order = np.argsort(-ms)[:5]
print(f"{n} boards: model(float NCHW) {t_f:.3f} ms ({n / t_f * 1e3:.0f} boards/s), forward_u8 {t_u:.3f} ms ({n / t_u * 1e3:.0f} boards/s)")
print("float path kernels: " + ", ".join(f"{model.PROF_NAMES[i].split('(')[0]} {ms[i]:.3f}" for i in order if ms[i] > 0))
o1, o2 = model(x), model.forward_u8(u8)
print("max |squares diff| float vs u8 path:", float((o1["squares"] - o2["squares"]).abs().max()), "of", float(o2["squares"].abs().max()))
