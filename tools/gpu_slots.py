#!/usr/bin/env python
"""Per-kernel CUDA-event times (cv_square_profile) of the bf16 path: python tools/gpu_slots.py [boards] [mask] [wave]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_vision_b200 as cv
from chess_vision_b200 import _native
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mask = int(sys.argv[2]) if len(sys.argv) > 2 else -1
wave = int(sys.argv[3]) if len(sys.argv) > 3 else 0
model = cv.build_model({"model": {"arch": "square", "pretrained": False}})
model.load_state_dict(bench.make_state(model.state_dict()))
model = model.cuda().eval()
if wave: model.set_wave(wave)
boards = torch.empty((n, 256, 256, 3), dtype=torch.uint8, device="cuda")
_native.check(_native.lib().cv_synth_boards(_native.ptr(boards), 0, 0, n, 256, 1, 1, None, _native.stream_ptr(boards.device)))
if mask >= 0: model.set_impl(mask)
for _ in range(3): model.predict_fen_device(boards)
model.profile(True)
iters = 3
for _ in range(iters): model.predict_fen_device(boards)
ms, cnt = model.profile_read()
model.profile(False)
order = np.argsort(-ms)[:6]
print(f"{n} boards total {ms.sum() / iters:.3f} ms: " + ", ".join(f"{model.PROF_NAMES[i].split('(')[0]} {ms[i] / iters:.3f}" for i in order if ms[i] > 0))
