#!/usr/bin/env python
"""Small fixed workload for ncu captures: N boards through predict_fen_device a few times.
    python tools/gpu_profile_run.py [boards=512] [iters=3] [mask=-1] [wave=0]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_vision_b200 as cv
from chess_vision_b200 import _native, synthetic
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mask = int(sys.argv[3]) if len(sys.argv) > 3 else -1
wave = int(sys.argv[4]) if len(sys.argv) > 4 else 0
model = cv.build_model({"model": {"arch": "square", "pretrained": False}})
model.load_state_dict(bench.make_state(model.state_dict()))
model = model.cuda().eval()
if wave:
    model.set_wave(wave)
boards = torch.empty((n, 256, 256, 3), dtype=torch.uint8, device="cuda")
_native.check(_native.lib().cv_synth_boards(_native.ptr(boards), 0, 0, n, 256, 1, 1, None, _native.stream_ptr(boards.device)))
if mask >= 0:
    model.set_impl(mask)
for _ in range(iters):
    fen, fen_len = model.predict_fen_device(boards)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    fen, fen_len = model.predict_fen_device(boards)
e1.record(); torch.cuda.synchronize()
if os.environ.get("CV_PROFILE_RESIZE", "1") != "0":          # the step before the path: one launch of the Pillow-exact resize kernel
    from chess_vision_b200.preprocess import resize_boards
    src = torch.randint(0, 256, (n, 400, 400, 3), dtype=torch.uint8, device="cuda")
    resize_boards(src, 256)
    torch.cuda.synchronize()
print(f"{n} boards: {e0.elapsed_time(e1) / iters:.3f} ms/iter, {n * iters / e0.elapsed_time(e1) * 1e3:.0f} boards/s", model.decode_fen_records(fen[:1], fen_len[:1])[0])
