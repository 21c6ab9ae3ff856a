#!/usr/bin/env python
"""ncu workload of the exact mode: one fp32_split forward of N boards between cudaProfilerStart / Stop (after a warm-up call).
    ncu --profile-from-start off --set full ... python tools/gpu_profile_split.py [boards=256]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_vision_b200 as cv
from chess_vision_b200 import _native
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model = cv.build_model({"model": {"arch": "square", "pretrained": False}})
model.load_state_dict(bench.make_state(model.state_dict()))
model = model.cuda().eval()
boards = torch.empty((n, 256, 256, 3), dtype=torch.uint8, device="cuda")
_native.check(_native.lib().cv_synth_boards(_native.ptr(boards), 0, 0, n, 256, 1, 1, None, _native.stream_ptr(boards.device)))
model.predict_fen_device(boards, precision="fp32_split")
torch.cuda.synchronize()
torch.cuda.profiler.start()
fen, fen_len = model.predict_fen_device(boards, precision="fp32_split")
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(model.decode_fen_records(fen[:1], fen_len[:1])[0])
