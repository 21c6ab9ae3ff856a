"""Throughput of every precision / kernel-selection mode on device-resident boards (boards/s), for DESIGN.md."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_vision_b200 as cv
from chess_vision_b200 import _native, synthetic

m = cv.build_model({"model": {"arch": "square", "pretrained": False}})
m.load_state_dict(synthetic.init_state_dict(m.state_dict(), 0), strict=True)
m = m.to("cuda").eval()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def run(B, prec, mask=None, reps=3):
    boards = torch.empty((B, 256, 256, 3), dtype=torch.uint8, device="cuda")
    _native.check(_native.lib().cv_synth_boards(_native.ptr(boards), 0, 0, B, 256, 1, 1, None, _native.stream_ptr(boards.device)))
    if mask is not None: m.set_impl(mask)
    try:
        m.predict_fen_device(boards, precision=prec)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps): m.predict_fen_device(boards, precision=prec)
        e1.record(); torch.cuda.synchronize()
    finally:
        if mask is not None: m.set_impl(1023)
    ms = e0.elapsed_time(e1) / reps
    print(f"{prec:5s} mask {mask}: {B} boards {ms:8.2f} ms = {B / ms:8.1f} k boards/s", flush=True)
run(4096, "fp16"); run(4096, "bf16")
run(1024, "bf16", 15); run(1024, "bf16", 7); run(1024, "bf16", 0)
run(256, "fp32"); run(1024, "fp32")
