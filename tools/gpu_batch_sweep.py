"""Batch sweep 1..65536 (BASELINE.json configs[3]): latency and throughput of predict_fen_device (boards resident in HBM, FEN
records left on the device), bf16, flipped flags on."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chess_vision_b200 as cv, bench
from chess_vision_b200 import _native
m = cv.build_model({"model": {"arch": "square", "pretrained": False}}); m.load_state_dict(bench.make_state(m.state_dict())); m = m.cuda().eval()
nmax = 65536
boards = torch.empty((nmax, 256, 256, 3), dtype=torch.uint8, device="cuda")
flipped = torch.empty((nmax,), dtype=torch.uint8, device="cuda")
for i in range(0, nmax, 8192):
    _native.check(_native.lib().cv_synth_boards(_native.ptr(boards[i:]), 0, i, 8192, 256, 1, 1, _native.ptr(flipped[i:]), _native.stream_ptr(boards.device)))
torch.cuda.synchronize()
print("| boards | ms per call | boards/s |\n|---|---|---|")
B = 1
while B <= nmax:
    reps = max(3, min(200, 4096 // B))
    for _ in range(3): m.predict_fen_device(boards[:B], flipped=flipped[:B])
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): m.predict_fen_device(boards[:B], flipped=flipped[:B])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"| {B} | {ms:.3f} | {B / ms * 1e3:,.0f} |")
    B *= 4
