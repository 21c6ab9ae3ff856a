"""fp32_split mode throughput (1024 boards) and per-kernel split; experiment builds: CV_SPLIT_SUB = boards per front sub-wave."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chess_vision_b200 as cv
from chess_vision_b200 import _native, synthetic
m = cv.build_model({"model": {"arch": "square", "pretrained": False}})
m.load_state_dict(synthetic.init_state_dict(m.state_dict(), 0), strict=True); m = m.to("cuda").eval()
B = 1024
boards = torch.empty((B, 256, 256, 3), dtype=torch.uint8, device="cuda")
_native.check(_native.lib().cv_synth_boards(_native.ptr(boards), 0, 0, B, 256, 1, 1, None, _native.stream_ptr(boards.device)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ref = None
for wave in [int(w) for w in os.environ.get("WAVES", "128").split(",")]:
    m.set_wave(wave)
    fen, _ = m.predict_fen_device(boards, precision="fp32_split")
    torch.cuda.synchronize(); e0.record()
    for _ in range(3): m.predict_fen_device(boards, precision="fp32_split")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    ref = fen.clone() if ref is None else ref
    m.profile(True); m.predict_fen_device(boards, precision="fp32_split"); pm, pc = m.profile_read(); m.profile(False)
    top = np.argsort(-pm)[:7]
    print(f"sub {os.environ.get('CV_SPLIT_SUB', 'default')} wave {wave}: {ms:.2f} ms = {B / ms:.1f} k boards/s same {bool(torch.equal(fen, ref))} | " +
          ", ".join(f"{m.PROF_NAMES[i].split(':')[0]} {pm[i]:.2f}" for i in top), flush=True)
