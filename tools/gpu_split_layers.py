"""fp32_split mode: per-layer time at 1024 boards against each layer's X2 traffic (4 bytes per element in + out + skip) and the measured HBM peak."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chess_vision_b200 as cv
from chess_vision_b200 import _native, synthetic, arch
m = cv.build_model({"model": {"arch": "square", "pretrained": False}})
m.load_state_dict(synthetic.init_state_dict(m.state_dict(), 0), strict=True); m = m.to("cuda").eval()
B = int(os.environ.get("BOARDS", "1024"))
boards = torch.empty((B, 256, 256, 3), dtype=torch.uint8, device="cuda")
_native.check(_native.lib().cv_synth_boards(_native.ptr(boards), 0, 0, B, 256, 1, 1, None, _native.stream_ptr(boards.device)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for wave in [int(w) for w in os.environ.get("WAVES", "512,1024").split(",")]:
    m.set_wave(wave)
    m.predict_fen_device(boards, precision="fp32_split")
    torch.cuda.synchronize(); e0.record()
    for _ in range(3): m.predict_fen_device(boards, precision="fp32_split")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    m.profile(True); m.predict_fen_device(boards, precision="fp32_split"); pm, pc = m.profile_read(); m.profile(False)
    print(f"wave {wave}: {ms:.2f} ms per {B} boards = {B / ms:.1f} k boards/s; profiled sum {pm.sum():.2f} ms")
    for i in np.argsort(-pm):
        if pm[i] <= 0: continue
        name = m.PROF_NAMES[i]
        extra = ""
        if 1 <= i <= 45:
            l = arch.LAYERS[i - 1]
            by = (l.in_elems + l.out_elems + (l.out_elems if l.skip >= 0 else 0)) * 4.0 * 64 * B
            if i == 1: by = (256 * 256 * 3 + l.out_elems * 4.0 * 64) * B
            extra = f"  {by / 1e9:6.2f} GB -> {by / (pm[i] * 1e-3) / 1e12:5.2f} TB/s ({by / (pm[i] * 1e-3) / 6525.2e9 * 100:4.1f} % of HBM peak)  launches {pc[i]}"
        print(f"   {pm[i]:7.3f} ms  {name}{extra}")
