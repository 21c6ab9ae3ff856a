"""Per-stage times (ms per 4096 boards) of the library given in CV_B200_LIB -- for A/B runs of experiment builds."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chess_vision_b200 as cv
from chess_vision_b200 import _native, synthetic
m = cv.build_model({"model": {"arch": "square", "pretrained": False}})
m.load_state_dict(synthetic.init_state_dict(m.state_dict(), 0), strict=True); m = m.to("cuda").eval()
B = 4096
boards = torch.empty((B, 256, 256, 3), dtype=torch.uint8, device="cuda")
_native.check(_native.lib().cv_synth_boards(_native.ptr(boards), 0, 0, B, 256, 1, 1, None, _native.stream_ptr(boards.device)))
prec = os.environ.get("PREC", "fp16")
for _ in range(3): fen, _ = m.predict_fen_device(boards, precision=prec)
ref = fen.clone()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10): m.predict_fen_device(boards, precision=prec)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
m.profile(True)
for _ in range(10): m.predict_fen_device(boards, precision=prec)
pm, pc = m.profile_read(); m.profile(False)
top = np.argsort(-pm)[:6]
import zlib
print(f"{os.path.basename(os.environ.get('CV_B200_LIB', 'default'))} {prec}: {ms:.3f} ms | " + ", ".join(f"{m.PROF_NAMES[i].split('(')[0]} {pm[i] / 10:.3f}" for i in top) + f" | crc {zlib.crc32(ref.cpu().numpy().tobytes()):08x}", flush=True)
