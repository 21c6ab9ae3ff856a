"""Cycle counters of the instrumented build (CV_B200_LIB = a library built with -DCV_FE_PROFILE -DCV_SC_PROFILE -DCV_EXPERIMENTS):
one 4096-board forward with CV_FE3_DEBUG=256 CV_SC_DEBUG=256 prints what each role of the front end waits for and what each
section of a stage C tile costs (block 0 only)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chess_vision_b200 as cv
from chess_vision_b200 import _native, synthetic
m = cv.build_model({"model": {"arch": "square", "pretrained": False}})
m.load_state_dict(synthetic.init_state_dict(m.state_dict(), 0), strict=True); m = m.to("cuda").eval()
B = int(os.environ.get("BOARDS", 4096))
boards = torch.empty((B, 256, 256, 3), dtype=torch.uint8, device="cuda")
_native.check(_native.lib().cv_synth_boards(_native.ptr(boards), 0, 0, B, 256, 1, 1, None, _native.stream_ptr(boards.device)))
for _ in range(2):
    m.predict_fen_device(boards, precision=os.environ.get("PREC", "fp16"))
torch.cuda.synchronize()
