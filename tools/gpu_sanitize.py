"""Smallest end-to-end case for compute-sanitizer (one tool per gpurun call): 2 boards through predict_fen in the fp16 default mode
(both passes are enqueued: the fp16 kernels run, the gated bf16 kernels exit), the bf16 mode and the host-buffer entry point."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_vision_b200 as cv
from chess_vision_b200 import synthetic

m = cv.build_model({"model": {"arch": "square", "pretrained": False}})
m.load_state_dict(synthetic.init_state_dict(m.state_dict(), 0), strict=True)
m = m.to("cuda").eval()
u8 = torch.from_numpy(synthetic.synth_boards(0, 2, 256, 1, synthetic.DIST_STRUCTURED))
a = m.predict_fen(u8.cuda(), precision="fp16")
b = m.predict_fen(u8.cuda(), precision="bf16")
c = m.predict_fen(u8.pin_memory(), precision="fp16")
if "--fp32" in sys.argv:
    m.predict_fen(u8.cuda(), precision="fp32")
torch.cuda.synchronize()
assert a == c
print("ok", a[0], "|", b[0])
