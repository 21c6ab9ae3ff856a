"""bf16 default path vs the fp32 oracle over board sizes (every multiple of 32 from 64 to 512) -- front-end table generality."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_vision_b200 as cv
from chess_vision_b200 import synthetic
from oracle import square_oracle as oracle
model = cv.build_model({"model": {"arch": "square", "pretrained": False}})
state = synthetic.init_state_dict(model.state_dict(), 0)
model.load_state_dict(state); model = model.cuda().eval()
for H in range(64, 513, 32):
    u8 = synthetic.synth_boards(0, 3, H, 1, synthetic.DIST_STRUCTURED)
    ref = oracle.forward(oracle.normalize_u8(u8), state, return_features=True)
    bd = torch.from_numpy(u8).cuda()
    out = model.forward_u8(bd, precision="bf16", return_features=True)
    o32 = model.forward_u8(bd, precision="fp32", return_features=True)
    e16 = float((out["features"].cpu() - ref["features"]).abs().max() / ref["features"].abs().max())
    e32 = float((o32["squares"].cpu() - ref["squares"]).abs().max() / ref["squares"].abs().max())
    fen_ok = model.predict_fen(bd, precision="fp32") == oracle.fen_strings(ref["squares"].numpy(), ref["turn"].numpy(), ref["castling"].numpy())
    print(f"H={H:3d}: bf16 features rel err {e16:.2e}, fp32 logits rel err {e32:.1e}, fp32 FEN equal {fen_ok}")
