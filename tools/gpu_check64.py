import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import chess_vision_b200 as cv
from chess_vision_b200 import synthetic
from oracle import square_oracle as oracle
arrays = dict(np.load(os.path.join(ROOT, "tests/golden/reference_outputs.npz")))
model = cv.build_model({"model": {"arch": "square", "pretrained": False}})
state = synthetic.init_state_dict(model.state_dict(), 0)
state = synthetic.calibrate_heads(state, {k[4:]: arrays[k] for k in arrays if k.startswith("cal_")}, 999)
model.load_state_dict(state); model = model.to("cuda").eval()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
u8 = synthetic.synth_boards(0, n, 256, 1)
ref = oracle.forward(oracle.normalize_u8(u8), state, return_features=True)
out = model.forward_u8(torch.from_numpy(u8).cuda(), precision="fp32", return_features=True)
for k in ("squares", "turn", "castling", "features"):
    d = (out[k].cpu() - ref[k]).abs().reshape(n, -1).max(1).values
    print(k, "per-board max abs err: max", float(d.max()), "argmax", int(d.argmax()), [f"{v:.0e}" for v in d.tolist()][:70])
f_ref = oracle.fen_strings(ref["squares"].numpy(), ref["turn"].numpy(), ref["castling"].numpy())
f_gpu = model.predict_fen(torch.from_numpy(u8).cuda(), precision="fp32")
print("fen agree:", [a == b for a, b in zip(f_ref, f_gpu)])
for a, b in list(zip(f_ref, f_gpu))[:12]:
    if a != b: print(a, "|", b)

from chess_vision_b200 import _native
dev = torch.empty((n, 256, 256, 3), dtype=torch.uint8, device="cuda")
_native.check(_native.lib().cv_synth_boards(_native.ptr(dev), 0, 0, n, 256, 1, 1, None, _native.stream_ptr(dev.device)))
print("device boards equal numpy:", bool((dev.cpu().numpy() == u8).all()))
f_dev = model.predict_fen(dev, precision="fp32")
print("fen agree (device boards):", sum(a == b for a, b in zip(f_ref, f_dev)), "of", n)
