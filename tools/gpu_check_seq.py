"""Replays bench.py's call sequence and checks parity at each step (bring-up diagnostics)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import chess_vision_b200 as cv
from chess_vision_b200 import synthetic, _native
from oracle import square_oracle as oracle
arrays = dict(np.load(os.path.join(ROOT, "tests/golden/reference_outputs.npz")))
model = cv.build_model({"model": {"arch": "square", "pretrained": False}})
state = synthetic.init_state_dict(model.state_dict(), 0)
state = synthetic.calibrate_heads(state, {k[4:]: arrays[k] for k in arrays if k.startswith("cal_")}, 999)
model.load_state_dict(state); model = model.to("cuda").eval()
B, n = 4096, 256
boards = torch.empty((B, 256, 256, 3), dtype=torch.uint8, device="cuda")
_native.check(_native.lib().cv_synth_boards(_native.ptr(boards), 0, 0, B, 256, 1, 1, None, _native.stream_ptr(boards.device)))
u8 = synthetic.synth_boards(0, n, 256, 1)
ref = oracle.forward(oracle.normalize_u8(u8), state, return_features=True)
f_ref = oracle.fen_strings(ref["squares"].numpy(), ref["turn"].numpy(), ref["castling"].numpy())

def report(tag, out):
    for k in ("features", "squares", "turn", "castling"):
        d = (out[k][: n * (64 if k == "features" else 1)].cpu() - ref[k]).abs()
        per = d.reshape(n, -1).max(1).values / ref[k].abs().max()
        print(f"{tag} {k}: max rel {float(per.max()):.3e} worst board {int(per.argmax())} median {float(per.median()):.3e}")

o_small = model.forward_u8(boards[:2].clone(), precision="bf16", return_features=True)
o_big = model.forward_u8(boards, precision="bf16", return_features=True)
report("bf16 B=4096", o_big)
print("bf16 board0 small-batch vs big-batch features equal:", bool(torch.equal(o_small["features"][:64], o_big["features"][:64])),
      float((o_small["features"][:64] - o_big["features"][:64]).abs().max()))
for mask in (0, 7):
    model.set_impl(mask)
    o = model.forward_u8(boards[:n].clone(), precision="bf16", return_features=True)
    report(f"bf16 mask{mask} B={n}", o)
model.set_impl(15)
o = model.forward_u8(boards[:n].clone(), precision="bf16", return_features=True)
report(f"bf16 mask15 B={n}", o)
f16 = model.predict_fen(boards[:n].clone(), precision="bf16")
print("bf16 fen agreement vs fp32 oracle:", np.mean([a == b for a, b in zip(f_ref, f16)]))
host = boards.cpu().pin_memory()
fh = model.predict_fen(host, precision="bf16")
print("host pipeline == device:", fh[:n] == f16)
f32 = model.predict_fen(boards[:64], precision="fp32")
print("fp32 after sequence agreement:", np.mean([a == b for a, b in zip(f_ref[:64], f32)]))
o32 = model.forward_u8(boards[:n].clone(), precision="fp32", return_features=True)
report("fp32", o32)
