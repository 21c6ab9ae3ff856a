#!/usr/bin/env python
"""bf16-path accuracy statistics over many boards (run on the GPU box): RMS and max error of every output for each
kernel-selection mask, next to PyTorch's own bf16 execution of the oracle graph (the yard-stick).
    python tools/gpu_accuracy.py [n_boards] [H] [mask ...]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_vision_b200 as cv
from chess_vision_b200 import synthetic
from oracle import square_oracle as oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
masks = [int(a) for a in sys.argv[3:]] or [63, 31, 15]
gold = os.path.join(ROOT, "tests", "golden")
arrays = dict(np.load(os.path.join(gold, "reference_outputs.npz")))
meta = json.load(open(os.path.join(gold, "reference_meta.json")))
model = cv.build_model({"model": {"arch": "square", "pretrained": False}})
state = synthetic.init_state_dict(model.state_dict(), meta["weight_seed"])
state = synthetic.calibrate_heads(state, {k[4:]: arrays[k] for k in arrays if k.startswith("cal_")}, meta["cal_seed"])
model.load_state_dict(state); model = model.cuda().eval()
u8 = synthetic.synth_boards(2000, n, H, meta["board_seed"], synthetic.DIST_STRUCTURED)
x = oracle.normalize_u8(u8)
torch.set_num_threads(os.cpu_count())
ref = oracle.forward(x, state, return_features=True)
keys = ("features", "squares", "turn", "castling")


def stats(out):
    r = {}
    for k in keys:
        d = (out[k].double().cpu() - ref[k].double())
        r[k] = (float(d.pow(2).mean().sqrt() / ref[k].double().pow(2).mean().sqrt()), float(d.abs().max() / ref[k].abs().max()))
    return r


def show(name, r, extra=""):
    print(f"{name:28s} " + "  ".join(f"{k} rms {r[k][0]:.3e} max {r[k][1]:.3e}" for k in keys) + extra)


ref_cls = ref["squares"].reshape(-1, 13).argmax(-1)
if n <= 64:
    yard = oracle.forward(x, state, return_features=True, dtype=torch.bfloat16)
    show("PyTorch-bf16 yard-stick", stats(yard), f"  argmax agree {(yard['squares'].reshape(-1, 13).argmax(-1) == ref_cls).float().mean():.4f}")
bd = torch.from_numpy(u8).cuda()
for m in masks:
    model.set_impl(m)
    out = model.forward_u8(bd, precision="bf16", return_features=True)
    agree = (out["squares"].cpu().reshape(-1, 13).argmax(-1) == ref_cls).float().mean()
    show(f"kernels mask {m}", stats(out), f"  argmax agree {agree:.4f}")
model.set_impl(63)
out32 = model.forward_u8(bd, precision="fp32", return_features=True)
show("kernels fp32 mode", stats(out32))
