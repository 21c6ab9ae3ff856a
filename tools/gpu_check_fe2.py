import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_vision_b200 as cv
from chess_vision_b200 import synthetic
from oracle import square_oracle as oracle
gold = os.path.join(ROOT, "tests", "golden")
arrays = dict(np.load(os.path.join(gold, "reference_outputs.npz")))
meta = json.load(open(os.path.join(gold, "reference_meta.json")))
model = cv.build_model({"model": {"arch": "square", "pretrained": False}})
state = synthetic.init_state_dict(model.state_dict(), meta["weight_seed"])
state = synthetic.calibrate_heads(state, {k[4:]: arrays[k] for k in arrays if k.startswith("cal_")}, meta["cal_seed"])
model.load_state_dict(state); model = model.cuda().eval()
for H, n in ((512, 3), (512, 40), (256, 40)):
    u8 = synthetic.synth_boards(0, n, H, 1, synthetic.DIST_STRUCTURED)
    ref = oracle.forward(oracle.normalize_u8(u8), state, return_features=True)["features"]
    bd = torch.from_numpy(u8).cuda()
    outs = []
    for rep in range(4):
        outs.append(model.forward_u8(bd, precision="bf16", return_features=True)["features"].cpu())
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    model.set_impl(255); v1 = model.forward_u8(bd, precision="bf16", return_features=True)["features"].cpu(); model.set_impl(1023)
    def rms(a): return float(((a - ref).double().pow(2).mean().sqrt()) / ref.double().pow(2).mean().sqrt())
    def mx(a): return float((a - ref).abs().max() / ref.abs().max())
    print(f"H={H} n={n}: deterministic={same}  v2 rms {rms(outs[0]):.3e} max {mx(outs[0]):.3e} | v1 rms {rms(v1):.3e} max {mx(v1):.3e}")
