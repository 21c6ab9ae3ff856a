"""fp16-operand mode against the CPU oracle, beside the bf16 and fp32 modes: logit errors, FEN agreement, overflow fall-back, stage times."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_vision_b200 as cv
from chess_vision_b200 import _native, synthetic
from oracle import square_oracle as oracle
sys.path.insert(0, os.path.join(ROOT, "tests"))

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

def gold():
    arrays = dict(np.load(os.path.join(ROOT, "tests/golden/reference_outputs.npz")))
    meta = json.load(open(os.path.join(ROOT, "tests/golden/reference_meta.json")))
    template = {k: torch.zeros(meta["shapes"][k], dtype=torch.long if ("num_batches" in k or k.startswith("class_to")) else torch.float32) for k in meta["keys"]}
    template["class_to_type"] = torch.tensor(meta["class_tables"]["type"]); template["class_to_color"] = torch.tensor(meta["class_tables"]["color"])
    state = synthetic.init_state_dict(template, meta["weight_seed"])
    stats = {k[4:]: arrays[k] for k in arrays if k.startswith("cal_")}
    return synthetic.calibrate_heads(state, stats, meta["cal_seed"]), meta

state, meta = gold()
cfg = {"model": {"arch": "square", "pretrained": False}}
m = cv.build_model(cfg); m.load_state_dict(state, strict=True); m = m.to("cuda").eval()
n = int(os.environ.get("N", "64"))
for H in (256, 512):
    nn_ = n if H == 256 else 8
    u8 = synthetic.synth_boards(0, nn_, H, meta["board_seed"], synthetic.DIST_STRUCTURED)
    ref = oracle.forward(oracle.normalize_u8(u8), state, return_features=True)
    ref_fen = oracle.fen_strings(ref["squares"].numpy(), ref["turn"].numpy(), ref["castling"].numpy())
    bd = torch.from_numpy(u8).cuda()
    for prec in ("fp16", "bf16", "fp32"):
        out = m.forward_u8(bd, precision=prec, return_features=True)
        errs = {k: rel(out[k].cpu().numpy(), ref[k].numpy()) for k in ("features", "squares", "turn", "castling")}
        fen = m.predict_fen(bd, precision=prec)
        sq_ref = ref["squares"].numpy().reshape(-1, 13).argmax(-1); sq_got = out["squares"].cpu().numpy().reshape(-1, 13).argmax(-1)
        print(f"H={H} {prec}: " + " ".join(f"{k} {v:.2e}" for k, v in errs.items()) +
              f" | square agreement {np.mean(sq_ref == sq_got):.4f} FEN agreement {np.mean([a == b for a, b in zip(fen, ref_fen)]):.3f}", flush=True)
    print("fp16 status (weights_fit, overflowed):", m.fp16_status())
    # the float entry point must equal the uint8 entry point in fp16 mode
    x = oracle.normalize_u8(u8).cuda()
    a, b = m(x, precision="fp16"), m.forward_u8(bd, precision="fp16")
    print("float entry == u8 entry (fp16):", all(torch.equal(a[k], b[k]) for k in ("squares", "turn", "castling")))
    chw = torch.from_numpy(np.ascontiguousarray(u8.transpose(0, 3, 1, 2))).cuda()
    c = m.forward_u8(chw, layout="chw", precision="fp16", return_features=True)
    print("chw (front end v1, fp16 out) features err", rel(c["features"].cpu().numpy(), ref["features"].numpy()))

# ---- overflow fall-back: blow one BatchNorm scale up so activations leave the fp16 range
st2 = {k: v.clone() for k, v in state.items()}
st2["backbone.blocks.2.1.pw_exp.bn.weight"] *= 3000.0
m2 = cv.build_model(cfg); m2.load_state_dict(st2, strict=True); m2 = m2.to("cuda").eval()
u8 = synthetic.synth_boards(0, 40, 256, 1, synthetic.DIST_STRUCTURED)
bd = torch.from_numpy(u8).cuda()
o16 = m2.forward_u8(bd, precision="fp16", return_features=True)
stt = m2.fp16_status()
ob = m2.forward_u8(bd, precision="bf16", return_features=True)
print("overflow case: status", stt, "fp16-mode == bf16-mode bit for bit:", all(torch.equal(o16[k], ob[k]) for k in o16),
      "finite:", bool(torch.isfinite(o16["squares"]).all()))
o16b = m.forward_u8(bd, precision="fp16"); print("normal weights after that: status", m.fp16_status())

# ---- stage times, 4096 boards
B = 4096
boards = torch.empty((B, 256, 256, 3), dtype=torch.uint8, device="cuda")
_native.check(_native.lib().cv_synth_boards(_native.ptr(boards), 0, 0, B, 256, 1, 1, None, _native.stream_ptr(boards.device)))
for prec in ("fp16", "bf16"):
    for _ in range(3): m.predict_fen_device(boards, precision=prec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): m.predict_fen_device(boards, precision=prec)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    m.profile(True)
    for _ in range(5): m.predict_fen_device(boards, precision=prec)
    pm, pc = m.profile_read(); m.profile(False)
    top = np.argsort(-pm)[:6]
    print(f"{prec}: {ms:.3f} ms / {B} boards = {B / ms:.1f} k boards/s | " + ", ".join(f"{m.PROF_NAMES[i]} {pm[i] / 5:.3f}" for i in top), flush=True)
for nb in (1, 64):
    b1 = boards[:nb].clone()
    for prec in ("fp16", "bf16"):
        for _ in range(5): m.predict_fen_device(b1, precision=prec)
        torch.cuda.synchronize(); e0.record()
        for _ in range(50): m.predict_fen_device(b1, precision=prec)
        e1.record(); torch.cuda.synchronize()
        print(f"{nb} boards {prec}: {e0.elapsed_time(e1) / 50 * 1000:.1f} us per call")
