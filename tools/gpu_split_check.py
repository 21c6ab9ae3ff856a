"""fp32_split mode (split fp16 operands on the tensor cores) against the CPU oracle: per-layer taps, logits, FEN strings, throughput."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_vision_b200 as cv
from chess_vision_b200 import _native, arch, synthetic
from oracle import square_oracle as oracle

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

arrays = dict(np.load(os.path.join(ROOT, "tests/golden/reference_outputs.npz")))
meta = json.load(open(os.path.join(ROOT, "tests/golden/reference_meta.json")))
m = cv.build_model({"model": {"arch": "square", "pretrained": False}})
state = synthetic.init_state_dict(m.state_dict(), meta["weight_seed"])
state = synthetic.calibrate_heads(state, {k[4:]: arrays[k] for k in arrays if k.startswith("cal_")}, meta["cal_seed"])
m.load_state_dict(state, strict=True); m = m.to("cuda").eval()
u8 = synthetic.synth_boards(0, 2, 256, meta["board_seed"], synthetic.DIST_STRUCTURED)
x = oracle.normalize_u8(u8)
taps = {}
oracle.forward(x, state, taps=taps)
xd = x.cuda()
worst = 0
for l in arch.LAYERS:
    got = m.tap_layer(xd, l.index, precision="fp32_split").cpu().numpy()
    ref = taps[l.key].permute(0, 2, 3, 1).numpy()
    e = rel(got, ref); worst = max(worst, e)
    print(f"L{l.index}:{e:.1e}", end=" ", flush=True)
print("\nworst per-layer", worst)
for H, n in ((256, 64), (512, 8)):
    u8 = synthetic.synth_boards(100, n, H, 1, synthetic.DIST_STRUCTURED)
    ref = oracle.forward(oracle.normalize_u8(u8), state, return_features=True)
    truth = oracle.forward(oracle.normalize_u8(u8), state, return_features=True, dtype=torch.float64)
    ref_fen = oracle.fen_strings(ref["squares"].numpy(), ref["turn"].numpy(), ref["castling"].numpy())
    bd = torch.from_numpy(u8).cuda()
    for prec in ("fp32_split", "fp32"):
        out = m.forward_u8(bd, precision=prec, return_features=True)
        fen = m.predict_fen(bd, precision=prec)
        print(f"H={H} {prec}: vs fp32 oracle " + " ".join(f"{k} {rel(out[k].cpu().numpy(), ref[k].numpy()):.2e}" for k in ("features", "squares", "turn", "castling")) +
              " | vs fp64 truth " + " ".join(f"{k} {rel(out[k].cpu().numpy(), truth[k].numpy()):.2e}" for k in ("squares", "turn", "castling")) +
              f" | FEN agreement {np.mean([a == b for a, b in zip(fen, ref_fen)]):.3f} status {m.fp16_status()}", flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
B = 1024
boards = torch.empty((B, 256, 256, 3), dtype=torch.uint8, device="cuda")
_native.check(_native.lib().cv_synth_boards(_native.ptr(boards), 0, 0, B, 256, 1, 1, None, _native.stream_ptr(boards.device)))
for wave in (32, 64, 128, 256):
    m.set_wave(wave)
    m.predict_fen_device(boards, precision="fp32_split")
    torch.cuda.synchronize(); e0.record()
    for _ in range(2): m.predict_fen_device(boards, precision="fp32_split")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    print(f"fp32_split wave {wave}: {ms:.2f} ms per {B} boards = {B / ms:.1f} k boards/s", flush=True)
m.set_wave(0)
m.profile(True)
m.predict_fen_device(boards, precision="fp32_split")
pm, pc = m.profile_read(); m.profile(False)
top = np.argsort(-pm)[:12]
print("top kernels (ms per 1024 boards): " + ", ".join(f"{m.PROF_NAMES[i]} {pm[i]:.2f}" for i in top))
