#!/usr/bin/env python
"""Layer-granular bf16 path (cv_square_set_impl mask 15: one tensor-core / vectorised kernel per layer, activations through HBM):
per-layer CUDA-event time, algorithmic bytes (input + output elements x 2 B per crop) and the fraction of the measured HBM peak.
    python tools/gpu_layer_roofline.py [boards=1024] [mask=15] [wave=128]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_vision_b200 as cv
from chess_vision_b200 import _native, arch
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mask = int(sys.argv[2]) if len(sys.argv) > 2 else 15
wave = int(sys.argv[3]) if len(sys.argv) > 3 else 128
peak = bench.measured_peaks()["hbm_gbs"]
model = cv.build_model({"model": {"arch": "square", "pretrained": False}})
model.load_state_dict(bench.make_state(model.state_dict()))
model = model.cuda().eval()
model.set_impl(mask)
model.set_wave(wave)
boards = torch.empty((n, 256, 256, 3), dtype=torch.uint8, device="cuda")
_native.check(_native.lib().cv_synth_boards(_native.ptr(boards), 0, 0, n, 256, 1, 1, None, _native.stream_ptr(boards.device)))
for _ in range(2): model.predict_fen_device(boards)
model.profile(True)
iters = 3
for _ in range(iters): model.predict_fen_device(boards)
ms, cnt = model.profile_read()
model.profile(False)
crops = 64 * n
kinds = {0: "dense3x3", 1: "pointwise", 2: "depthwise"}
tot = {k: [0.0, 0.0] for k in kinds.values()}
print(f"{n} boards, wave {wave}, mask {mask}: {ms.sum() / iters:.2f} ms per call; HBM peak {peak:.0f} GB/s (measured)")
print("| layer | kind | in+out el/crop | ms | GB/s | of HBM peak |\n|---|---|---|---|---|---|")
for l in arch.LAYERS:
    t = ms[1 + l.index] / iters
    if t <= 0: continue
    el = l.cin * l.hin * l.hin + l.cout * l.hout * l.hout
    if l.skip >= 0: el += l.cout * l.hout * l.hout
    gbs = el * 2 * crops / (t / 1e3) / 1e9
    tot[kinds[l.kind]][0] += el * 2 * crops; tot[kinds[l.kind]][1] += t
    print(f"| L{l.index} {l.key} | {kinds[l.kind]} | {el} | {t:.3f} | {gbs:.0f} | {100 * gbs / peak:.0f} % |")
for k, (b, t) in tot.items():
    if t > 0: print(f"| all {k} | | | {t:.3f} | {b / (t / 1e3) / 1e9:.0f} | {100 * b / (t / 1e3) / 1e9 / peak:.0f} % |")
