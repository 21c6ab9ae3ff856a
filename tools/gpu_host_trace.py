import os, sys, time, torch
sys.path.insert(0, "/root/repo")
import chess_vision_b200 as cv, bench
from chess_vision_b200 import synthetic, _native
m = cv.build_model({"model": {"arch": "square", "pretrained": False}}); m.load_state_dict(bench.make_state(m.state_dict())); m = m.cuda().eval()
base = torch.from_numpy(synthetic.synth_boards(0, 512, 256, 1, synthetic.DIST_STRUCTURED))
host = base.repeat(8, 1, 1, 1).contiguous().pin_memory()
out = (torch.empty((4096, 80), dtype=torch.uint8).pin_memory(), torch.empty((4096,), dtype=torch.uint8).pin_memory())
for _ in range(3): m.predict_fen_host(host, out=out)
os.environ["CV_HOST_TRACE"] = "1"
t = time.perf_counter(); m.predict_fen_host(host, out=out); print("wall ms", (time.perf_counter() - t) * 1e3)
