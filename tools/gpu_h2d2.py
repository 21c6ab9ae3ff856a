"""H2D bandwidth from pinned memory alone and while the fused kernels run on another stream."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chess_vision_b200 as cv, bench
from chess_vision_b200 import synthetic
n = 4096
host = torch.empty((n, 256, 256, 3), dtype=torch.uint8).pin_memory()
dev = torch.empty_like(host, device="cuda")
cs = torch.cuda.Stream()
def copy_all():
    with torch.cuda.stream(cs):
        for i in range(0, n, 512):
            dev[i:i + 512].copy_(host[i:i + 512], non_blocking=True)
for _ in range(2): copy_all()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5): copy_all()
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print(f"H2D alone: {dt*1e3:.2f} ms per 805 MB = {host.numel()/dt/1e9:.1f} GB/s")
m = cv.build_model({"model": {"arch": "square", "pretrained": False}}); m.load_state_dict(bench.make_state(m.state_dict())); m = m.cuda().eval()
boards = torch.from_numpy(synthetic.synth_boards(0, 512, 256, 1, synthetic.DIST_STRUCTURED)).cuda()
for _ in range(3): m.predict_fen_device(boards)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    copy_all()
    for _ in range(8): m.predict_fen_device(boards)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print(f"H2D of 805 MB + 8 x 512-board compute concurrently: {dt*1e3:.2f} ms per step")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    for _ in range(8): m.predict_fen_device(boards)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print(f"8 x 512-board compute alone: {dt*1e3:.2f} ms per step")
