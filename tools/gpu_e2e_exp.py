"""End-to-end host path (pinned boards -> FEN records on the host) per batch size: where the time beyond the H2D copy goes."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chess_vision_b200 as cv, bench
from chess_vision_b200 import synthetic
m = cv.build_model({"model": {"arch": "square", "pretrained": False}}); m.load_state_dict(bench.make_state(m.state_dict())); m = m.cuda().eval()
base = torch.from_numpy(synthetic.synth_boards(0, 512, 256, 1, synthetic.DIST_STRUCTURED))
for B in (128, 512, 1024, 2048, 4096, 8192):
    host = base.repeat((B + 511) // 512, 1, 1, 1)[:B].contiguous().pin_memory()
    for _ in range(3): m.predict_fen_host(host)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): m.predict_fen_host(host)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    copy_ms = host.numel() / 55.2e9 * 1e3
    print(f"B={B:5d}: {dt*1e3:7.2f} ms/step  {B/dt:8.0f} boards/s   (H2D alone {copy_ms:6.2f} ms, rest {dt*1e3-copy_ms:5.2f} ms)")
