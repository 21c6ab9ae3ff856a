"""Host-path schedules (experiment build, CV_HOST_SCHED / CV_HOST_CHUNK / CV_HOST_PIECE): ms per 4096-board step, best of several."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chess_vision_b200 as cv, bench
from chess_vision_b200 import synthetic
m = cv.build_model({"model": {"arch": "square", "pretrained": False}}); m.load_state_dict(bench.make_state(m.state_dict())); m = m.cuda().eval()
base = torch.from_numpy(synthetic.synth_boards(0, 512, 256, 1, synthetic.DIST_STRUCTURED))
host = base.repeat(8, 1, 1, 1).contiguous().pin_memory()
out = (torch.empty((4096, 80), dtype=torch.uint8).pin_memory(), torch.empty((4096,), dtype=torch.uint8).pin_memory())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ref = None
for sched, chunk, piece in ((0, 512, 128), (8, 512, 128), (16, 512, 128), (24, 512, 128), (32, 512, 128), (8, 512, 64)):
    os.environ["CV_HOST_SCHED"], os.environ["CV_HOST_CHUNK"], os.environ["CV_HOST_PIECE"] = str(sched), str(chunk), str(piece)
    for _ in range(3): m.predict_fen_host(host, out=out)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): m.predict_fen_host(host, out=out)
    e1.record(); torch.cuda.synchronize()
    fens = m.decode_fen_records(out[0][:64], out[1][:64])
    ref = ref or fens
    print(f"sched {sched} chunk {chunk} piece {piece}: {e0.elapsed_time(e1) / 10:.3f} ms per 4096 boards  same {fens == ref}", flush=True)
os.environ["CV_HOST_SCHED"], os.environ["CV_HOST_CHUNK"], os.environ["CV_HOST_PIECE"] = "8", "512", "128"
os.environ["CV_HOST_TRACE"] = "1"
m.predict_fen_host(host, out=out)
