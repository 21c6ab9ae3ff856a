"""JPEG input path: decode rate of cv_jpeg_decode_batch per batch size (experiment build + CV_JPEG_TRACE=1 prints the stage split)."""
import io, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from PIL import Image
from chess_vision_b200 import preprocess, synthetic
base = synthetic.synth_boards(0, 64, 256, 1, synthetic.DIST_STRUCTURED)
files = []
for i in range(64):
    b = io.BytesIO(); Image.fromarray(base[i]).save(b, "JPEG", quality=90, subsampling=2); files.append(b.getvalue())
print("mean file bytes", np.mean([len(f) for f in files]))
for n in (1, 8, 32, 256, 1024, 4096, 16384):
    batch = [files[i % 64] for i in range(n)]
    out = torch.empty((n, 256, 256, 3), dtype=torch.uint8, device="cuda")
    for host in (False, True):
        if host and n > 256:
            continue
        preprocess.decode_jpegs(batch, "cuda", entropy_on_host=host, out=out)
        reps = 5 if n <= 256 else 2
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(reps):
            preprocess.decode_jpegs(batch, "cuda", entropy_on_host=host, out=out)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / reps
        print(f"n={n} entropy_on_host={host}: {dt * 1e3:.2f} ms = {n / dt / 1e3:.1f} k files/s", flush=True)
