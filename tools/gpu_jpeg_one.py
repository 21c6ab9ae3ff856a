"""One cv_jpeg_decode_batch call of N files (for ncu captures of the JPEG kernels)."""
import io, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from PIL import Image
from chess_vision_b200 import preprocess, synthetic
n = int(os.environ.get("N", 4096))
base = synthetic.synth_boards(0, 64, 256, 1, synthetic.DIST_STRUCTURED)
files = []
for i in range(64):
    b = io.BytesIO(); Image.fromarray(base[i]).save(b, "JPEG", quality=90, subsampling=2); files.append(b.getvalue())
batch = [files[i % 64] for i in range(n)]
out = torch.empty((n, 256, 256, 3), dtype=torch.uint8, device="cuda")
for _ in range(int(os.environ.get("REPS", 2))):
    preprocess.decode_jpegs(batch, "cuda", out=out)
torch.cuda.synchronize()
print("ok", int(out[0].sum()))
