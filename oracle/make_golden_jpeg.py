"""Generate tests/golden/jpeg_reference.npz with Pillow itself (build container; Pillow 12.2.0 on libjpeg-turbo).

Small JPEG files written by Pillow with several qualities, chroma subsamplings, optimised Huffman tables, restart intervals, odd sizes
and a grayscale file, each stored with the RGB pixels ``PIL.Image.open(...).convert("RGB")`` decodes from it -- the call the reference
makes (predict.py:19).  tests/test_jpeg_cpu.py pins oracle/jpeg_oracle.py to them bit for bit; tests/test_jpeg_gpu.py the CUDA path.

    python oracle/make_golden_jpeg.py
"""
import io
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import resize_oracle  # noqa: E402  (its seeded synthetic photographs)

CASES = [  # (height, width, subsampling 0=4:4:4 1=4:2:2 2=4:2:0, quality, extra save options)
    (40, 56, 0, 90, {}), (40, 56, 1, 90, {}), (40, 56, 2, 90, {}),
    (17, 23, 2, 50, {}), (17, 23, 1, 75, {"optimize": True}), (17, 23, 0, 100, {}),
    (64, 64, 2, 90, {"restart_marker_blocks": 3}), (33, 70, 1, 85, {"restart_marker_rows": 1}), (33, 70, 2, 95, {"optimize": True}),
    (5, 3, 2, 90, {}), (8, 8, 1, 60, {}), (9, 35, 2, 30, {}),
    (100, 90, 2, 90, {}), (128, 128, 2, 90, {"restart_marker_blocks": 8}),
]


def main():
    out = {}
    names = []
    for i, (h, w, ss, q, kw) in enumerate(CASES):
        img = resize_oracle.synth_image(1000 + i, h, w)
        b = io.BytesIO()
        Image.fromarray(img).save(b, "JPEG", quality=q, subsampling=ss, **kw)
        data = b.getvalue()
        out[f"file{i}"] = np.frombuffer(data, np.uint8)
        out[f"rgb{i}"] = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
        names.append(f"{h}x{w} ss{ss} q{q} {kw}")
    g = resize_oracle.synth_image(7, 30, 41)[:, :, 0]
    b = io.BytesIO()
    Image.fromarray(g).save(b, "JPEG", quality=80)
    i = len(CASES)
    out[f"file{i}"] = np.frombuffer(b.getvalue(), np.uint8)
    out[f"rgb{i}"] = np.asarray(Image.open(io.BytesIO(b.getvalue())).convert("RGB"))
    names.append("30x41 grayscale q80")
    out["names"] = np.array(names)
    path = os.path.join(ROOT, "tests", "golden", "jpeg_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(names), "cases")


if __name__ == "__main__":
    main()
