"""Generate tests/golden/resize_reference.npz with Pillow ITSELF (build container only): the reference's eval transform
(/root/reference/dataset.py:177-181) is `torchvision.transforms.Resize((S, S))` on a PIL image, i.e. `Image.resize((S, S), BILINEAR)`.

Each case is a seeded image (oracle/resize_oracle.synth_image) resized by torchvision's own `transforms.Resize` (checked equal to
`Image.resize`).  Small cases are stored whole, large ones as CRC32 of the output bytes.

    python oracle/make_golden_resize.py
"""
import os
import sys
import zlib

import numpy as np
import PIL
from PIL import Image
from torchvision import transforms

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import resize_oracle  # noqa: E402

# (seed, in_h, in_w, out_h, out_w, store whole output)
CASES = [(1, 40, 36, 32, 32, True), (2, 24, 24, 32, 32, True), (3, 33, 50, 32, 32, True), (4, 32, 48, 32, 32, True), (5, 50, 32, 32, 32, True),
         (6, 400, 400, 256, 256, False), (7, 512, 512, 256, 256, False), (8, 300, 280, 256, 256, False), (9, 200, 200, 256, 256, False),
         (10, 1024, 1024, 256, 256, False), (11, 256, 256, 256, 256, False), (12, 800, 800, 512, 512, False), (13, 257, 255, 256, 256, False)]


def main():
    out = {"pillow_version": np.array(PIL.__version__)}
    meta = []
    for seed, h, w, oh, ow, whole in CASES:
        img = resize_oracle.synth_image(seed, h, w)
        ref = np.asarray(transforms.Resize((oh, ow))(Image.fromarray(img)))
        direct = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
        assert np.array_equal(ref, direct)
        meta.append((seed, h, w, oh, ow, int(whole), zlib.crc32(ref.tobytes())))
        if whole:
            out[f"out_{seed}"] = ref
    out["cases"] = np.array(meta, dtype=np.int64)
    path = os.path.join(ROOT, "tests", "golden", "resize_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; Pillow", PIL.__version__)


if __name__ == "__main__":
    main()
