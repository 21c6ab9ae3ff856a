"""CPU restatement of the board resize in the reference's input transform (TEST INFRASTRUCTURE -- only tests/, smoke() and
bench.py's CPU legs may import this; the product path is chess_vision_b200/csrc/resize.cu).

SURVEY section 8f, N1: the step right before the hot path is `transforms.Resize((S, S))` on a PIL image
(/root/reference/dataset.py:177-181, fed at /root/reference/predict.py:19-20).  torchvision hands a PIL image to
`Image.resize((S, S), BILINEAR)`; the arithmetic lives in Pillow (third-party, unpinned in the reference's
requirements.txt; this container has Pillow 12.2.0), file src/libImaging/Resample.c, restated here from its published algorithm:

  * `precompute_coeffs`: scale = in/out, filterscale = max(scale, 1), support = 1.0 * filterscale (bilinear "triangle" filter with
    antialiasing when shrinking); per output index: center = (xx + 0.5) * scale, xmin = int(center - support + 0.5) clipped at 0,
    xmax = int(center + support + 0.5) clipped at `in`, weights triangle((x + xmin - center + 0.5) / filterscale) normalised to sum 1
    in double precision;
  * `normalize_coeffs_8bpc`: weights -> int32 fixed point with 22 fractional bits, rounded half away from zero;
  * `ImagingResampleHorizontal_8bpc` then `ImagingResampleVertical_8bpc`: int32 accumulation starting at 1 << 21, `>> 22`, clipped to
    0..255; the horizontal result is rounded to uint8 BEFORE the vertical pass; a pass whose size does not change is skipped, and
    `Image.resize` returns a copy when neither changes.

Pinned: tests/golden/resize_reference.npz holds outputs of Pillow itself (oracle/make_golden_resize.py, run in the build container)
for seeded images of several sizes; tests/test_resize_cpu.py checks this restatement against them bit for bit.
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bilinear_coeffs(in_size: int, out_size: int):
    """-> (ksize, bounds int32 (out, 2) = (first tap, tap count), coefficients int32 (out, ksize)) as Pillow computes them."""
    scale = float(in_size) / float(out_size)
    filterscale = scale if scale > 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)           # C cast: truncation (the value is positive whenever it matters)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            w[x] = 1.0 - a if a < 1.0 else 0.0
            ww += w[x]
        if ww != 0.0:
            for x in range(xmax):
                w[x] /= ww
        for x in range(ksize):
            v = w[x] * float(1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _resample_axis0(img: np.ndarray, out_size: int) -> np.ndarray:
    """One Pillow pass along axis 0 of a uint8 array (n, ...): int32 fixed point, rounded and clipped to uint8."""
    _, bounds, kk = bilinear_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], dtype=np.uint8)
    src = img.astype(np.int64)
    for i in range(out_size):
        x0, n = int(bounds[i, 0]), int(bounds[i, 1])
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for k in range(n):
            acc += src[x0 + k] * int(kk[i, k])
        out[i] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """(H, W, C) uint8 -> (out_h, out_w, C) uint8, bit-exact with `PIL.Image.resize((out_w, out_h), BILINEAR)`."""
    assert img.dtype == np.uint8 and img.ndim == 3
    out = img
    if out.shape[1] != out_w:                                   # horizontal pass first (Resample.c: ImagingResampleInner)
        out = _resample_axis0(np.ascontiguousarray(out.transpose(1, 0, 2)), out_w).transpose(1, 0, 2)
    if out.shape[0] != out_h:
        out = _resample_axis0(np.ascontiguousarray(out), out_h)
    return np.ascontiguousarray(out).copy()


def synth_image(seed: int, h: int, w: int) -> np.ndarray:
    """Seeded test image: board-like blocks + noise + a gradient (so that rounding cases of every kind occur)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    blocks = (((yy * 8) // max(h, 1) + (xx * 8) // max(w, 1)) % 2) * 150
    base = blocks[..., None] + rng.integers(0, 106, size=(h, w, 3))
    base[: h // 4] = rng.integers(0, 256, size=(h // 4, w, 3))   # a band of pure noise
    base[-(h // 8 + 1):, :, 0] = (xx[-(h // 8 + 1):] * 255) // max(w - 1, 1)
    return np.clip(base, 0, 255).astype(np.uint8)
