"""Board resize of the input transform, CPU side: the oracle against Pillow's own outputs (golden vectors made by
oracle/make_golden_resize.py with Pillow 12.2.0 through torchvision's transforms.Resize), and the host-side coefficient tables of the
C-ABI library (cv_resize_coeffs_host, no GPU) against the oracle's."""
import os
import zlib

import numpy as np
import pytest

from oracle import resize_oracle as ro

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "resize_reference.npz"))
CASES = [tuple(int(v) for v in c) for c in GOLD["cases"]]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[1]}x{c[2]}to{c[3]}x{c[4]}")
def test_oracle_is_bit_exact_with_pillow(case):
    seed, h, w, oh, ow, whole, crc = case
    out = ro.resize_bilinear_u8(ro.synth_image(seed, h, w), oh, ow)
    assert out.shape == (oh, ow, 3) and out.dtype == np.uint8
    assert zlib.crc32(out.tobytes()) == crc
    if whole:
        assert np.array_equal(out, GOLD[f"out_{seed}"])


def test_oracle_against_installed_pillow_when_present():
    """Where Pillow is importable (it is in this image) the oracle is also checked live on sizes outside the golden set."""
    Image = pytest.importorskip("PIL.Image")
    for seed, (h, w, oh, ow) in enumerate([(100, 77, 64, 64), (64, 64, 96, 80), (31, 200, 17, 23), (400, 400, 224, 224)]):
        img = ro.synth_image(100 + seed, h, w)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
        assert np.array_equal(ro.resize_bilinear_u8(img, oh, ow), ref)


def test_identity_and_constant_images():
    img = ro.synth_image(5, 48, 48)
    assert np.array_equal(ro.resize_bilinear_u8(img, 48, 48), img)                 # Image.resize copies when the size is unchanged
    flat = np.full((40, 52, 3), 201, dtype=np.uint8)
    assert np.all(ro.resize_bilinear_u8(flat, 32, 32) == 201)                      # weights sum to one: constants are preserved
    assert np.all(ro.resize_bilinear_u8(flat, 64, 80) == 201)


@pytest.mark.parametrize("sizes", [(400, 256), (512, 256), (200, 256), (1024, 256), (257, 256), (255, 256), (36, 32), (256, 256),
                                   (800, 512), (1, 7), (7, 1), (3000, 224)])
def test_library_tables_match_oracle(sizes):
    from chess_vision_b200 import preprocess
    ks, bounds, kk = preprocess.resize_coeffs(*sizes)
    ks2, bounds2, kk2 = ro.bilinear_coeffs(*sizes)
    assert ks == ks2 and np.array_equal(bounds, bounds2) and np.array_equal(kk, kk2)
    assert np.all(np.abs(kk.sum(axis=1) - (1 << 22)) <= ks)                        # 22-bit fixed point, rows sum to one


def test_c_abi_argument_errors():
    import ctypes
    from chess_vision_b200 import _native
    L = _native.lib()
    ks = ctypes.c_int(0)
    assert L.cv_resize_coeffs_host(0, 4, ctypes.byref(ks), None, None, 0) == -1
    small = np.zeros(3, dtype=np.int32)
    assert L.cv_resize_coeffs_host(400, 256, ctypes.byref(ks), None, small.ctypes.data_as(ctypes.c_void_p), 3) == -1
    assert b"too small" in L.cv_last_error()
    assert L.cv_resize_bilinear_u8(None, -1, 4, 4, None, 4, 4, None) == -1
    assert L.cv_resize_bilinear_u8(None, 1, 4, 4, None, 2, 2, None) == -1           # null pointers
    assert L.cv_resize_bilinear_u8(None, 0, 4, 4, None, 2, 2, None) == 0            # empty batch is a no-op
