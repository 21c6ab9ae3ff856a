"""The JPEG decode restatement (oracle/jpeg_oracle.py: libjpeg's ISLOW IDCT, fancy upsampling, YCbCr tables) against Pillow itself:
the committed golden files Pillow wrote and decoded in the build container (tests/golden/jpeg_reference.npz) and, where Pillow is
importable, live decodes of freshly written files.  Integer work: every comparison is bit-exact."""
import io
import os

import numpy as np
import pytest

from oracle import jpeg_oracle as jo

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "jpeg_reference.npz"))
N_CASES = len(GOLD["names"])


@pytest.mark.parametrize("i", range(N_CASES))
def test_oracle_decodes_golden_files_bit_exactly(i):
    got = jo.decode(GOLD[f"file{i}"].tobytes())
    assert got.shape == GOLD[f"rgb{i}"].shape and np.array_equal(got, GOLD[f"rgb{i}"]), str(GOLD["names"][i])


def test_oracle_matches_live_pillow():
    Image = pytest.importorskip("PIL.Image")
    from oracle import resize_oracle
    for seed, (h, w) in enumerate(((24, 31), (48, 48), (7, 50))):
        img = resize_oracle.synth_image(seed, h, w)
        for ss in (0, 1, 2):
            b = io.BytesIO()
            Image.fromarray(img).save(b, "JPEG", quality=70 + 10 * ss, subsampling=ss)
            want = np.asarray(Image.open(io.BytesIO(b.getvalue())).convert("RGB"))
            assert np.array_equal(jo.decode(b.getvalue()), want), (h, w, ss)


def test_unsupported_files_are_refused():
    Image = pytest.importorskip("PIL.Image")
    img = np.zeros((16, 16, 3), np.uint8)
    b = io.BytesIO()
    Image.fromarray(img).save(b, "JPEG", progressive=True)
    with pytest.raises(jo.UnsupportedJpeg, match="progressive"):
        jo.decode(b.getvalue())
    b = io.BytesIO()
    Image.fromarray(np.zeros((16, 16, 4), np.uint8), "CMYK").save(b, "JPEG")
    with pytest.raises(jo.UnsupportedJpeg):
        jo.decode(b.getvalue())
    with pytest.raises(jo.UnsupportedJpeg):
        jo.decode(b"\x89PNG....")


# ---- the PRODUCT's entropy decoder, host build (csrc/jpeg.cu huff_decode_interval: the routine the device kernel runs) ------------------
@pytest.mark.parametrize("i", range(N_CASES))
def test_native_entropy_decoder_matches_the_oracle(i):
    from chess_vision_b200 import preprocess
    data = GOLD[f"file{i}"].tobytes()
    hdr = jo.parse_headers(data)
    want = jo.decode_coefficients(data, hdr)
    got = preprocess.jpeg_coefficients_host(data)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.shape == w.shape and np.array_equal(g, w), str(GOLD["names"][i])
    assert preprocess.jpeg_info(data) == (hdr["width"], hdr["height"], len(hdr["comps"]))


def test_native_parser_refuses_what_the_oracle_refuses():
    Image = pytest.importorskip("PIL.Image")
    from chess_vision_b200 import _native, preprocess
    b = io.BytesIO()
    Image.fromarray(np.zeros((16, 16, 3), np.uint8)).save(b, "JPEG", progressive=True)
    assert preprocess.jpeg_info(b.getvalue()) is None and b"progressive" in _native.lib().cv_last_error()
    assert preprocess.jpeg_info(b"\x89PNG\r\n\x1a\n" + b"\0" * 32) is None
    assert preprocess.jpeg_info(GOLD["file0"].tobytes()[:100]) is None          # truncated header
