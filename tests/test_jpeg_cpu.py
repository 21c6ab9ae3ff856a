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


# ---- the chunked scheme of the device path (huff_decode_chunk: speculative rounds, chain check, writing pass, DC prefix sums), run on
#      the host round by round: whatever the chunk size and however few rounds (chains that do not close fall back to the serial walk),
#      the coefficients are those of the serial decoder -----------------------------------------------------------------------------------
@pytest.mark.parametrize("i", range(N_CASES))
def test_chunked_entropy_decoder_equals_the_serial_one_on_golden_files(i):
    from chess_vision_b200 import preprocess
    data = GOLD[f"file{i}"].tobytes()
    want = preprocess.jpeg_coefficients_host(data)
    for chunk, rounds in ((-1, 6), (16, 1), (16, 40), (48, 3), (100, 2), (256, 4), (4096, 1)):
        st = []
        got = preprocess.jpeg_coefficients_host(data, chunk_bytes=chunk, rounds=rounds, stats=st)
        assert all(np.array_equal(g, w) for g, w in zip(got, want)), (str(GOLD["names"][i]), chunk, rounds, st)


def test_chunked_entropy_decoder_on_fresh_files_and_its_convergence():
    Image = pytest.importorskip("PIL.Image")
    from chess_vision_b200 import preprocess, synthetic
    rng = np.random.default_rng(5)
    boards = synthetic.synth_boards(0, 3, 256, 1, synthetic.DIST_STRUCTURED)
    images = [boards[0], boards[1], rng.integers(0, 256, (120, 200, 3), dtype=np.uint8), (boards[2] // 2 + rng.integers(0, 128, (256, 256, 3))).astype(np.uint8)]
    fell_back_with_one_round = 0
    for k, img in enumerate(images):
        for quality, sub, kw in ((90, 2, {}), (75, 1, {"optimize": True}), (98, 0, {}), (60, 2, {"restart_marker_blocks": 7})):
            b = io.BytesIO()
            Image.fromarray(img).save(b, "JPEG", quality=quality, subsampling=sub, **kw)
            data = b.getvalue()
            want = preprocess.jpeg_coefficients_host(data)
            for chunk, rounds in ((-1, 6), (64, 1), (64, 3), (200, 2), (512, 8)):
                st = []
                got = preprocess.jpeg_coefficients_host(data, chunk_bytes=chunk, rounds=rounds, stats=st)
                assert all(np.array_equal(g, w) for g, w in zip(got, want)), (k, quality, sub, chunk, rounds, st)
                if chunk == 64 and rounds == 1:
                    fell_back_with_one_round += st[2]
                if chunk == -1 and k < 2 and quality == 90:
                    # the reference's own kind of file (datagen/generate.js: quality 90, 4:2:0): the device path's size rule closes every chain
                    # in at most 4 of its 6 rounds, re-decoding each chunk about once
                    assert st[2] == 0 and st[3] <= 3 and st[1] <= 2.5 * st[0], st
    assert fell_back_with_one_round > 0          # one round cannot close a chain of guesses: the fall-back path is what was compared there
