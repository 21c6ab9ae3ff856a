"""The CUDA JPEG decoder (cv_jpeg_decode_batch: device Huffman, integer IDCT, fancy upsampling, YCbCr -> RGB) against Pillow's golden
decodes, the CPU oracle and -- where Pillow is importable -- live Pillow; and `predict_images` on .jpg files against the
reference-surface `predict()` (predict.py:18-42, PIL decode + torchvision transform).  Integer work: bit-exact."""
import io
import os

import numpy as np
import pytest
import torch

from chess_vision_b200 import _native, preprocess
from oracle import jpeg_oracle as jo
from oracle import resize_oracle

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "jpeg_reference.npz"))
N_CASES = len(GOLD["names"])


@pytest.mark.parametrize("host_entropy", [False, True])
def test_golden_files_decode_bit_exactly(host_entropy):
    for i in range(N_CASES):
        got = preprocess.decode_jpegs([GOLD[f"file{i}"].tobytes()], "cuda", entropy_on_host=host_entropy)[0].cpu().numpy()
        assert got.shape == GOLD[f"rgb{i}"].shape and np.array_equal(got, GOLD[f"rgb{i}"]), (str(GOLD["names"][i]), host_entropy)


def test_mixed_batch_of_one_size():
    """One call, many files of one size with different subsamplings / qualities / restart intervals, more files than one launch block."""
    Image = pytest.importorskip("PIL.Image")
    files, want = [], []
    for k in range(150):
        img = resize_oracle.synth_image(500 + k, 72, 88)
        b = io.BytesIO()
        kw = {"restart_marker_blocks": 1 + k % 5} if k % 3 == 0 else ({"optimize": True} if k % 3 == 1 else {})
        Image.fromarray(img).save(b, "JPEG", quality=40 + (k * 7) % 60, subsampling=k % 3, **kw)
        files.append(b.getvalue())
        want.append(np.asarray(Image.open(io.BytesIO(files[-1])).convert("RGB")))
    want = np.stack(want)
    for host_entropy in (False, True):
        got = preprocess.decode_jpegs(files, "cuda", entropy_on_host=host_entropy).cpu().numpy()
        assert np.array_equal(got, want), host_entropy
    assert np.array_equal(jo.decode(files[7]), want[7])                      # and the CPU oracle agrees with both


def test_size_mismatch_and_unsupported_files_raise():
    a, b = GOLD["file0"].tobytes(), GOLD["file3"].tobytes()                 # 40x56 and 17x23
    with pytest.raises(_native.NativeError, match="the batch is"):
        preprocess.decode_jpegs([a, b], "cuda")
    with pytest.raises(_native.NativeError):
        preprocess.decode_jpegs([b"\x89PNG\r\n\x1a\n" + b"\0" * 64], "cuda")
    assert preprocess.decode_jpegs([], "cuda").shape[0] == 0


def test_predict_images_on_jpeg_files_equals_reference_predict(gpu_model, tmp_path):
    """.jpg boards (datagen's format: quality 90, 4:2:0) of two sizes + a PNG: `predict_images` (device decode + device resize) must print
    the strings `predict()` prints through PIL + torchvision (predict.py:19-20), in fp32 mode where the path is exact."""
    Image = pytest.importorskip("PIL.Image")
    import chess_vision_b200 as cv
    from chess_vision_b200 import synthetic
    transform = cv.get_transform("mobilenetv4_conv_small_050.e3000_r224_in1k", False, 256)
    u8 = synthetic.synth_boards(0, 5, 256, 1, synthetic.DIST_STRUCTURED)
    paths = []
    for i in range(5):
        img = Image.fromarray(u8[i])
        if i in (1, 3):
            img = img.resize((400, 400), Image.BILINEAR)                      # a larger render: decode at 400x400, then the Pillow-exact resize
        p = tmp_path / (f"b{i}.png" if i == 4 else f"b{i}.jpg")
        if i == 4:
            img.save(p)
        else:
            img.save(p, "JPEG", quality=90, subsampling=2)
        paths.append(str(p))
    gpu_model.precision = "fp32"
    try:
        want = [cv.predict(gpu_model, p, transform, torch.device("cuda")) for p in paths]
        got = preprocess.predict_images(gpu_model, paths, 256)
    finally:
        gpu_model.precision = "fp16"
    assert got == want


def test_chunked_device_entropy_decoder_on_hard_streams():
    """Files whose Huffman streams are slow to resynchronise (noise at quality 100: the chain of chunk states needs many rounds, some
    files fall back to the one-thread kernel) and files far larger than a chunk, many per call: still Pillow's pixels, byte for byte."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(9)
    files, want = [], []
    for k in range(40):
        img = rng.integers(0, 256, (208, 176, 3), dtype=np.uint8)
        if k % 4 == 1:
            img = (img // 8 + resize_oracle.synth_image(900 + k, 208, 176) // 2).astype(np.uint8)
        b = io.BytesIO()
        Image.fromarray(img).save(b, "JPEG", quality=(100, 97, 90, 35)[k % 4], subsampling=(0, 2, 1, 2)[(k // 4) % 4])
        files.append(b.getvalue())
        want.append(np.asarray(Image.open(io.BytesIO(files[-1])).convert("RGB")))
    got = preprocess.decode_jpegs(files, "cuda").cpu().numpy()
    assert np.array_equal(got, np.stack(want))
    again = preprocess.decode_jpegs(files, "cuda").cpu().numpy()          # scratch arrays are reused: stale chunk states must not matter
    assert np.array_equal(again, got)
    host = preprocess.decode_jpegs(files, "cuda", entropy_on_host=True).cpu().numpy()
    assert np.array_equal(host, got)


def test_pipelined_jpeg_prediction_equals_the_plain_path(gpu_model, tmp_path):
    """`predict_jpeg_files` (decode of chunk k + 1 on a helper thread under the forward of chunk k, ragged last chunk, a resize in between,
    flipped boards) and `predict_images` on the same files as paths: the strings of decode -> resize -> predict_fen done plainly."""
    Image = pytest.importorskip("PIL.Image")
    from chess_vision_b200 import synthetic
    u8 = synthetic.synth_boards(0, 37, 256, 1, synthetic.DIST_STRUCTURED)
    files, paths = [], []
    for i in range(37):
        b = io.BytesIO()
        Image.fromarray(u8[i]).resize((320, 320), Image.BILINEAR).save(b, "JPEG", quality=90, subsampling=2)
        files.append(b.getvalue())
        p = tmp_path / f"b{i}.jpg"
        p.write_bytes(files[-1])
        paths.append(str(p))
    flipped = [i % 3 == 0 for i in range(37)]
    boards = preprocess.resize_boards(preprocess.decode_jpegs(files, "cuda"), 256)
    want = gpu_model.predict_fen(boards, flipped=torch.tensor(flipped, dtype=torch.uint8))
    assert preprocess.predict_jpeg_files(gpu_model, files, 256, flipped=flipped, chunk=16) == want
    assert preprocess.predict_jpeg_files(gpu_model, files, 256, flipped=flipped) == want
    assert preprocess.predict_images(gpu_model, paths, 256, flipped=flipped) == want
    fen, ln = preprocess.predict_jpeg_files(gpu_model, files[:5], 256, device_records=True)
    assert fen.is_cuda and fen.shape == (5, 80) and gpu_model.decode_fen_records(fen, ln) == gpu_model.predict_fen(boards[:5])
    with pytest.raises(_native.NativeError):
        preprocess.predict_jpeg_files(gpu_model, files[:20] + [b"\x89PNG not a jpeg"] + files[20:], 256, chunk=8)
