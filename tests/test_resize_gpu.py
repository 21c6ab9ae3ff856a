"""Board resize kernel (cv_resize_bilinear_u8 through chess_vision_b200.preprocess) on the GPU: bit-exact against Pillow's golden
outputs, against the CPU oracle on ragged / batched inputs, and end to end (PNG files -> FEN) against the host transform."""
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import resize_oracle as ro

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "resize_reference.npz"))
CASES = [tuple(int(v) for v in c) for c in GOLD["cases"]]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[1]}x{c[2]}to{c[3]}x{c[4]}")
def test_kernel_is_bit_exact_with_pillow_golden(case):
    from chess_vision_b200.preprocess import resize_boards
    seed, h, w, oh, ow, whole, crc = case
    img = torch.from_numpy(ro.synth_image(seed, h, w)).cuda()
    out = resize_boards(img[None], (oh, ow))[0].cpu().numpy()
    assert zlib.crc32(out.tobytes()) == crc                                        # integer work: bit-exact
    if whole:
        assert np.array_equal(out, GOLD[f"out_{seed}"])


def test_batch_ragged_sizes_and_unaligned_views_match_oracle():
    from chess_vision_b200.preprocess import resize_boards
    for seed, (B, h, w, oh, ow) in enumerate([(5, 61, 47, 33, 29), (3, 400, 400, 256, 256), (2, 97, 131, 256, 256), (4, 16, 16, 64, 64),
                                              (2, 640, 480, 256, 256), (1, 1200, 1200, 256, 256)]):
        imgs = np.stack([ro.synth_image(1000 + 10 * seed + b, h, w) for b in range(B)])
        ref = np.stack([ro.resize_bilinear_u8(imgs[b], oh, ow) for b in range(B)])
        out = resize_boards(torch.from_numpy(imgs).cuda(), (oh, ow))
        assert np.array_equal(out.cpu().numpy(), ref)
    # a destination whose base is not 4-byte aligned (a view one byte into a buffer): the byte-store path
    imgs = np.stack([ro.synth_image(77, 50, 50)])
    buf = torch.zeros(1 + 1 * 21 * 21 * 3, dtype=torch.uint8, device="cuda")
    out = buf[1:].view(1, 21, 21, 3)
    resize_boards(torch.from_numpy(imgs).cuda(), 21, out=out)
    assert np.array_equal(out.cpu().numpy()[0], ro.resize_bilinear_u8(imgs[0], 21, 21))
    # empty batch
    assert resize_boards(torch.empty((0, 40, 40, 3), dtype=torch.uint8, device="cuda"), 32).shape == (0, 32, 32, 3)


def test_full_size_batch_properties():
    """BASELINE-size batch (4096 boards of 400x400 -> 256x256): every image equals the same image resized alone; a constant image
    stays constant; the result does not depend on how the batch is split."""
    from chess_vision_b200.preprocess import resize_boards
    base = torch.from_numpy(np.stack([ro.synth_image(s, 400, 400) for s in range(8)])).cuda()
    big = base.repeat(512, 1, 1, 1)
    big[4095] = 173
    out = resize_boards(big, 256)
    ref8 = resize_boards(base, 256)
    assert torch.equal(out[:4088].view(511, 8, 256, 256, 3), ref8[None].expand(511, -1, -1, -1, -1))
    assert bool((out[4095] == 173).all())
    assert np.array_equal(ref8[3].cpu().numpy(), ro.resize_bilinear_u8(base[3].cpu().numpy(), 256, 256))
    halves = torch.cat([resize_boards(big[:1000], 256), resize_boards(big[1000:], 256)])
    assert torch.equal(halves, out)


def test_files_to_fen_matches_host_transform(tmp_path, gold_state):
    """predict_images (decode on host, resize + model + FEN on the device) returns the strings of predict() with the reference's
    host transform (PIL Resize -> ToTensor -> Normalize) in fp32 mode: the resized bytes are identical, so the inputs are."""
    from PIL import Image
    import chess_vision_b200 as cv
    model = cv.build_model({"model": {"arch": "square", "pretrained": False}})
    model.load_state_dict(gold_state)
    model = model.cuda().eval()
    model.precision = "fp32"
    paths = []
    for i, (h, w) in enumerate([(400, 400), (400, 400), (512, 512), (300, 280)]):
        p = str(tmp_path / f"b{i}.png")
        Image.fromarray(ro.synth_image(200 + i, h, w)).save(p)
        paths.append(p)
    fast = cv.predict_images(model, paths, input_size=256)
    tf = cv.get_transform(input_size=256)
    slow = [cv.predict(model, p, tf, torch.device("cuda")) for p in paths]
    assert fast == slow
