"""Checkpoint bridge: the reference's checkpoint dict (train.py:458-467 / predict.py:53-57) and the packed-blob cache."""
import os

import numpy as np
import pytest
import torch

import chess_vision_b200 as cv
from chess_vision_b200 import arch, checkpoint, synthetic, weights

CFG = {"model": {"arch": "square", "name": "mobilenetv4_conv_small_050.e3000_r224_in1k", "pretrained": True, "input_size": 256}}


def make_state():
    m = cv.build_model({"model": {"arch": "square", "pretrained": False}})
    return synthetic.init_state_dict(m.state_dict(), 3)


def test_reference_checkpoint_file_roundtrip(tmp_path):
    state = make_state()
    wrapped = {"_orig_mod.module." + k: v for k, v in state.items()}            # torch.compile + DDP prefixes
    ckpt = {"epoch": 3, "model": wrapped, "optimizer": {"state": {}, "param_groups": []}, "scheduler": {}, "scaler": {},
            "best_val_acc": 0.5, "config": CFG}
    p = str(tmp_path / "latest.pth")
    torch.save(ckpt, p)
    model, cfg = checkpoint.load_checkpoint(p, device="cpu")
    assert cfg["model"]["pretrained"] is False and not model.training
    got = model.state_dict()
    assert list(got.keys()) == list(state.keys()) and len(got) == 288
    assert all(torch.equal(got[k], state[k]) for k in state)
    with pytest.raises(KeyError):
        torch.save({"model": state}, p)
        checkpoint.load_checkpoint(p, device="cpu")
    with pytest.raises(RuntimeError):                                             # strict load, like the reference
        bad = dict(state)
        bad.pop("turn_head.bias")
        torch.save({"model": bad, "config": CFG}, p)
        checkpoint.load_checkpoint(p, device="cpu")


def test_packed_file_roundtrip_and_corruption(tmp_path):
    state = make_state()
    p = str(tmp_path / "w.cvb")
    crc = checkpoint.save_packed(p, state, CFG)
    blob, cfg = checkpoint.load_packed(p)
    assert cfg == CFG and blob.numel() == arch.BLOB_FLOATS
    assert torch.equal(blob, weights.pack_state_dict(state))
    import zlib
    assert crc == zlib.crc32(blob.numpy().tobytes())
    raw = bytearray(open(p, "rb").read())
    raw[-5] ^= 0x40
    open(p, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="CRC"):
        checkpoint.load_packed(p)
    open(p, "wb").write(bytes(raw[:-8]))
    with pytest.raises(ValueError, match="floats"):
        checkpoint.load_packed(p)
    open(p, "wb").write(b"not a weights file at all")
    with pytest.raises(ValueError, match="not a chess_vision_b200"):
        checkpoint.load_packed(p)


@pytest.mark.gpu
def test_packed_model_equals_state_dict_model(tmp_path):
    state = make_state()
    p, q = str(tmp_path / "latest.pth"), str(tmp_path / "w.cvb")
    torch.save({"model": state, "config": CFG, "epoch": 0}, p)
    m1, cfg = checkpoint.load_checkpoint(p, device="cuda")
    checkpoint.save_packed(q, m1.state_dict(), cfg)
    m2, _ = checkpoint.model_from_packed(q, device="cuda")
    u8 = torch.from_numpy(synthetic.synth_boards(0, 16, 256, 1, synthetic.DIST_STRUCTURED)).cuda()
    for prec in ("fp32", "bf16"):
        a, b = m1.forward_u8(u8, precision=prec), m2.forward_u8(u8, precision=prec)
        assert all(torch.equal(a[k], b[k]) for k in a), prec
    assert m1.predict_fen(u8) == m2.predict_fen(u8)
