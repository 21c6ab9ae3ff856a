"""Drop-in boundary on the host side: build_model contract, state_dict layout, weight packer, C-ABI symbols.
No GPU compute here (every compute entry point needs a B200)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import chess_vision_b200 as cv
from chess_vision_b200 import _native, arch, synthetic, weights
from oracle import square_oracle as oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_build_model_contract(square_cfg):
    m = cv.build_model(square_cfg)
    assert isinstance(m, cv.ChessSquareCNN) and isinstance(m, torch.nn.Module)
    assert (m.square_overlap, m.square_input_size, m.feature_dim) == (1.5, 64, 480)
    with pytest.raises(KeyError):                                   # models/__init__.py:16 cfg["model"]
        cv.build_model({})
    with pytest.raises(ValueError, match="Unknown architecture: 'resnet'"):      # models/__init__.py:25-28
        cv.build_model({"model": {"arch": "resnet"}})
    with pytest.raises(NotImplementedError):                        # default arch is "vit" (models/__init__.py:17)
        cv.build_model({"model": {}})
    with pytest.raises(RuntimeError, match="pretrained"):           # build_square default pretrained=True, offline
        cv.build_model({"model": {"arch": "square"}})


def test_state_dict_is_the_reference_layout(square_cfg, golden, gold_state):
    _, meta = golden
    m = cv.build_model(square_cfg)
    sd = m.state_dict()
    assert sorted(sd.keys()) == sorted(meta["keys"]) and len(sd) == 288
    for k, v in sd.items():
        assert list(v.shape) == meta["shapes"][k], k
    assert sum(p.numel() for p in m.parameters()) == meta["n_params"] == 2_929_231
    assert sd["class_to_type"].dtype == torch.int64 and sd["class_to_type"].tolist() == meta["class_tables"]["type"]
    m.load_state_dict(gold_state, strict=True)                       # predict.py:57
    bad = dict(gold_state); bad.pop("turn_head.bias")
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad, strict=True)


def test_no_cpu_fallback(square_cfg):
    m = cv.build_model(square_cfg)
    with pytest.raises(RuntimeError, match="inference-only"):
        m(torch.zeros(1, 3, 256, 256))
    m.eval()
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(torch.zeros(1, 3, 256, 256))
    with pytest.raises(RuntimeError, match="CUDA"):
        cv.models.common.combine_type_color(torch.zeros(1, 7), torch.zeros(1, 3))


def test_arch_table_matches_native_and_oracle(golden):
    _, meta = golden
    table = _native.layer_table()
    assert len(table) == arch.NUM_LAYERS == 45
    lay = {n: o for n, o, _ in arch.BLOB_LAYOUT}
    for l, t in zip(arch.LAYERS, table):
        assert (t.kind, t.cin, t.cout, t.k, t.stride, t.relu, t.hin, t.hout, t.skip) == \
               (l.kind, l.cin, l.cout, l.k, l.stride, int(l.relu), l.hin, l.hout, l.skip)
        assert t.w_offset == lay[f"L{l.index}.w"] and t.b_offset == lay[f"L{l.index}.b"]
        shape = meta["shapes"]["backbone." + l.conv_key]
        assert shape == [l.cout, l.cin // l.groups, l.k, l.k], l.key
    assert _native.lib().cv_weight_blob_floats() == arch.BLOB_FLOATS
    assert arch.TRUNK_MACS_PER_CROP == 5_130_368
    kinds = [l.kind for l in arch.LAYERS]
    assert (kinds.count(arch.DENSE), kinds.count(arch.POINTWISE), kinds.count(arch.DEPTHWISE)) == (3, 27, 15)


def test_packer_folds_bn_like_the_oracle(gold_state):
    blob = weights.pack_state_dict(gold_state).numpy()
    off = {n: (o, c) for n, o, c in arch.BLOB_LAYOUT}
    for l in (arch.LAYERS[0], arch.LAYERS[3], arch.LAYERS[7], arch.LAYERS[23], arch.LAYERS[44]):
        w, b = oracle.fold_bn(gold_state, "backbone." + l.conv_key, "backbone." + l.bn_key)
        o, n = off[f"L{l.index}.w"]
        got = blob[o:o + n]
        if l.kind == arch.DEPTHWISE:
            want = w[:, 0].permute(1, 2, 0).reshape(-1)
        else:
            want = w.permute(2, 3, 1, 0).reshape(-1)
        np.testing.assert_allclose(got, want.numpy(), rtol=1e-6, atol=1e-9)
        o, n = off[f"L{l.index}.b"]
        np.testing.assert_allclose(blob[o:o + n], b.numpy(), rtol=1e-6, atol=1e-9)
    o, n = off["glob_w"]
    assert np.array_equal(blob[o:o + n], gold_state["global_head.1.weight"].numpy().reshape(-1))
    o, n = off["tc_w"]
    assert np.array_equal(blob[o:o + 64], gold_state["turn_head.weight"].numpy().reshape(-1))
    o, n = off["head_w"]
    assert np.array_equal(blob[o + 7 * 480:o + n], gold_state["color_head.1.weight"].numpy().reshape(-1))


def test_layer_scale_folding(gold_state):
    sd = dict(gold_state)
    g = torch.full((48,), 0.5)
    sd["backbone.blocks.2.1.layer_scale.gamma"] = g
    l = arch.layer_by_key("blocks.2.1.pw_proj")
    w0, b0 = weights.fold_layer(gold_state, l)
    w1, b1 = weights.fold_layer(sd, l)
    np.testing.assert_allclose(w1, w0 * 0.5); np.testing.assert_allclose(b1, b0 * 0.5)


def test_norm_lut_is_the_reference_transform():
    lut = weights.norm_lut()
    u8 = np.arange(256, dtype=np.uint8).reshape(1, 16, 16, 1).repeat(3, axis=3)
    ref = oracle.normalize_u8(u8)                                    # ToTensor + Normalize arithmetic
    for c in range(3):
        assert torch.equal(lut[c], ref[0, c].reshape(-1))


def test_c_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "chessvision_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = set(re.findall(r"\b(cv_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 20
    lib = ctypes.CDLL(_native.LIB_PATH)
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/chessvision_b200.h but not exported"
    assert lib.cv_abi_version() == 1


def test_c_abi_argument_errors_without_gpu():
    L = _native.lib()
    info = _native.LayerInfo()
    assert L.cv_layer_info_get(45, ctypes.byref(info)) == -1 and b"out of range" in L.cv_last_error()
    h = ctypes.c_void_p()
    if not torch.cuda.is_available():
        assert L.cv_square_create(0, ctypes.byref(h)) == -2          # CV_ERR_CUDA: no CPU fallback
        assert b"no CPU fallback" in L.cv_last_error()
    assert L.cv_square_fen(None, None, None, None, -1, None, None, None) == -1


def test_crop_index_table_bit_exact():
    L = _native.lib()
    for H in (256, 512, 224, 320, 64):
        y0 = np.zeros((8, 64), np.int32); y1 = np.zeros((8, 64), np.int32); lam = np.zeros(64, np.float32)
        _native.check(L.cv_crop_index_table(H, y0.ctypes.data, y1.ctypes.data, lam.ctypes.data))
        oy0, oy1, olam = oracle.crop_index_table(H)
        assert np.array_equal(y0, oy0) and np.array_equal(y1, oy1) and np.array_equal(lam, olam), H
    assert L.cv_crop_index_table(100, y0.ctypes.data, y1.ctypes.data, lam.ctypes.data) == -1


def test_synth_host_matches_numpy():
    L = _native.lib()
    for H, dist, lay in ((256, 1, 0), (256, 0, 0), (64, 1, 1), (512, 1, 0)):
        n = 2
        buf = np.zeros((n, H, H, 3) if lay == 0 else (n, 3, H, H), np.uint8)
        fl = np.zeros(n, np.uint8)
        _native.check(L.cv_synth_boards_host(buf.ctypes.data, lay, 12345678901, n, H, 7, dist, fl.ctypes.data))
        want = synthetic.synth_boards(12345678901, n, H, 7, dist, lay)
        assert np.array_equal(buf, want)
        assert np.array_equal(fl, synthetic.synth_flipped(12345678901, n, 7))
    a = synthetic.synth_boards(0, 4, 64, 1)                          # any sharding regenerates the same boards
    b = np.concatenate([synthetic.synth_boards(0, 1, 64, 1), synthetic.synth_boards(1, 3, 64, 1)])
    assert np.array_equal(a, b)
    assert 0.3 < synthetic.synth_flipped(0, 4096, 1).mean() < 0.7


def test_fen_host_hook_matches_reference(golden):
    arrays, meta = golden
    L = _native.lib()
    rec = ctypes.create_string_buffer(80)
    cast = (ctypes.c_float * 4)(-1.0, 0.5, 0.0, 2.0)
    for labels, want in zip(arrays["rand_labels"], meta["rand_labels_fen"]):
        lab = np.ascontiguousarray(labels.astype(np.int8))
        n = L.cv_fen_from_classes_host(lab.ctypes.data, ctypes.c_float(0.25), cast, rec)
        assert n > 0 and rec.raw[:n].decode() == want + " b Qq" and rec.raw[n:] == b"\0" * (80 - n)
    empty = np.zeros(64, np.int8)
    none = (ctypes.c_float * 4)(0, 0, 0, 0)
    n = L.cv_fen_from_classes_host(empty.ctypes.data, ctypes.c_float(0.0), none, rec)
    assert rec.raw[:n] == b"8/8/8/8/8/8/8/8 w -"                    # shortest record
    full = np.full(64, 12, np.int8)
    allc = (ctypes.c_float * 4)(1, 1, 1, 1)
    n = L.cv_fen_from_classes_host(full.ctypes.data, ctypes.c_float(1.0), allc, rec)
    assert n == 78 and rec.raw[:n].decode() == "/".join(["k" * 8] * 8) + " b KQkq"   # longest record
    bad = np.full(64, 13, np.int8)
    assert L.cv_fen_from_classes_host(bad.ctypes.data, ctypes.c_float(0.0), none, rec) == -1


def test_python_constants_follow_the_header():
    """The enums of include/chessvision_b200.h and their Python mirrors (_native.py, ChessSquareCNN.IMPL_*, evaluate.py counters) must
    not drift apart (round 1 shipped IMPL_DEFAULT = 511 against CV_IMPL_DEFAULT = 1023)."""
    import os
    import re
    from chess_vision_b200 import _native, evaluate
    from chess_vision_b200.models.square import ChessSquareCNN
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "chessvision_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    enums = {k: int(v) for k, v in re.findall(r"\b(CV_[A-Z0-9_]+)\s*=\s*(-?\d+)", hdr)}
    assert (enums["CV_PRECISION_FP32"], enums["CV_PRECISION_BF16"], enums["CV_PRECISION_FP16"], enums["CV_PRECISION_FP32_SPLIT"]) == \
        (_native.PRECISION_FP32, _native.PRECISION_BF16, _native.PRECISION_FP16, _native.PRECISION_FP32_SPLIT)
    assert (enums["CV_LAYOUT_HWC"], enums["CV_LAYOUT_CHW"], enums["CV_FEN_STRIDE"]) == (_native.LAYOUT_HWC, _native.LAYOUT_CHW, _native.FEN_STRIDE)
    for name in ("POINTWISE_UMMA", "DENSE_UMMA", "DEPTHWISE_VEC", "SPLIT_WEIGHTS", "FRONTEND", "TAIL", "MID", "EARLY", "FRONTEND3", "DEFAULT", "ALL"):
        assert enums["CV_IMPL_" + name] == getattr(_native, "IMPL_" + name) == getattr(ChessSquareCNN, "IMPL_" + name), name
    assert enums["CV_PROF_SLOTS"] == ChessSquareCNN.PROF_SLOTS == len(ChessSquareCNN.PROF_NAMES)
    assert enums["CV_EVAL_COUNTERS"] == evaluate.N_COUNTERS and enums["CV_EVAL_CONFUSION"] == evaluate.CONFUSION
    assert enums["CV_EVAL_TURN_CONFUSION"] == evaluate.TURN_CONFUSION and enums["CV_EVAL_PIECE_TOTAL"] == evaluate.PIECE_TOTAL
    assert _native.lib().cv_abi_version() == int(re.search(r"#define CV_ABI_VERSION (\d+)", hdr).group(1))

