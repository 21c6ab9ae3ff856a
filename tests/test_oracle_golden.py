"""The CPU oracle (oracle/square_oracle.py) pinned against outputs of the reference itself.

tests/golden/ was produced in the build container by oracle/make_golden.py, which imports the reference's
own models/square.py, models/common.py, dataset.py and predict.py unmodified (timm restated by a shim).
"""
import numpy as np
import pytest
import torch

from chess_vision_b200 import synthetic
from oracle import square_oracle as oracle


def _boards(H, n, seed):
    return synthetic.synth_boards(0, n, H, seed, synthetic.DIST_STRUCTURED)


@pytest.mark.parametrize("H,n", [(256, 2), (512, 1)])
def test_crop_stage_matches_reference(golden, H, n):
    arrays, meta = golden
    x = oracle.normalize_u8(_boards(H, n, meta["board_seed"]))
    assert torch.equal(x, synthetic.normalize_boards(_boards(H, n, meta["board_seed"])))
    crops = oracle.crop_squares(x).numpy()
    assert crops.shape == (n * 64, 3, 64, 64)
    np.testing.assert_allclose(crops[[0, 7, 27, 63]], arrays[f"crops{H}_sample"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(crops[9, :, 0, :], arrays[f"crops{H}_crop9_row0"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(crops.astype(np.float64).sum(axis=(1, 2, 3)), arrays[f"crops{H}_sum"], rtol=0, atol=2e-3)


def test_crop_geometry_values():
    assert oracle.crop_geometry(256) == (32, 48, 8)          # square.py:53-55 at config_square.yaml:12
    assert oracle.crop_geometry(512) == (64, 96, 16)
    i0, i1, lam = oracle.bilinear_taps(48)
    assert set(np.round(lam, 6).tolist()) <= {0.0, 0.125, 0.375, 0.625, 0.875}
    assert i0[0] == 0 and lam[0] == 0.0 and i1[-1] == 47
    _, _, lam512 = oracle.bilinear_taps(96)
    assert set(lam512.tolist()) == {0.25, 0.75}
    y0, y1, _ = oracle.crop_index_table(256)
    assert y0.min() == 0 and y1.max() == 255 and y0[0, 0] == 0 and y0[1, 0] == 24     # replicate-pad clamp / stride


@pytest.mark.parametrize("H,n", [(256, 8), (512, 2)])
def test_forward_matches_reference(golden, gold_state, H, n):
    arrays, meta = golden
    x = oracle.normalize_u8(_boards(H, n, meta["board_seed"]))
    out = oracle.forward(x, gold_state, return_features=True)
    for k in ("squares", "turn", "castling"):
        ref = arrays[f"{k}{H}"]
        got = out[k].numpy()
        assert got.shape == ref.shape
        err = np.abs(got - ref).max() / np.abs(ref).max()
        assert err < 1e-5, (k, err)
    np.testing.assert_allclose(out["features"][:64].numpy(), arrays[f"features{H}_board0"], rtol=1e-4, atol=1e-5)
    fens = oracle.fen_strings(out["squares"].numpy(), out["turn"].numpy(), out["castling"].numpy())
    assert fens == meta[f"fen{H}"]
    if H == 256:
        assert fens[:2] == meta["predict_png_fen"]           # the reference's predict() through a PNG


def test_golden_is_not_vacuous(golden):
    arrays, meta = golden
    classes = np.concatenate([arrays["squares256"].reshape(-1, 13).argmax(-1), arrays["squares512"].reshape(-1, 13).argmax(-1)])
    assert set(classes.tolist()) == set(range(13))           # all 13 classes occur
    assert any(ch.isdigit() for f in meta["fen256"] for ch in f.split()[0])
    assert len({f.split()[1] for f in meta["fen256"]}) == 2 or len({f.split()[2] for f in meta["fen256"]}) > 1


def test_state_dict_layout(golden, gold_state):
    _, meta = golden
    assert meta["n_params"] == 2_929_231 and len(meta["keys"]) == 288          # README.md:11
    assert list(gold_state.keys()) == meta["keys"]


def test_combine_is_raw_logit_addition():
    t = torch.arange(14, dtype=torch.float32).reshape(2, 7)
    c = torch.tensor([[100., 200., 300.], [1000., 2000., 3000.]])
    j = oracle.combine_type_color(t, c)
    assert j.shape == (2, 13)
    assert j[0].tolist() == [100.] + [200. + i for i in range(1, 7)] + [300. + i for i in range(1, 7)]


def test_fold_bn_equals_conv_then_bn(gold_state):
    import torch.nn.functional as F
    w, b = oracle.fold_bn(gold_state, "backbone.blocks.1.0.conv.weight", "backbone.blocks.1.0.bn1")
    x = torch.randn(2, 16, 16, 16, dtype=torch.float64)
    sd = {k: v.double() if v.is_floating_point() else v for k, v in gold_state.items()}
    ref = F.batch_norm(F.conv2d(x, sd["backbone.blocks.1.0.conv.weight"], None, 2, 1), sd["backbone.blocks.1.0.bn1.running_mean"],
                       sd["backbone.blocks.1.0.bn1.running_var"], sd["backbone.blocks.1.0.bn1.weight"],
                       sd["backbone.blocks.1.0.bn1.bias"], False, 0.0, 1e-5)
    np.testing.assert_allclose(F.conv2d(x, w, b, 2, 1).numpy(), ref.numpy(), rtol=1e-10, atol=1e-10)
