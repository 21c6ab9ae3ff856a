"""Guard at the third-party boundary: the trunk arithmetic of the reference lives in `timm` (unpinned, requirements.txt:3), which
cannot be installed here, so `arch.py` / `oracle/timm_shim` restate timm's published `mobilenetv4_conv_small_050`.  Wherever a
GENUINE timm is importable these tests pin the restatement to it -- state_dict keys and shapes, parameter count, num_features,
pretrained_cfg mean/std, and the forward_features + global_pool arithmetic on identical weights -- and fail loudly if timm's
model ever differs.  Without timm (this container, the GPU box) they skip; what then pins the trunk is listed in DESIGN.md
("Oracle and parity")."""
import importlib
import os
import sys

import pytest
import torch


def _real_timm():
    """The installed timm package, never the oracle's shim (which also answers `import timm` when oracle/timm_shim is on the path)."""
    shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "timm_shim")
    saved = list(sys.path)
    sys.path[:] = [p for p in sys.path if os.path.abspath(p) != shim]
    cached = sys.modules.pop("timm", None)
    try:
        mod = importlib.import_module("timm")
        if "chessvision-oracle-shim" in getattr(mod, "__version__", ""):
            return None
        return mod
    except ImportError:
        return None
    finally:
        sys.path[:] = saved
        if cached is not None and "timm" not in sys.modules:
            sys.modules["timm"] = cached


timm = _real_timm()
pytestmark = pytest.mark.skipif(timm is None, reason="genuine timm is not installed (third-party, unpinned; no network here)")
NAME = "mobilenetv4_conv_small_050.e3000_r224_in1k"


@pytest.fixture(scope="module")
def real():
    torch.manual_seed(0)
    return timm.create_model(NAME, pretrained=False, num_classes=0).eval()      # models/square.py:121-125


def test_state_dict_keys_shapes_and_counts(real):
    from chess_vision_b200 import arch
    from chess_vision_b200.models.backbone import create_backbone
    ours = create_backbone(NAME, pretrained=False)
    a, b = real.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert tuple(a[k].shape) == tuple(b[k].shape), k
    assert real.num_features == arch.FEATURE_DIM == ours.num_features
    assert sum(p.numel() for p in real.parameters()) == sum(p.numel() for p in ours.parameters()) == 957_952
    cfg = real.pretrained_cfg
    assert tuple(cfg["mean"]) == (0.485, 0.456, 0.406) and tuple(cfg["std"]) == (0.229, 0.224, 0.225)      # dataset.py:157-160


def test_forward_features_matches_the_oracle_trunk(real):
    """Same weights through genuine timm and through the oracle's restatement of the trunk (what every parity test trusts)."""
    from chess_vision_b200 import synthetic
    from oracle import square_oracle as oracle
    import chess_vision_b200 as cv
    model = cv.build_model({"model": {"arch": "square", "pretrained": False}})
    state = synthetic.init_state_dict(model.state_dict(), 3)
    real.load_state_dict({k[len("backbone."):]: v for k, v in state.items() if k.startswith("backbone.")}, strict=True)
    x = oracle.normalize_u8(synthetic.synth_boards(0, 1, 256, 1, synthetic.DIST_STRUCTURED))
    crops = oracle.crop_squares(x)
    with torch.no_grad():
        want = real.global_pool(real.forward_features(crops)).flatten(1)                                   # models/square.py:86-90
    got = oracle.forward(x, state, return_features=True)["features"]
    assert got.shape == want.shape
    assert float((got - want).abs().max() / want.abs().max()) < 1e-5
