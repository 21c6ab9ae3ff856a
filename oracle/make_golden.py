"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (build container only).

The reference's own ``models/square.py``, ``models/common.py``, ``models/__init__.py``, ``dataset.py`` and
``predict.py`` are imported UNMODIFIED from /root/reference over ``oracle/timm_shim`` (timm itself is not
installable offline; see the shim's docstring for what that leaves unpinned).  Nothing here is used at
test time on the GPU box: only the committed .npz files travel.

    python oracle/make_golden.py           # rewrites tests/golden/

Inputs are regenerated from seeds at test time (chess_vision_b200/synthetic.py), so the fixtures hold
only reference OUTPUTS plus the small calibration statistics the head weights are derived from.
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "timm_shim"))
sys.path.insert(0, REF)

from chess_vision_b200 import synthetic  # noqa: E402  (input/weight recipes only -- no product compute)

import dataset as ref_dataset  # noqa: E402  /root/reference/dataset.py
import predict as ref_predict  # noqa: E402  /root/reference/predict.py
from models import build_model as ref_build_model  # noqa: E402  /root/reference/models/__init__.py

GOLD = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEED, CAL_SEED, BOARD_SEED = 0, 999, 1
N_CAL = 8


def ref_model():
    cfg = yaml.safe_load(open(os.path.join(REF, "config_square.yaml")))
    cfg["model"]["pretrained"] = False
    m = ref_build_model(cfg)
    m.eval()
    return m, cfg


def batched_ref_fen(outputs):
    """predict.py:27-42 applied per board with the reference's own labels_to_fen."""
    fens = []
    for b in range(outputs["squares"].shape[0]):
        sq = outputs["squares"][b].view(ref_dataset.NUM_SQUARES, ref_dataset.NUM_CLASSES)
        placement = ref_dataset.labels_to_fen(sq.argmax(dim=-1).cpu())
        turn = "b" if outputs["turn"][b].item() > 0 else "w"
        flags = (outputs["castling"][b] > 0).tolist()
        chars = "".join(ch for f, ch in zip(flags, ["K", "Q", "k", "q"]) if f)
        fens.append(f"{placement} {turn} {chars or '-'}")
    return fens


def main():
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    model, cfg = ref_model()
    template = model.state_dict()
    keys = list(template.keys())
    shapes = {k: list(v.shape) for k, v in template.items()}
    n_params = sum(p.numel() for p in model.parameters())
    assert n_params == 2_929_231 and len(keys) == 288

    # --- weights: H1 recipe, heads calibrated on the reference's own features ------------------------
    state = synthetic.init_state_dict(template, WEIGHT_SEED)
    model.load_state_dict(state, strict=True)
    cal_u8 = synthetic.synth_boards(0, N_CAL, 256, CAL_SEED, synthetic.DIST_STRUCTURED)
    with torch.no_grad():
        x = synthetic.normalize_boards(cal_u8)
        feats = model._extract_features(model._crop_squares(x))          # square.py:95-96
    stats = synthetic.calibration_stats(feats, CAL_SEED)
    state = synthetic.calibrate_heads(state, stats, CAL_SEED)
    model.load_state_dict(state, strict=True)

    out = {}
    meta = {"weight_seed": WEIGHT_SEED, "cal_seed": CAL_SEED, "board_seed": BOARD_SEED, "n_cal": N_CAL,
            "n_params": n_params, "keys": keys, "shapes": shapes, "torch": torch.__version__,
            "numpy": np.__version__}
    for k, v in stats.items():
        out["cal_" + k] = np.asarray(v)

    # --- crop stage: reference _crop_squares on float boards (both resolutions) ----------------------
    for H, nb in ((256, 2), (512, 1)):
        u8 = synthetic.synth_boards(0, nb, H, BOARD_SEED, synthetic.DIST_STRUCTURED)
        xb = synthetic.normalize_boards(u8)
        with torch.no_grad():
            crops = model._crop_squares(xb)                               # (nb*64,3,64,64)
        c = crops.numpy()
        out[f"crops{H}_sum"] = c.astype(np.float64).sum(axis=(1, 2, 3))   # per-crop checksum
        out[f"crops{H}_sample"] = c[[0, 7, 27, 63]].copy()                # corner/edge/interior crops in full
        out[f"crops{H}_crop9_row0"] = c[9, :, 0, :].copy()

    # --- full forward + FEN: 256 (config 1 shape) and 512 (config 5 shape) ---------------------------
    for H, nb in ((256, 8), (512, 2)):
        u8 = synthetic.synth_boards(0, nb, H, BOARD_SEED, synthetic.DIST_STRUCTURED)
        xb = synthetic.normalize_boards(u8)
        with torch.no_grad():
            o = model(xb)
            f = model._extract_features(model._crop_squares(xb))
        out[f"squares{H}"] = o["squares"].numpy()
        out[f"turn{H}"] = o["turn"].numpy()
        out[f"castling{H}"] = o["castling"].numpy()
        out[f"features{H}_board0"] = f[:64].numpy()
        meta[f"fen{H}"] = batched_ref_fen(o)

    # --- the reference's predict() end to end through a lossless PNG ---------------------------------
    from PIL import Image
    transform = ref_dataset.get_transform(cfg["model"]["name"], is_training=False, input_size=256)
    u8 = synthetic.synth_boards(0, 2, 256, BOARD_SEED, synthetic.DIST_STRUCTURED)
    fens = []
    with tempfile.TemporaryDirectory() as td:
        for i in range(2):
            p = os.path.join(td, f"b{i}.png")
            Image.fromarray(u8[i]).save(p)
            fens.append(ref_predict.predict(model, p, transform, torch.device("cpu")))
    meta["predict_png_fen"] = fens
    assert fens == meta["fen256"][:2], (fens, meta["fen256"][:2])

    # --- FEN codec known answers straight from the reference's functions ----------------------------
    kat = ["rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR",          # README.md:116
           "1B1B1K2/3p1N2/8/8/8/8/8/1B6",                           # dataset.py:74 style
           "8/8/8/8/8/8/8/8", "pppppppp/" * 7 + "pppppppp", "p1p1p1p1/1p1p1p1p/8/PPPPPPPP/7k/K7/3Q4/4q3"]
    meta["kat_roundtrip"] = {s: ref_dataset.labels_to_fen(ref_dataset.fen_to_labels(s)) for s in kat}
    meta["kat_labels"] = {s: ref_dataset.fen_to_labels(s).tolist() for s in kat}
    meta["filename_kat"] = ref_dataset.filename_to_fen("1B1B1K2-3p1N2-8-8-8-8-8-1B6.jpeg")
    pf = ref_dataset.parse_full_fen("rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR b KQkq -")
    meta["parse_full_kat"] = {"turn": pf["turn"].tolist(), "castling": pf["castling"].tolist(),
                              "squares": pf["squares"].tolist()}
    rng = np.random.default_rng(7)
    rand_labels = rng.integers(0, 13, size=(64, 64)) * (rng.random((64, 64)) < 0.45)
    out["rand_labels"] = rand_labels.astype(np.int8)
    meta["rand_labels_fen"] = [ref_dataset.labels_to_fen(torch.from_numpy(r)) for r in rand_labels]
    meta["class_tables"] = {"type": ref_dataset.CLASS_TO_TYPE, "color": ref_dataset.CLASS_TO_COLOR,
                            "pieces": "".join(ref_dataset.INDEX_TO_PIECE[i] for i in range(13))}

    np.savez_compressed(os.path.join(GOLD, "reference_outputs.npz"), **out)
    json.dump(meta, open(os.path.join(GOLD, "reference_meta.json"), "w"), indent=1)
    hist = np.bincount(np.concatenate([o_.reshape(-1, 13).argmax(-1) for o_ in (out["squares256"], out["squares512"])]), minlength=13)
    print("class histogram:", hist.tolist())
    print("fen256[0]:", meta["fen256"][0])
    print("fen512[0]:", meta["fen512"][0])
    print("wrote", GOLD)


if __name__ == "__main__":
    main()
