"""Checkpoint bridge: the reference's training checkpoints -> this package, and a packed-blob cache.

The reference writes ``{"epoch", "model": state_dict, "optimizer", "scheduler", "scaler", "best_val_acc", "config"}`` with
``torch.save`` (train.py:458-467) and reads ``ckpt["model"]`` / ``ckpt["config"]`` back with ``weights_only=True``
(predict.py:53-57, evaluate.py:303-306).  ``load_checkpoint`` accepts exactly that file.  ``save_packed`` / ``load_packed``
store the BN-folded fp32 blob the device consumes (weights.pack_state_dict, 2.3 M floats) with its CRC and the config, so a
serving process starts without torch-unpickling optimizer state and without re-folding BatchNorm.
"""
import json
import struct
import zlib

import numpy as np
import torch

from . import arch
from .models import build_model
from .weights import pack_state_dict

MAGIC = b"CVB200W1"
RAW_MAGIC = b"CVB200S1"
_PREFIXES = ("_orig_mod.", "module.")        # torch.compile / DistributedDataParallel wrappers around the trained model


def clean_state_dict(sd):
    """Strip wrapper prefixes that training may have left on the 288 reference keys."""
    out = {}
    for k, v in sd.items():
        changed = True
        while changed:
            changed = False
            for p in _PREFIXES:
                if k.startswith(p):
                    k, changed = k[len(p):], True
        out[k] = v
    return out


def load_checkpoint(path, device="cuda"):
    """Reference checkpoint file -> (model on `device` in eval mode, config dict).  Same calls as predict.py:53-57."""
    ckpt = torch.load(path, map_location="cpu", weights_only=True)
    if "model" not in ckpt or "config" not in ckpt:
        raise KeyError("checkpoint has no 'model' / 'config' entries (train.py:458-467 writes both)")
    cfg = ckpt["config"]
    cfg["model"]["pretrained"] = False           # the weights come from the file; nothing is downloaded
    model = build_model(cfg)
    model.load_state_dict(clean_state_dict(ckpt["model"]))       # strict: the 288 reference keys
    return model.to(device).eval(), cfg


def save_packed(path, state_dict, cfg=None):
    """BN-folded fp32 blob + CRC32 + config -> one flat file: MAGIC | u32 header bytes | header JSON | blob."""
    blob = pack_state_dict(clean_state_dict(state_dict)).numpy()
    raw = blob.astype("<f4").tobytes()
    header = json.dumps({"floats": int(blob.size), "crc32": zlib.crc32(raw), "config": cfg or {}, "layers": len(arch.LAYERS)}).encode()
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<I", len(header)) + header + raw)
    return zlib.crc32(raw)


def save_raw_state_dict(path, state_dict):
    """The reference state_dict as a flat list of named fp32 tensors, for hosts without Python / torch (examples/c_abi_demo.c
    feeds it to ``cv_square_pack_weights``): RAW_MAGIC | u32 count | per tensor: u16 name bytes | name | i64 numel | fp32 data.
    Integer bookkeeping tensors (num_batches_tracked, class_to_*) are not written: the path never reads them."""
    sd = clean_state_dict(state_dict)
    items = [(k, v.detach().to("cpu", torch.float32).contiguous().numpy()) for k, v in sd.items()
             if not (k.endswith("num_batches_tracked") or k.startswith("class_to_"))]
    with open(path, "wb") as f:
        f.write(RAW_MAGIC + struct.pack("<I", len(items)))
        for k, a in items:
            kb = k.encode()
            f.write(struct.pack("<H", len(kb)) + kb + struct.pack("<q", a.size) + a.astype("<f4").tobytes())
    return len(items)


def load_packed(path):
    """-> (blob fp32 CPU tensor of arch.BLOB_FLOATS, config).  Raises ValueError on a truncated / corrupted / foreign file."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != MAGIC or len(data) < 12:
        raise ValueError(f"{path}: not a chess_vision_b200 packed-weights file")
    n = struct.unpack("<I", data[8:12])[0]
    header = json.loads(data[12:12 + n].decode())
    raw = data[12 + n:]
    if header["floats"] != arch.BLOB_FLOATS or len(raw) != 4 * arch.BLOB_FLOATS:
        raise ValueError(f"{path}: blob has {len(raw) // 4} floats, this build expects {arch.BLOB_FLOATS}")
    if zlib.crc32(raw) != header["crc32"]:
        raise ValueError(f"{path}: CRC mismatch (file corrupted)")
    return torch.from_numpy(np.frombuffer(raw, dtype="<f4").copy()), header["config"]


def model_from_packed(path, device="cuda"):
    """Packed file -> serving model: the blob goes straight to the device (no state_dict, no BN fold)."""
    blob, cfg = load_packed(path)
    cfg = cfg or {"model": {"arch": "square"}}
    cfg.setdefault("model", {})["pretrained"] = False
    model = build_model(cfg).to(device).eval()
    model.load_packed_blob(blob.to(device))
    return model, cfg
