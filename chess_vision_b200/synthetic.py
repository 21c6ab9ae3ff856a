"""Synthetic boards and random-init weights (hard part H1).

Boards come from a counter-based integer hash keyed by ``(seed, global_board_index)`` so any
sharding of a board stream over ranks yields bit-identical boards.  The same arithmetic is
implemented on the device by ``cv_synth_boards`` (csrc/synth.cu); ``tests/test_gpu_synth.py``
checks the two agree byte for byte.

Distributions
  DIST_UNIFORM    i.i.d. bytes -- throughput runs only (parity-vacuous: every square looks alike)
  DIST_STRUCTURED per-square base colour averaged with a 32x32 blocky pattern -- parity runs
"""
import numpy as np
import torch

from . import arch

DIST_UNIFORM = 0
DIST_STRUCTURED = 1
LAYOUT_HWC = 0
LAYOUT_CHW = 1

_U32 = np.uint32


def hash32(x):
    """lowbias32 integer finaliser on uint32 arrays (wrap-around arithmetic)."""
    x = np.asarray(x, dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        x ^= x >> _U32(16)
        x *= _U32(0x7FEB352D)
        x ^= x >> _U32(15)
        x *= _U32(0x846CA68B)
        x ^= x >> _U32(16)
    return x


def board_key(seed: int, board_index):
    """Per-board 32-bit key; board_index may be an array of global indices."""
    idx = (np.asarray(board_index, dtype=np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    with np.errstate(over="ignore"):
        return hash32(_U32(seed & 0xFFFFFFFF) ^ hash32(idx + _U32(0x9E3779B9)))


def synth_boards(first_board: int, n: int, size: int = 256, seed: int = 1,
                 dist: int = DIST_STRUCTURED, layout: int = LAYOUT_HWC) -> np.ndarray:
    """uint8 boards ``first_board .. first_board+n-1``: (n,size,size,3) or (n,3,size,size)."""
    assert size % 32 == 0
    keys = board_key(seed, np.arange(first_board, first_board + n))[:, None, None, None]   # (n,1,1,1)
    c = np.arange(3, dtype=np.uint32)[None, None, None, :]
    yy = np.arange(size, dtype=np.uint32)[None, :, None, None]
    xx = np.arange(size, dtype=np.uint32)[None, None, :, None]
    with np.errstate(over="ignore"):
        if dist == DIST_UNIFORM:
            lin = (yy * _U32(size) + xx) * _U32(3) + c
            img = (hash32(keys + _U32(5000) + lin) & _U32(0xFF)).astype(np.uint8)
        else:
            sq, cell = size // 8, size // 32
            sq_id = (yy // _U32(sq)) * _U32(8) + (xx // _U32(sq))
            base = hash32(keys + _U32(1) + sq_id * _U32(3) + c) & _U32(0xFF)
            cell_id = (yy // _U32(cell)) * _U32(32) + (xx // _U32(cell))
            pat = (hash32(keys + _U32(1000) + cell_id * _U32(3) + c) >> _U32(8)) & _U32(0xFF)
            img = ((base + pat + _U32(1)) >> _U32(1)).astype(np.uint8)
    if layout == LAYOUT_CHW:
        img = np.ascontiguousarray(img.transpose(0, 3, 1, 2))
    return img


def synth_flipped(first_board: int, n: int, seed: int = 1) -> np.ndarray:
    """Bernoulli(0.5) 'rendered from Black's side' flag per board (BASELINE.json config 4)."""
    keys = board_key(seed, np.arange(first_board, first_board + n))
    with np.errstate(over="ignore"):
        return (hash32(keys + _U32(7777)) & _U32(1)).astype(np.uint8)


def normalize_boards(u8_hwc: np.ndarray) -> torch.Tensor:
    """uint8 (n,H,H,3) -> fp32 (n,3,H,H), the arithmetic of ToTensor + Normalize
    (dataset.py:177-181): ``(u8 / 255 - mean) / std`` in fp32."""
    from .dataset import NORM_MEAN, NORM_STD
    t = torch.from_numpy(u8_hwc).permute(0, 3, 1, 2).to(torch.float32).div(255)
    mean = torch.tensor(NORM_MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(NORM_STD, dtype=torch.float32).view(1, 3, 1, 1)
    return t.sub(mean).div(std).contiguous()


# -------------------------------------------------------------------------------------------------
# Random-init weights (no pretrained weights are obtainable offline).  Recipe H1: conv
# N(0, sqrt(2/fan_out)); BatchNorm statistics perturbed so that folding is actually exercised.
# -------------------------------------------------------------------------------------------------
def init_state_dict(template: dict, seed: int = 0) -> dict:
    """Fill a state_dict with the reference's 288 keys from ``numpy.random.default_rng(seed)``."""
    rng = np.random.default_rng(seed)
    out = {}
    for key, ref in template.items():
        shape = tuple(ref.shape)
        if key in ("class_to_type", "class_to_color"):
            out[key] = ref.clone()
        elif key.endswith("num_batches_tracked"):
            out[key] = torch.zeros(shape, dtype=torch.long)
        elif key.endswith("running_mean"):
            out[key] = torch.from_numpy(rng.normal(0.0, 0.1, shape).astype(np.float32))
        elif key.endswith("running_var"):
            out[key] = torch.from_numpy(rng.uniform(0.8, 1.2, shape).astype(np.float32))
        elif len(shape) == 4:                                   # conv weight (O, I/g, kh, kw)
            fan_out = shape[0] * shape[2] * shape[3]
            if shape[1] == 1 and shape[0] > 1 and "conv_stem" not in key:
                fan_out = shape[2] * shape[3]                   # depthwise: groups == out
            out[key] = torch.from_numpy(rng.normal(0.0, (2.0 / fan_out) ** 0.5, shape).astype(np.float32))
        elif ".bn" in key or "norm_head" in key:                # BN affine
            if key.endswith("weight"):
                out[key] = torch.from_numpy(rng.uniform(0.8, 1.2, shape).astype(np.float32))
            else:
                out[key] = torch.from_numpy(rng.normal(0.0, 0.1, shape).astype(np.float32))
        elif len(shape) == 2:                                   # Linear weight
            out[key] = torch.from_numpy((rng.standard_normal(shape) / np.sqrt(shape[1])).astype(np.float32))
        else:                                                   # Linear bias
            out[key] = torch.from_numpy(rng.normal(0.0, 0.05, shape).astype(np.float32))
    return out


def calibration_stats(features: torch.Tensor, seed: int = 999) -> dict:
    """Small summary of a calibration batch's pooled trunk features (n_boards*64, 480) from which
    ``calibrate_heads`` derives the five Linear heads.  Stored in ``tests/golden/`` so every machine
    rebuilds bit-identical head weights without re-running the calibration forward."""
    f = features.detach().to(torch.float64).cpu().numpy()
    n_boards = f.shape[0] // arch.NUM_SQUARES
    g = f.reshape(n_boards, -1)
    stats = {"f_mean": f.mean(0), "f_std": np.float64(f.std()),
             "g_mean": g.mean(0), "g_std": np.float64(g.std())}
    w, b = _draw_global(np.random.default_rng(seed + 1), stats)
    h = np.maximum(g @ w.T + b, 0.0)
    stats["h_mean"] = h.mean(0)
    stats["h_std"] = np.float64(h.std())
    return stats


def _lin(rng, rows, mean, std, gain):
    w = rng.standard_normal((rows, mean.shape[0])) * gain / (float(std) * np.sqrt(mean.shape[0]) + 1e-12)
    return w, -(w @ mean)


def _draw_global(rng, stats):
    w, b = _lin(rng, arch.GLOBAL_HIDDEN, stats["g_mean"], stats["g_std"], 2.0)
    return w, b + 0.5


def calibrate_heads(state: dict, stats: dict, seed: int = 999) -> dict:
    """Re-draw the five Linear heads so predictions are non-degenerate (H1c): all 13 classes,
    digit runs and mixed turn/castling bits appear.  Both sides of a parity test load the SAME
    resulting tensors, so calibration cannot hide a mismatch."""
    out = dict(state)

    def put(name, w, b):
        out[name + ".weight"] = torch.from_numpy(w.astype(np.float32))
        out[name + ".bias"] = torch.from_numpy(b.astype(np.float32))
    rng = np.random.default_rng(seed)
    put("type_head.1", *_lin(rng, 7, stats["f_mean"], stats["f_std"], 4.0))
    put("color_head.1", *_lin(rng, 3, stats["f_mean"], stats["f_std"], 4.0))
    put("global_head.1", *_draw_global(np.random.default_rng(seed + 1), stats))
    rng2 = np.random.default_rng(seed + 2)
    put("turn_head", *_lin(rng2, 1, stats["h_mean"], stats["h_std"], 2.0))
    put("castling_head", *_lin(rng2, 4, stats["h_mean"], stats["h_std"], 2.0))
    return out
