"""FEN codec and class tables of the hot path (host side).

Mirrors the names and behaviour of the reference's ``dataset.py:14-116`` (class tables,
``fen_to_labels``, ``labels_to_fen``, ``filename_to_fen``, ``parse_full_fen``) so code written
against the reference keeps working.  The batched, on-device encoder is ``cv_square_fen``
(csrc/fen.cu); these host functions are the small-scale mirror used by ``predict`` callers and
by the parsers the tests need.  Dataset / augmentation classes are out of scope.
"""
import os

import torch

# dataset.py:14-19 -- 13 joint classes: empty, 6 white, 6 black
_PIECES = ".PNBRQKpnbrqk"
PIECE_TO_INDEX = {ch: i for i, ch in enumerate(_PIECES)}
INDEX_TO_PIECE = dict(enumerate(_PIECES))
NUM_CLASSES = 13
NUM_SQUARES = 64

# dataset.py:26-32 -- type(7) x color(3) decomposition
NUM_PIECE_TYPES = 7
NUM_PIECE_COLORS = 3
CLASS_TO_TYPE = [0] + [1, 2, 3, 4, 5, 6] * 2
CLASS_TO_COLOR = [0] + [1] * 6 + [2] * 6

# get_transform eval branch (dataset.py:157-160, 177-181): timm pretrained_cfg of the trunk
NORM_MEAN = (0.485, 0.456, 0.406)
NORM_STD = (0.229, 0.224, 0.225)


def fen_to_labels(fen: str) -> torch.Tensor:
    """Placement field -> (64,) int64 class indices, index 0 = a8 ... 63 = h1 (dataset.py:35-49).

    Raises ``KeyError`` on an unknown piece letter and ``AssertionError`` when the ranks do not
    add up to 64 squares, like the reference.
    """
    cells = []
    for rank in fen.split("/"):
        for ch in rank:
            if ch.isdigit():
                cells += [0] * int(ch)
            else:
                cells.append(PIECE_TO_INDEX[ch])
    assert len(cells) == 64, f"Expected 64 squares, got {len(cells)} from FEN: {fen}"
    return torch.tensor(cells, dtype=torch.long)


def _encode_rank(classes) -> str:
    out, run = [], 0
    for c in classes:
        if c == 0:
            run += 1
            continue
        if run:
            out.append(str(run))
            run = 0
        out.append(_PIECES[c])
    if run:
        out.append(str(run))
    return "".join(out)


def labels_to_fen(labels) -> str:
    """(64,) class indices -> placement field with run-length digits (dataset.py:52-70)."""
    flat = [int(v) for v in (labels.tolist() if hasattr(labels, "tolist") else labels)]
    return "/".join(_encode_rank(flat[r:r + 8]) for r in range(0, 64, 8))


def filename_to_fen(filename: str) -> str:
    """'1B1B1K2-3p1N2-...-1B6.jpeg' -> '1B1B1K2/3p1N2/.../1B6' (dataset.py:73-76)."""
    return os.path.splitext(filename)[0].replace("-", "/")


def parse_full_fen(fen_str: str) -> dict:
    """Full FEN (2-6 fields) -> squares/turn/castling targets (dataset.py:79-116)."""
    fields = fen_str.strip().split()
    turn = fields[1] if len(fields) > 1 else "w"
    rights = fields[2] if len(fields) > 2 else "-"
    flags = [0.0 if rights == "-" else float(ch in rights) for ch in "KQkq"]
    return {
        "squares": fen_to_labels(fields[0]),
        "turn": torch.tensor([1.0 if turn == "b" else 0.0], dtype=torch.float),
        "castling": torch.tensor(flags, dtype=torch.float),
    }


def assemble_fen(placement: str, turn_logit: float, castling_logits) -> str:
    """Three-field FEN exactly as ``predict.py:31-42`` builds it."""
    turn = "b" if turn_logit > 0 else "w"
    rights = "".join(ch for v, ch in zip(castling_logits, "KQkq") if v > 0)
    return f"{placement} {turn} {rights or '-'}"
