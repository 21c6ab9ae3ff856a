"""the other architectures of `build_model` (ChessCNN, models/cnn.py:36-53; ChessViT, models/vit.py:28-49) return the
same {"squares" (B,832), "turn" (B,1), "castling" (B,4)} contract as the square model, built from type (7) + color (3) logits on an
8x8 grid.  Their backbones are out of scope, but everything AFTER the backbone is this package's: the device type+color combine,
the device FEN encoder and the on-device evaluation bookkeeping.  Here the reference's CNN head is restated on random stride-32
features and driven through those three entry points, against the CPU oracles."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from chess_vision_b200.evaluate import EvalAccumulator
from chess_vision_b200.models.common import combine_type_color
from chess_vision_b200.predict import fen_from_outputs
from oracle import eval_oracle as eo
from oracle import square_oracle as oracle

pytestmark = pytest.mark.gpu


def cnn_head_outputs(B, C=96, seed=0):
    """models/cnn.py:36-53 on random features (B, C, 8, 8): 1x1-conv type / color heads -> permute -> combine; pooled turn / castling."""
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, C, 8, 8, generator=g)
    type_head, color_head = nn.Conv2d(C, 7, 1), nn.Conv2d(C, 3, 1)
    turn_head, castling_head = nn.Linear(C, 1), nn.Linear(C, 4)
    with torch.no_grad():
        for m in (type_head, color_head, turn_head, castling_head):
            m.weight.copy_(torch.randn(m.weight.shape, generator=g) * 0.5)
            m.bias.copy_(torch.randn(m.bias.shape, generator=g))
        spatial = F.adaptive_avg_pool2d(feats, (8, 8))
        t = type_head(spatial).permute(0, 2, 3, 1).contiguous()            # (B, 8, 8, 7)
        c = color_head(spatial).permute(0, 2, 3, 1).contiguous()           # (B, 8, 8, 3)
        pooled = feats.mean((2, 3))
        return t, c, turn_head(pooled), castling_head(pooled)


@pytest.mark.parametrize("B", [1, 5, 130])
def test_cnn_shaped_heads_through_combine_fen_and_eval(B):
    t, c, turn, cast = cnn_head_outputs(B)
    joint = combine_type_color(t.cuda(), c.cuda(), None, None)               # device op (cv_combine_type_color), (B, 8, 8, 13)
    want = oracle.combine_type_color(t, c)
    assert joint.shape == (B, 8, 8, 13) and torch.equal(joint.cpu(), want)
    outputs = {"squares": joint.reshape(B, -1), "turn": turn.cuda(), "castling": cast.cuda()}    # cnn.py:49-53
    sq = want.reshape(B, 64, 13).numpy()
    assert fen_from_outputs(outputs) == oracle.fen_strings(sq, turn.numpy(), cast.numpy())
    fl = (np.arange(B) % 2).astype(np.uint8)
    assert fen_from_outputs(outputs, flipped=torch.from_numpy(fl)) == oracle.fen_strings(sq, turn.numpy(), cast.numpy(), flipped=fl)
    # evaluation bookkeeping on these logits against labels that are right 85 % of the time
    rng = np.random.default_rng(B)
    pred = sq.argmax(-1)
    labels = {"squares": torch.from_numpy(np.where(rng.random((B, 64)) < 0.85, pred, rng.integers(0, 13, (B, 64))).astype(np.int64)),
              "turn": torch.from_numpy((rng.random((B, 1)) < 0.5).astype(np.float32)), "castling": torch.from_numpy((rng.random((B, 4)) < 0.5).astype(np.float32)),
              "legal": torch.from_numpy((rng.random((B, 1)) < 0.7).astype(np.float32))}
    acc = EvalAccumulator("cuda")
    acc.update(outputs, labels)
    batch = {"squares": want.reshape(B, 832).numpy(), "turn": turn.numpy(), "castling": cast.numpy(), "sq_labels": labels["squares"].numpy(),
             "turn_labels": labels["turn"].numpy(), "castling_labels": labels["castling"].numpy(), "legal": labels["legal"].numpy()}
    c_ref, per_ref, _ = eo.evaluate_batch(batch)
    c_got, per_got, _ = acc.results()
    assert np.array_equal(c_got, c_ref) and np.array_equal(per_got, per_ref)
