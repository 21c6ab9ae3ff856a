"""On-device evaluation bookkeeping (cv_eval_accumulate through EvalAccumulator) against the CPU restatement and against the
report the REFERENCE's own evaluate() printed for the same seeded logits (tests/golden/eval_reference.json)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import eval_oracle as eo

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "eval_reference.json")))


def to_torch(b):
    outputs = {k: torch.from_numpy(b[k]).cuda() for k in ("squares", "turn", "castling")}
    labels = {"squares": torch.from_numpy(b["sq_labels"]), "turn": torch.from_numpy(b["turn_labels"]),
              "castling": torch.from_numpy(b["castling_labels"]), "legal": torch.from_numpy(b["legal"])}
    return outputs, labels


def test_counters_per_sample_and_loss_match_oracle_exactly():
    from chess_vision_b200.evaluate import EvalAccumulator
    acc = EvalAccumulator("cuda")
    c_ref = np.zeros(eo.N_COUNTERS, dtype=np.int64)
    per_ref, loss_ref = [], []
    for seed, n in GOLD["batches"] + [[7, 1], [8, 300], [9, 1025]]:          # ragged sizes incl. a single board and > one grid wave
        b = eo.synth_eval_batch(seed, n)
        acc.update(*to_torch(b))
        ci, pi, li = eo.evaluate_batch(b)
        c_ref += ci
        per_ref.append(pi)
        loss_ref.append(li)
    c, per, loss = acc.results()
    assert np.array_equal(c, c_ref)                                           # integer work: bit-exact
    assert np.array_equal(per, np.concatenate(per_ref))
    lr = np.concatenate(loss_ref)
    assert np.max(np.abs(loss - lr) / np.maximum(np.abs(lr), 1.0)) < 1e-5     # fp32 log-sum-exp vs fp64, tolerance 1e-5
    acc2 = EvalAccumulator("cuda")                                            # empty batch is a no-op
    acc2.update({"squares": torch.zeros((0, 832)), "turn": torch.zeros((0, 1)), "castling": torch.zeros((0, 4))},
                {"squares": torch.zeros((0, 64), dtype=torch.int64), "turn": torch.zeros((0, 1)), "castling": torch.zeros((0, 4)), "legal": torch.zeros((0, 1))})
    assert int(acc2.counters.sum()) == 0


def test_report_and_summary_equal_the_reference_run():
    from chess_vision_b200.evaluate import EvalAccumulator
    acc = EvalAccumulator("cuda")
    for seed, n in GOLD["batches"]:
        acc.update(*to_torch(eo.synth_eval_batch(seed, n)), keep_predictions=True)
    s, ref = acc.summary(), GOLD["summary"]
    for k in ref:
        if k == "loss":
            assert abs(s[k] - ref[k]) < 1e-5 * ref[k]
        else:
            assert s[k] == ref[k], k
    ours = [l.rstrip() for l in acc.report().splitlines()]
    theirs = [l.rstrip() for l in GOLD["report"].splitlines()]
    theirs = theirs[:len(ours)]                                               # the reference goes on with the manifest-grouped tables
    assert ours == theirs, [(a, b) for a, b in zip(ours, theirs) if a != b][:3]


def test_evaluate_drop_in_on_the_model(gpu_model):
    """evaluate(model, dataset, loader, device) end to end on synthetic boards: labels taken from the model's own fp32
    prediction, so every accuracy must be exactly 1."""
    from chess_vision_b200.evaluate import evaluate
    from chess_vision_b200 import synthetic
    u8 = synthetic.synth_boards(0, 24, 256, 1, synthetic.DIST_STRUCTURED)
    x = synthetic.normalize_boards(u8)
    loader = []
    for i in range(0, 24, 8):
        xb = x[i:i + 8]
        o = gpu_model(xb.cuda(), precision="fp32")
        labels = {"squares": o["squares"].view(-1, 64, 13).argmax(-1).cpu(), "turn": (o["turn"] > 0).float().cpu(),
                  "castling": (o["castling"] > 0).float().cpu(), "legal": torch.ones(xb.shape[0], 1)}
        loader.append((xb, labels))
    prev = gpu_model.precision
    gpu_model.precision = "fp32"
    try:
        s = evaluate(gpu_model, None, loader, torch.device("cuda"), verbose=False)
    finally:
        gpu_model.precision = prev
    assert s["total_boards"] == 24 and s["total_legal"] == 24
    assert s["square_acc"] == 1.0 and s["board_acc"] == 1.0 and s["turn_acc"] == 1.0 and s["castling_acc"] == 1.0 and s["full_fen_acc"] == 1.0
