"""CPU restatement of the reference's evaluation bookkeeping (TEST INFRASTRUCTURE -- only tests/, smoke() and bench.py's CPU
legs may import this; the product path is chess_vision_b200/csrc/eval.cu).

Follows /root/reference/evaluate.py:48-155 (`evaluate`): per-square argmax (first maximum wins, torch.argmax), exact integer
counts, the 13x13 confusion matrix (rows = true class), per-piece totals, turn / castling statistics over LEGAL positions only
(evaluate.py:101-122), the full-FEN count (evaluate.py:121-122) and the cross-entropy loss (evaluate.py:95-96).
Pinned: tests/golden/eval_reference.json holds the outputs of the reference's own `evaluate()` run on seeded synthetic
logits (oracle/make_golden_eval.py), and tests/test_eval_cpu.py checks this restatement against them.
"""
import numpy as np

NUM_CLASSES, NUM_SQUARES = 13, 64          # dataset.py:21-22

# counter layout shared with include/chessvision_b200.h (CV_EVAL_*)
TOTAL_BOARDS, TOTAL_SQUARES, CORRECT_SQUARES, CORRECT_BOARDS, TOTAL_LEGAL, CORRECT_TURN = 0, 1, 2, 3, 4, 5
CORRECT_CASTLING_RIGHT, CORRECT_CASTLING_ALL, CORRECT_FULL_FEN = 6, 10, 11
PIECE_CORRECT, PIECE_TOTAL, CONFUSION, TURN_CONFUSION, N_COUNTERS = 12, 25, 38, 207, 211


def synth_eval_batch(seed, n):
    """Seeded synthetic logits + labels of one batch (the same recipe feeds the reference run and every test)."""
    rng = np.random.default_rng(seed)
    squares = (rng.standard_normal((n, NUM_SQUARES * NUM_CLASSES)) * 3).astype(np.float32)
    turn = rng.standard_normal((n, 1)).astype(np.float32)
    castling = rng.standard_normal((n, 4)).astype(np.float32)
    pred = squares.reshape(n, NUM_SQUARES, NUM_CLASSES).argmax(-1)
    rnd = rng.integers(0, NUM_CLASSES, size=(n, NUM_SQUARES))
    sq_labels = np.where(rng.random((n, NUM_SQUARES)) < 0.9, pred, rnd).astype(np.int64)
    sq_labels[::3] = pred[::3]                                            # some boards fully correct
    turn_labels = np.where(rng.random((n, 1)) < 0.8, turn > 0, rng.random((n, 1)) < 0.5).astype(np.float32)
    castling_labels = np.where(rng.random((n, 4)) < 0.85, castling > 0, rng.random((n, 4)) < 0.5).astype(np.float32)
    legal = (rng.random((n, 1)) < 0.75).astype(np.float32)
    # exact ties between classes: torch.argmax takes the first maximum (evaluate.py:87)
    squares[0, 0:13] = 0.0
    squares[n - 1, 13 * 5 + 3] = squares[n - 1, 13 * 5 + 9] = 50.0
    return {"squares": squares, "turn": turn, "castling": castling, "sq_labels": sq_labels, "turn_labels": turn_labels,
            "castling_labels": castling_labels, "legal": legal}


def evaluate_batch(batch):
    """One batch of evaluate.py:74-155 -> (counters int64[N_COUNTERS], per_sample uint8 (n,4), board_loss float64 (n,))."""
    sq = batch["squares"].reshape(-1, NUM_SQUARES, NUM_CLASSES).astype(np.float32)
    n = sq.shape[0]
    labels = batch["sq_labels"].astype(np.int64)
    c = np.zeros(N_COUNTERS, dtype=np.int64)
    preds = sq.argmax(-1)                                                 # evaluate.py:87 (first maximum)
    ok = preds == labels                                                  # :88
    board_ok = ok.all(1)                                                  # :90
    c[TOTAL_BOARDS] = n
    c[TOTAL_SQUARES] = labels.size                                        # :92
    c[CORRECT_SQUARES] = ok.sum()                                         # :89
    c[CORRECT_BOARDS] = board_ok.sum()                                    # :91
    x = sq.astype(np.float64)
    lse = np.log(np.exp(x - x.max(-1, keepdims=True)).sum(-1)) + x.max(-1)
    ce = lse - np.take_along_axis(x, labels[..., None], -1)[..., 0]       # :95 CrossEntropyLoss per square
    board_loss = ce.sum(1)
    turn_pred = (batch["turn"] > 0).reshape(n)                            # :99
    turn_true = batch["turn_labels"].reshape(n) > 0.5
    turn_ok = turn_pred == turn_true                                      # :100
    cast_ok = (batch["castling"] > 0) == (batch["castling_labels"] > 0.5)  # :101-102
    cast_all = cast_ok.all(1)                                             # :103
    legal = batch["legal"].reshape(n) > 0
    c[TOTAL_LEGAL] = legal.sum()                                          # :107
    c[CORRECT_TURN] = (turn_ok & legal).sum()                             # :108
    for j in range(n):                                                    # :113-115
        if legal[j]:
            c[TURN_CONFUSION + 2 * int(turn_true[j]) + int(turn_pred[j])] += 1
    for r in range(4):                                                    # :117-118
        c[CORRECT_CASTLING_RIGHT + r] = (cast_ok[:, r] & legal).sum()
    c[CORRECT_CASTLING_ALL] = (cast_all & legal).sum()                    # :119
    c[CORRECT_FULL_FEN] = (board_ok & turn_ok & cast_all & legal).sum()   # :121-122
    for k in range(NUM_CLASSES):                                          # :127-130
        m = labels == k
        c[PIECE_TOTAL + k] = m.sum()
        c[PIECE_CORRECT + k] = (preds[m] == k).sum()
    np.add.at(c, CONFUSION + labels.reshape(-1) * NUM_CLASSES + preds.reshape(-1), 1)   # :132-133
    per = np.zeros((n, 4), dtype=np.uint8)                                # :136-146 sample_results
    per[:, 0] = (~ok).sum(1)
    per[:, 1] = board_ok
    per[:, 2] = np.where(legal, turn_ok, 255)
    per[:, 3] = np.where(legal, cast_all, 255)
    return c, per, board_loss


def summary(counters, loss_sum):
    """The dict evaluate.py:222-231 returns."""
    c = counters
    tl = max(int(c[TOTAL_LEGAL]), 1)
    return {"loss": float(loss_sum) / (NUM_SQUARES * int(c[TOTAL_BOARDS])), "square_acc": int(c[CORRECT_SQUARES]) / int(c[TOTAL_SQUARES]),
            "board_acc": int(c[CORRECT_BOARDS]) / int(c[TOTAL_BOARDS]), "turn_acc": int(c[CORRECT_TURN]) / tl,
            "castling_acc": int(c[CORRECT_CASTLING_ALL]) / tl, "full_fen_acc": int(c[CORRECT_FULL_FEN]) / tl,
            "total_boards": int(c[TOTAL_BOARDS]), "total_legal": int(c[TOTAL_LEGAL])}


# ---- grouped metrics (evaluate.py:233-287): seeded manifest rows + per-sample results ---------------------------------------
def synth_manifest(seed=5, n=97):
    """Manifest rows (strings, as csv.DictReader yields them; the fields evaluate.py:239-246 groups by) + the per-sample table (N,4) uint8 of
    EvalAccumulator, seeded: inputs of the grouped-metrics golden (oracle/make_golden_eval_grouped.py, tests/test_eval_cpu.py)."""
    rng = np.random.default_rng(seed)
    styles = ["alpha", "cburnett", "merida", "wood"]
    meta = [{"piece_count": str(int(rng.integers(2, 33))), "castling": ["-", "KQkq", "Kq", "k"][int(rng.integers(0, 4))],
             "turn": "wb"[int(rng.integers(0, 2))], "has_highlight": str(int(rng.integers(0, 2))), "style": styles[int(rng.integers(0, 4))],
             "flipped": str(int(rng.integers(0, 2)))} for _ in range(n)]
    legal = rng.random(n) < 0.8
    per = np.zeros((n, 4), np.uint8)
    per[:, 0] = rng.integers(0, 4, n) * (rng.random(n) < 0.4)
    per[:, 1] = per[:, 0] == 0
    per[:, 2] = np.where(legal, rng.random(n) < 0.9, 255)
    per[:, 3] = np.where(legal, rng.random(n) < 0.7, 255)
    return meta, per


class ManifestStub:
    use_manifest = True

    def __init__(self, meta):
        self.meta = meta

    def get_metadata(self, i):
        return self.meta[i]
