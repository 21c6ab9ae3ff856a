"""Evaluation with the bookkeeping on the GPU: mirrors /root/reference/evaluate.py:47-231 (`evaluate`).

The reference turns every batch into Python scalars on the host (about twenty `.item()` / `.cpu()` synchronisations and an
O(B*64) loop for the confusion matrix, evaluate.py:87-155).  Here one kernel (`cv_eval_accumulate`, csrc/eval.cu) adds a batch
to exact int64 counters that stay on the device; the host reads them once at the end.  Same names, same report, same summary
dict; the grouped-by-manifest tables (evaluate.py:233-287) are host dictionary work over the per-sample table the kernel emits
(`grouped_metrics` / `grouped_report`).
"""
import numpy as np
import torch

from . import _native
from .dataset import NUM_CLASSES, NUM_SQUARES, labels_to_fen

# counter layout = enum CV_EVAL_* of include/chessvision_b200.h
TOTAL_BOARDS, TOTAL_SQUARES, CORRECT_SQUARES, CORRECT_BOARDS, TOTAL_LEGAL, CORRECT_TURN = 0, 1, 2, 3, 4, 5
CORRECT_CASTLING_RIGHT, CORRECT_CASTLING_ALL, CORRECT_FULL_FEN = 6, 10, 11
PIECE_CORRECT, PIECE_TOTAL, CONFUSION, TURN_CONFUSION, N_COUNTERS = 12, 25, 38, 207, 211
PIECE_NAMES = {0: "empty", 1: "P", 2: "N", 3: "B", 4: "R", 5: "Q", 6: "K", 7: "p", 8: "n", 9: "b", 10: "r", 11: "q", 12: "k"}


def _u8(t, device, shape):
    """Labels arrive as the reference's dataset yields them (int64 classes, float 0/1 flags): compact device uint8."""
    t = torch.as_tensor(t)
    if t.dtype.is_floating_point:
        t = t > 0.5
    return t.to(device=device, non_blocking=True).to(torch.uint8).reshape(shape).contiguous()


class EvalAccumulator:
    """Device-resident state of one evaluation run (the local variables of evaluate.py:52-70)."""

    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _native.NativeError("EvalAccumulator needs a CUDA device (chess_vision_b200 has no CPU fallback)")
        self.counters = torch.zeros(N_COUNTERS, dtype=torch.int64, device=self.device)
        self._per_sample, self._loss, self._preds = [], [], []

    @torch.no_grad()
    def update(self, outputs, labels, keep_predictions=False):
        """outputs: the model's dict (squares (B,832), turn (B,1), castling (B,4)); labels: dict with squares (B,64),
        turn (B,1), castling (B,4), legal (B,1) as in evaluate.py:75-80."""
        sq = outputs["squares"].to(self.device, torch.float32).contiguous()
        B = sq.shape[0]
        turn = outputs["turn"].to(self.device, torch.float32).reshape(B).contiguous()
        cast = outputs["castling"].to(self.device, torch.float32).reshape(B, 4).contiguous()
        lab = _u8(labels["squares"], self.device, (B, NUM_SQUARES))
        tl, cl, lg = _u8(labels["turn"], self.device, (B,)), _u8(labels["castling"], self.device, (B, 4)), _u8(labels["legal"], self.device, (B,))
        per = torch.empty((B, 4), dtype=torch.uint8, device=self.device)
        loss = torch.empty(B, dtype=torch.float32, device=self.device)
        p = _native.ptr
        with torch.cuda.device(self.device):                  # the launch must target the counters' device, not the caller's current one
            _native.check(_native.lib().cv_eval_accumulate(p(sq), p(turn), p(cast), p(lab), p(tl), p(cl), p(lg), B, p(self.counters), p(per),
                                                           p(loss), _native.stream_ptr(self.device)))
        self._per_sample.append(per)
        self._loss.append(loss)
        if keep_predictions:                                  # only for the "worst predictions" list (evaluate.py:147-153)
            self._preds.append((sq.view(B, NUM_SQUARES, NUM_CLASSES).argmax(-1).to(torch.uint8), lab))

    def results(self):
        """One device->host read: counters (int64 numpy), per-sample table (uint8 (N,4)), per-board loss sums."""
        c = self.counters.cpu().numpy()
        per = torch.cat(self._per_sample).cpu().numpy() if self._per_sample else np.zeros((0, 4), np.uint8)
        loss = torch.cat(self._loss).cpu().numpy().astype(np.float64) if self._loss else np.zeros(0)
        return c, per, loss

    def summary(self):
        """The dict evaluate.py:222-231 returns."""
        c, _, loss = self.results()
        boards, legal = int(c[TOTAL_BOARDS]), int(c[TOTAL_LEGAL])
        return {"loss": float(loss.sum()) / (NUM_SQUARES * max(boards, 1)), "square_acc": int(c[CORRECT_SQUARES]) / max(int(c[TOTAL_SQUARES]), 1),
                "board_acc": int(c[CORRECT_BOARDS]) / max(boards, 1), "turn_acc": int(c[CORRECT_TURN]) / max(legal, 1),
                "castling_acc": int(c[CORRECT_CASTLING_ALL]) / max(legal, 1), "full_fen_acc": int(c[CORRECT_FULL_FEN]) / max(legal, 1),
                "total_boards": boards, "total_legal": legal}

    def report(self):
        """The text evaluate.py:157-216 prints (same lines, same formats)."""
        c, per, loss = self.results()
        s = self.summary()
        tb, ts, tl = s["total_boards"], int(c[TOTAL_SQUARES]), s["total_legal"]
        L = ["", "=" * 60, "EVALUATION RESULTS", "=" * 60, "", f"Overall ({tb} images, {tl} legal):", f"  Loss:            {s['loss']:.4f}",
             f"  Per-square acc:  {s['square_acc']:.4f} ({int(c[CORRECT_SQUARES])}/{ts})",
             f"  Full-board acc:  {s['board_acc']:.4f} ({int(c[CORRECT_BOARDS])}/{tb})"]
        if tl > 0:
            t = c[TURN_CONFUSION:TURN_CONFUSION + 4]
            L += ["", "Turn prediction (legal positions only):", f"  Accuracy:        {s['turn_acc']:.4f} ({int(c[CORRECT_TURN])}/{tl})",
                  "  Confusion (rows=true, cols=pred):", "             White  Black", f"    White  {int(t[0]):>6d} {int(t[1]):>6d}",
                  f"    Black  {int(t[2]):>6d} {int(t[3]):>6d}", "", "Castling prediction (legal positions only):"]
            for r, name in enumerate(["K", "Q", "k", "q"]):
                n = int(c[CORRECT_CASTLING_RIGHT + r])
                L.append(f"  {name:>1s}: {n / tl:.4f} ({n}/{tl})")
            L += [f"  All-4-correct:   {s['castling_acc']:.4f} ({int(c[CORRECT_CASTLING_ALL])}/{tl})", "",
                  "Full FEN accuracy (position + turn + castling, legal only):", f"  {s['full_fen_acc']:.4f} ({int(c[CORRECT_FULL_FEN])}/{tl})"]
        else:
            L += ["", "No legal positions in dataset — turn/castling metrics skipped."]
        L += ["", "Per-piece accuracy:"]
        for k in range(NUM_CLASSES):
            tot, ok = int(c[PIECE_TOTAL + k]), int(c[PIECE_CORRECT + k])
            if tot > 0:
                L.append(f"  {PIECE_NAMES[k]:>5s}: {ok / tot:.4f}  ({ok}/{tot})")
        L += ["", "Confusion matrix (rows=true, cols=predicted):", "       " + "".join(f"{PIECE_NAMES[k]:>6s}" for k in range(NUM_CLASSES))]
        conf = c[CONFUSION:CONFUSION + 169].reshape(13, 13)
        for t in range(NUM_CLASSES):
            L.append(f"  {PIECE_NAMES[t]:>4s} " + "".join(f"{int(conf[t, p]):>6d}" for p in range(NUM_CLASSES)))
        if self._preds:
            preds = torch.cat([a for a, _ in self._preds]).cpu()
            labs = torch.cat([b for _, b in self._preds]).cpu()
            worst = sorted(((int(per[i, 0]), i) for i in range(len(per)) if per[i, 0] > 0), key=lambda x: -x[0])    # stable, as list.sort
            L += ["", "Top 10 worst predictions:"]
            for num_wrong, i in worst[:10]:
                L += [f"  Image {i}: {num_wrong}/64 squares wrong", f"    True: {labels_to_fen(labs[i])}", f"    Pred: {labels_to_fen(preds[i])}"]
        return "\n".join(L)


# ---- grouped metrics (evaluate.py:30-45, 233-287): accuracy by manifest field, from the per-sample table ---------------------------------
def piece_count_bucket(count):
    count = int(count)
    return "endgame (2-10)" if count <= 10 else "midgame (11-20)" if count <= 20 else "opening (21-32)"


def castling_category(castling_str):
    return "none" if castling_str == "-" else "has_rights"


GROUPING_FIELDS = {                                           # field -> bucket function, in the reference's order (evaluate.py:239-246)
    "piece_count": piece_count_bucket,
    "castling": castling_category,
    "turn": lambda x: "white" if x == "w" else "black",
    "has_highlight": lambda x: "highlighted" if x == "1" else "no highlight",
    "style": lambda x: x,
    "flipped": lambda x: "flipped" if x == "1" else "normal",
}


def grouped_metrics(dataset, per_sample):
    """per_sample: the (N,4) uint8 table of ``EvalAccumulator.results()`` (squares wrong, board correct, turn correct, castling-all
    correct; 255 = not a legal position, the reference's None) in dataset order -> {field: {bucket: counts}} exactly as
    evaluate.py:253-272 accumulates them.  Needs ``dataset.use_manifest`` and ``dataset.get_metadata(i)`` like the reference."""
    if not getattr(dataset, "use_manifest", False):
        return {}
    per = np.asarray(per_sample)
    out = {}
    first = dataset.get_metadata(0)
    for field, bucket_fn in GROUPING_FIELDS.items():
        if field not in first:
            continue
        groups = {}
        for i in range(per.shape[0]):
            g = groups.setdefault(bucket_fn(dataset.get_metadata(i).get(field, "")),
                                  {"total": 0, "board_correct": 0, "turn_correct": 0, "turn_total": 0, "castling_correct": 0, "castling_total": 0})
            g["total"] += 1
            g["board_correct"] += int(per[i, 1])
            if per[i, 2] != 255:
                g["turn_total"] += 1
                g["turn_correct"] += int(per[i, 2])
            if per[i, 3] != 255:
                g["castling_total"] += 1
                g["castling_correct"] += int(per[i, 3])
        out[field] = groups
    return out


def grouped_report(dataset, per_sample):
    """The text evaluate.py:248-287 prints (same lines, same formats); empty string without a manifest."""
    groups_by_field = grouped_metrics(dataset, per_sample)
    if not getattr(dataset, "use_manifest", False):
        return ""
    L = ["", "=" * 60, "GROUPED METRICS", "=" * 60]
    for field, groups in groups_by_field.items():
        L += ["", f"By {field}:"]
        for bucket in sorted(groups):
            g = groups[bucket]
            line = f"  {bucket:>20s}: board_acc={(g['board_correct'] / g['total'] if g['total'] > 0 else 0):.4f} (n={g['total']})"
            if g["turn_total"] > 0:
                line += f"  turn={g['turn_correct'] / g['turn_total']:.4f}"
            if g["castling_total"] > 0:
                line += f"  castling={g['castling_correct'] / g['castling_total']:.4f}"
            L.append(line)
    return "\n".join(L)


@torch.no_grad()
def evaluate(model, dataset, loader, device, verbose=True):
    """Drop-in for evaluate.py:47 `evaluate(model, dataset, loader, device)`: same arguments, same summary dict."""
    model.eval()
    acc = EvalAccumulator(device)
    for images, labels in loader:
        images = images.to(device, non_blocking=True)
        outputs = model.forward_u8(images) if images.dtype == torch.uint8 else model(images)
        acc.update(outputs, labels, keep_predictions=verbose)
    if verbose:
        print(acc.report())
        grouped = grouped_report(dataset, acc.results()[1])       # evaluate.py:218
        if grouped:
            print(grouped)
    return acc.summary()
