"""Generate tests/golden/eval_reference.json by running the REFERENCE's own evaluate() (build container only).

/root/reference/evaluate.py is imported unmodified (over oracle/timm_shim, because dataset.py imports timm) and its
`evaluate(model, dataset, loader, device)` (evaluate.py:47-231) is called with a stub model that returns seeded synthetic
logits (oracle/eval_oracle.synth_eval_batch) and a list of (images, labels) batches.  The returned summary and the printed
report (confusion matrix, per-piece lines, turn confusion, castling lines) are stored.

    python oracle/make_golden_eval.py
"""
import contextlib
import io
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "timm_shim"))
sys.path.insert(0, REF)

import evaluate as ref_evaluate  # noqa: E402  /root/reference/evaluate.py
from oracle import eval_oracle  # noqa: E402

BATCHES = [(101, 7), (102, 16), (103, 5), (104, 33)]          # (seed, boards)


class StubModel:
    def __init__(self, batches):
        self.batches, self.i = batches, 0

    def eval(self):
        return self

    def __call__(self, images):
        b = self.batches[self.i]
        self.i += 1
        return {k: torch.from_numpy(b[k]) for k in ("squares", "turn", "castling")}


def main():
    batches = [eval_oracle.synth_eval_batch(s, n) for s, n in BATCHES]
    loader = [(torch.zeros(b["squares"].shape[0], 1),
               {"squares": torch.from_numpy(b["sq_labels"]), "turn": torch.from_numpy(b["turn_labels"]),
                "castling": torch.from_numpy(b["castling_labels"]), "legal": torch.from_numpy(b["legal"])}) for b in batches]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(io.StringIO()):
        summary = ref_evaluate.evaluate(StubModel(batches), object(), loader, torch.device("cpu"))
    out = {"batches": BATCHES, "summary": summary, "report": buf.getvalue(), "torch": torch.__version__}
    path = os.path.join(ROOT, "tests", "golden", "eval_reference.json")
    json.dump(out, open(path, "w"), indent=1)
    print(buf.getvalue()[:1500])
    print("wrote", path, summary)


if __name__ == "__main__":
    main()
