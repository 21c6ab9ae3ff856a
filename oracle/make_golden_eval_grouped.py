"""Generate tests/golden/eval_grouped_reference.json with the REFERENCE's own print_grouped_metrics (build container only).

/root/reference/evaluate.py:233-287 is imported unmodified (over oracle/timm_shim) and run on a stub manifest dataset and seeded
per-sample results; the printed text is stored with the inputs so that chess_vision_b200.evaluate.grouped_report can be compared
line for line (tests/test_eval_cpu.py).

    python oracle/make_golden_eval_grouped.py
"""
import contextlib
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "timm_shim"))
sys.path.insert(0, "/root/reference")

import evaluate as ref_evaluate  # noqa: E402  /root/reference/evaluate.py
from oracle.eval_oracle import ManifestStub, synth_manifest  # noqa: E402


def main():
    meta, per = synth_manifest(5, 97)
    results = [{"idx": i, "num_wrong": int(per[i, 0]), "board_correct": int(per[i, 1]),
                "turn_correct": None if per[i, 2] == 255 else int(per[i, 2]),
                "castling_correct": None if per[i, 3] == 255 else int(per[i, 3])} for i in range(len(meta))]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ref_evaluate.print_grouped_metrics(ManifestStub(meta), results)
    path = os.path.join(ROOT, "tests", "golden", "eval_grouped_reference.json")
    json.dump({"seed": 5, "n": 97, "report": buf.getvalue()}, open(path, "w"), indent=1)
    print(buf.getvalue())
    print("wrote", path)


if __name__ == "__main__":
    main()
