"""The evaluation restatement (oracle/eval_oracle.py) against the REFERENCE's own evaluate() (tests/golden/eval_reference.json,
written by oracle/make_golden_eval.py from /root/reference/evaluate.py:47-231 run unmodified)."""
import json
import os
import re

import numpy as np

from oracle import eval_oracle as eo

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "eval_reference.json")))


def run_oracle():
    c = np.zeros(eo.N_COUNTERS, dtype=np.int64)
    per, loss = [], []
    for seed, n in GOLD["batches"]:
        ci, pi, li = eo.evaluate_batch(eo.synth_eval_batch(seed, n))
        c += ci
        per.append(pi)
        loss.append(li)
    return c, np.concatenate(per), np.concatenate(loss)


def parse_report(text):
    """Numbers out of the reference's printed report (evaluate.py:157-216)."""
    out = {}
    rows = []
    in_conf = False
    for line in text.splitlines():
        if line.startswith("Confusion matrix"):
            in_conf = True
            continue
        if in_conf:
            m = re.match(r"\s+(empty|[PNBRQKpnbrqk])\s+((?:-?\d+\s*){13})$", line)
            if m:
                rows.append([int(v) for v in m.group(2).split()])
            elif rows and len(rows) == 13:
                in_conf = False
    out["confusion"] = np.array(rows, dtype=np.int64)
    out["piece"] = {m.group(1): (int(m.group(2)), int(m.group(3))) for m in re.finditer(r"^\s+(empty|[PNBRQKpnbrqk]): [\d.]+\s+\((\d+)/(\d+)\)$", text, re.M)}
    out["castling"] = [int(m.group(1)) for m in re.finditer(r"^  [KQkq]: [\d.]+ \((\d+)/\d+\)$", text, re.M)]
    w = re.search(r"White\s+(\d+)\s+(\d+)\n\s+Black\s+(\d+)\s+(\d+)", text)
    out["turn_confusion"] = [int(w.group(i)) for i in range(1, 5)]
    out["worst"] = [(int(m.group(1)), int(m.group(2))) for m in re.finditer(r"Image (\d+): (\d+)/64 squares wrong", text)]
    return out


def test_oracle_matches_reference_summary():
    c, per, loss = run_oracle()
    s, ref = eo.summary(c, loss.sum()), GOLD["summary"]
    for k in ("square_acc", "board_acc", "turn_acc", "castling_acc", "full_fen_acc"):
        assert s[k] == ref[k], k                       # ratios of identical integers
    assert s["total_boards"] == ref["total_boards"] and s["total_legal"] == ref["total_legal"]
    assert abs(s["loss"] - ref["loss"]) < 1e-5 * ref["loss"]      # the reference averages in fp32 per batch


def test_oracle_matches_reference_report():
    c, per, _ = run_oracle()
    r = parse_report(GOLD["report"])
    assert r["confusion"].shape == (13, 13)
    assert np.array_equal(c[eo.CONFUSION:eo.CONFUSION + 169].reshape(13, 13), r["confusion"])
    names = ["empty"] + list("PNBRQKpnbrqk")
    for k, nm in enumerate(names):
        assert (int(c[eo.PIECE_CORRECT + k]), int(c[eo.PIECE_TOTAL + k])) == r["piece"][nm], nm
    assert [int(v) for v in c[eo.CORRECT_CASTLING_RIGHT:eo.CORRECT_CASTLING_RIGHT + 4]] == r["castling"]
    assert [int(v) for v in c[eo.TURN_CONFUSION:eo.TURN_CONFUSION + 4]] == r["turn_confusion"]
    worst = sorted(((int(per[i, 0]), i) for i in range(len(per)) if per[i, 0] > 0), key=lambda x: -x[0])[:10]
    assert [(i, n) for n, i in worst] == r["worst"]


def test_counter_layout_matches_header():
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "chessvision_b200.h")).read()
    for name in ("TOTAL_BOARDS", "TOTAL_SQUARES", "CORRECT_SQUARES", "CORRECT_BOARDS", "TOTAL_LEGAL", "CORRECT_TURN", "CORRECT_CASTLING_RIGHT",
                 "CORRECT_CASTLING_ALL", "CORRECT_FULL_FEN", "PIECE_CORRECT", "PIECE_TOTAL", "CONFUSION", "TURN_CONFUSION"):
        m = re.search(rf"CV_EVAL_{name} = (\d+)", hdr)
        assert m and int(m.group(1)) == getattr(eo, name), name
    assert int(re.search(r"CV_EVAL_COUNTERS = (\d+)", hdr).group(1)) == eo.N_COUNTERS
    from chess_vision_b200 import evaluate as ev
    for name in ("TOTAL_BOARDS", "CORRECT_FULL_FEN", "PIECE_CORRECT", "PIECE_TOTAL", "CONFUSION", "TURN_CONFUSION", "N_COUNTERS"):
        assert getattr(ev, name) == getattr(eo, name)


def test_grouped_report_matches_the_reference_line_for_line():
    """chess_vision_b200.evaluate.grouped_report (host dictionary work over the kernel's per-sample table) against the text the
    reference's own print_grouped_metrics (evaluate.py:233-287) printed for the same manifest rows and per-sample results
    (tests/golden/eval_grouped_reference.json, oracle/make_golden_eval_grouped.py)."""
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "eval_grouped_reference.json")))
    meta, per = eo.synth_manifest(gold["seed"], gold["n"])
    from chess_vision_b200.evaluate import grouped_metrics, grouped_report
    assert grouped_report(eo.ManifestStub(meta), per) + "\n" == gold["report"]
    g = grouped_metrics(eo.ManifestStub(meta), per)
    assert list(g) == ["piece_count", "castling", "turn", "has_highlight", "style", "flipped"]
    assert sum(v["total"] for v in g["flipped"].values()) == gold["n"]

    class NoManifest:
        use_manifest = False
    assert grouped_report(NoManifest(), per) == "" and grouped_metrics(NoManifest(), per) == {}
