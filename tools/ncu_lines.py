#!/usr/bin/env python
"""Per-source-line summary of an ncu report (needs -lineinfo + --import-source on):
    python tools/ncu_lines.py gpurun_out/x.ncu-rep [kernel-regex] [top N]
Prints, for the lines with the most warp-stall samples: samples, share, warp instructions executed, the
dominant stall reasons."""
import csv
import io
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    cur_file = ""
    agg = {}
    kern = ""
    want = re.compile(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] else None
    seen_kernels = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            kern = r[1]
            if kern not in seen_kernels:
                seen_kernels.append(kern)
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] == "" or not r[0].isdigit():
            continue
        if want and not want.search(kern):
            continue
        if len(seen_kernels) > 1 and kern != seen_kernels[0] and not want:
            continue
        d = dict(zip(hdr[4:], r[4:]))
        key = (cur_file, int(r[0]), r[1].strip()[:90])
        a = agg.setdefault(key, {"samples": 0, "inst": 0, "stalls": {}})
        def num(v):
            try:
                return int(v)
            except (TypeError, ValueError):
                return 0
        a["samples"] += num(d.get("# Samples"))
        a["inst"] += num(d.get("Instructions Executed"))
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "-", "0"):
                a["stalls"][k[6:]] = a["stalls"].get(k[6:], 0) + num(v)
    tot = sum(a["samples"] for a in agg.values()) or 1
    toti = sum(a["inst"] for a in agg.values()) or 1
    print(f"total samples {tot}, warp instructions {toti}")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = ", ".join(f"{k}:{v}" for k, v in sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3])
        print(f"{a['samples']:7d} {100 * a['samples'] / tot:5.1f}%  inst {a['inst']:9d} {100 * a['inst'] / toti:5.1f}%  {key[0]}:{key[1]:<4d} {key[2]}  [{st}]")


if __name__ == "__main__":
    main()
