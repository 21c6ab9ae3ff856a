#!/usr/bin/env python
"""Shared-memory wavefronts per source line of an ncu report (needs -lineinfo + --import-source on):
    python tools/ncu_smem_lines.py gpurun_out/x.ncu-rep <kernel-regex> [top N]
For the first captured launch of the matching kernel: L1 wavefronts, the excess over the conflict-free ideal, warp instructions."""
import csv, io, re, subprocess, sys

rep, want = sys.argv[1], re.compile(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
hdr, kern, cur, agg, seen = None, "", "", {}, []
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        kern = r[1]
        if want.search(kern):
            if (kern, cur) in seen:
                kern = ""               # a later launch of the same kernel: skip
            else:
                seen.append((kern, cur))
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r[0].isdigit() or not kern or not want.search(kern):
        continue
    d = dict(zip(hdr[4:], r[4:]))
    num = lambda k: int(d[k]) if d.get(k, "").isdigit() else 0
    wf = num("L1 Wavefronts Shared")
    if wf:
        key = (cur, int(r[0]), r[1].strip()[:110])
        a = agg.setdefault(key, [0, 0, 0, 0])
        a[0] += wf; a[1] += num("L1 Wavefronts Shared Excessive"); a[2] += num("L1 Wavefronts Shared Ideal"); a[3] += num("Instructions Executed")
tot, exc = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f"{seen[0][0].split('(')[0] if seen else '?'}: shared-memory wavefronts {tot}, of which excess over the ideal {exc} ({100 * exc / max(tot, 1):.1f} %)")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0]:10d} wf {100 * a[0] / tot:5.1f}%  excess {a[1]:10d}  ideal {a[2]:10d}  inst {a[3]:9d}  {key[0]}:{key[1]:<4d} {key[2]}")
