"""Per-kernel SASS opcode counts of libchessvision_b200.so (cuobjdump -sass): the mnemonics that prove the Blackwell paths
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk, SYNCS = mbarrier, FFMA2 / HFMA2
packed math).  Usage: python tools/sass_summary.py [lib.so] > profiles/rNN_sass_opcodes.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "chess_vision_b200", "libchessvision_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip() or n
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FFMA", "HFMA2", "HADD2", "LDS", "STS", "LDG", "STG", "BAR", "ELECT", "F2FP"]
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); kernels[cur] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        kernels[cur]["total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                kernels[cur][w] += 1
print(f"# {os.path.relpath(lib, ROOT)}: SASS opcode counts per kernel (cuobjdump -sass, sm_100a)")
print("kernel | total | " + " | ".join(WATCH))
for k, c in kernels.items():
    name = demangle(k).replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    if name.endswith(")"):                       # drop the trailing parameter list, keep the template arguments
        depth = 0
        for i in range(len(name) - 1, -1, -1):
            depth += name[i] == ")"
            depth -= name[i] == "("
            if depth == 0:
                name = name[:i]
                break
    name = name.replace("void ", "")
    print(f"{name} | {c['total']} | " + " | ".join(str(c[w]) for w in WATCH))
