"""Per-kernel SASS opcode counts of the shipped library (no GPU needed): the mnemonics that show what the kernels are
made of -- bulk asynchronous copies (UBLKCP), mbarrier ops (SYNCS), PDL (ACQBULK / PREEXIT), 128-bit streaming stores
(STG.E.EF.128), the integer work of the permute transposes (LOP3 / PRMT / SHF), redux.sync (REDUX), warp votes (VOTE).
No tensor-core instruction is expected: the path has no contraction.

    python tools/sass_summary.py [lib] > profiles/r2_sass_summary.txt
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "csgn_b200/lib/libcsgn.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
COLS = ["UBLKCP", "SYNCS", "PDL", "STG.EF.128", "STG", "LDG", "LDS", "STS", "LOP3", "PRMT", "SHF", "REDUX", "VOTE", "BAR", "MMA"]
counts, total, fn = collections.defaultdict(collections.Counter), collections.Counter(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
    if not m or fn is None:
        continue
    op = m.group(1)
    total[fn] += 1
    c = counts[fn]
    if op.startswith("UBLKCP"): c["UBLKCP"] += 1
    if op.startswith("SYNCS"): c["SYNCS"] += 1
    if op.startswith("ACQBULK") or op.startswith("PREEXIT"): c["PDL"] += 1
    if op.startswith("STG.E.EF.128"): c["STG.EF.128"] += 1
    for k in ("STG", "LDG", "LDS", "STS", "LOP3", "PRMT", "SHF", "REDUX", "VOTE", "BAR"):
        if op.startswith(k): c[k] += 1
    if "MMA" in op: c["MMA"] += 1
names = subprocess.run(["c++filt"], input="\n".join(total), stdout=subprocess.PIPE, text=True).stdout.splitlines()
rows = []
for mangled, name in zip(total, names):
    short = re.sub(r"\(anonymous namespace\)::", "", name)
    short = re.sub(r"^void ", "", short).replace("csgn::", "")
    short = short.split("(")[0]
    rows.append((short, mangled))
print("%-60s %6s " % ("kernel (%s)" % lib, "instr") + " ".join("%10s" % c for c in COLS))
for short, mangled in sorted(rows):
    print("%-60s %6d " % (short[:60], total[mangled]) + " ".join("%10d" % counts[mangled][c] for c in COLS))
print("\n%d kernels; tensor-core (MMA) instructions: %d" % (len(rows), sum(counts[m]["MMA"] for m in total)))
