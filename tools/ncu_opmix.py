"""Executed-instruction mix by SASS opcode for one kernel of an .ncu-rep.  python tools/ncu_opmix.py rep regex [n]"""
import csv
import subprocess
import sys
from collections import Counter

rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ci = {k: i for i, k in enumerate(hdr)}
seen = set()
mix = Counter()
for r in rows[2:]:
    if len(r) != len(hdr) or not r[ci["Instructions Executed"]].isdigit():
        continue
    if r[ci["Address"]] in seen:
        continue
    seen.add(r[ci["Address"]])
    src = r[ci["Source"]].strip()
    toks = [t for t in src.split() if not t.startswith("@")]
    op = toks[0].split(".")[0] if toks else "?"
    mix[op] += int(r[ci["Instructions Executed"]])
tot = sum(mix.values())
print("total warp instructions", tot)
for op, c in mix.most_common(n):
    print(f"{100 * c / tot:5.1f}%  {c:>9}  {op}")
