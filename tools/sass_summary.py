"""Per-kernel counts of the Blackwell-only SASS opcodes in the in-tree library (runs without a GPU):
    python tools/sass_summary.py > profiles/r2_sass_summary.txt
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG / UTMASTG = TMA tensor load / store,
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, UCGABAR = cluster barrier, DSMEM = generic-address ST.E (st.shared::cluster into the peer CTA)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "vit_b200", "libvitb200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
ops = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTMAPF", "UCGABAR", "SYNCS", "MUFU.EX2", "ELECT"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    counts[cur]["_n"] += 1
    for o in ops:
        if op.startswith(o):
            counts[cur][o] += 1
            total[o] += 1
    if op in ("ST.E", "ST.E.128"):   # generic-address stores: st.shared::cluster into the peer CTA (no other generic stores exist)
        counts[cur]["DSMEM"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.relpath(lib, ROOT)}: SASS opcode counts per kernel (cuobjdump -sass, sm_100a); kernels without tcgen05 / TMA omitted")
print(f"# {'kernel':<70} instr " + " ".join(f"{o:>8}" for o in ops) + "    DSMEM")
for (name, c), dm in zip(counts.items(), demangle):
    if not any(c[o] for o in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG")):
        continue
    short = re.sub(r"\(.*", "", dm).replace("vb::", "")
    print(f"{short:<72} {c['_n']:>5} " + " ".join(f"{c[o]:>8}" for o in ops) + f" {c['DSMEM']:>8}")
print("# total: " + ", ".join(f"{o} {total[o]}" for o in ops))
