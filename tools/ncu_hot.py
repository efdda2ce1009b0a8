"""Top stall locations of one kernel from an .ncu-rep source page (SASS rows with sampling counts).

    python tools/ncu_hot.py rep.ncu-rep kernel_regex [n]
"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ci = {k: i for i, k in enumerate(hdr)}
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
body = [r for r in rows[2:] if len(r) == len(hdr) and r[ci["# Samples"]].isdigit()]
tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
agg = {k: sum(int(r[ci[k]] or 0) for r in body) for k in stalls}
print("total samples", tot, "instructions", sum(int(r[ci["Instructions Executed"]] or 0) for r in body))
print("stall mix:", ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
body.sort(key=lambda r: -int(r[ci["# Samples"]] or 0))
for r in body[:n]:
    s = int(r[ci["# Samples"]] or 0)
    top = max(stalls, key=lambda k: int(r[ci[k]] or 0))
    print(f"{100 * s / max(tot, 1):5.1f}%  {r[ci['Instructions Executed']]:>8}  {top[6:]:<12} {r[ci['Source']][:110]}")
