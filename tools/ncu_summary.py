"""Summarise an .ncu-rep (raw page): one line per launch with the counters DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.csv]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = [
    ("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"), ("gpu__time_duration.sum", "dur_ns"),
    ("sm__cycles_active.max", "sm_cyc_max"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_pct"),
    ("smsp__inst_executed.sum", "inst"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    # tensor-pipe utilisation: share of elapsed cycles with the tensor pipe active (tcgen05.mma shows up in the hmma
    # sub-pipe counters), and the share of cycles the tensor-memory (TMEM) path is busy
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_pct"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "hmma_cyc"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tmem_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit"),
]


def find(key):   # section-qualified columns look like "TPC.TriageCompute.<metric>"
    for i, c in enumerate(hdr):
        if c == key or c.endswith("." + key):
            return i
    return None


idx = [(find(k), n) for k, n in want if find(k) is not None]
units = rows[1]
lines = [[n for _, n in idx]]
for r in rows[2:]:
    line = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.split("(")[0].replace("void ", "").replace("vb::", "")[:34]
        elif n in ("dram_rd", "dram_wr"):
            f = float(v.replace(",", "")) if v else 0.0
            u = units[i]
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            v = f"{f * mult / 1e6:.3f}MB"
        elif n == "dur_ns":
            f = float(v.replace(",", "")) if v else 0.0
            mult = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(units[i], 1e-3)
            v = f"{f * mult:.2f}us"
        line.append(v)
    lines.append(line)
# derived: tensor-pipe active share of the kernel = hmma sub-pipe active cycles / busiest SM's active cycles
names = lines[0]
if "hmma_cyc" in names and "sm_cyc_max" in names and "tensor_active_pct" not in names:
    a, b = names.index("hmma_cyc"), names.index("sm_cyc_max")
    names.append("tensor_active_pct")
    for l in lines[1:]:
        try:
            l.append(f"{100.0 * float(l[a].replace(',', '')) / max(float(l[b].replace(',', '')), 1.0):.1f}")
        except ValueError:
            l.append("")
w = [max(len(l[c]) for l in lines) for c in range(len(lines[0]))]
txt = "\n".join("  ".join(v.ljust(w[c]) for c, v in enumerate(l)) for l in lines)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
