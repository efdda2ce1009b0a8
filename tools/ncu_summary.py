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
    ("sm__pipe_tensor_subpipe_mma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__inst_executed_pipe_tensor_subpipe_mma.sum", "tensor_inst"),
    ("lts__t_sector_hit_rate.pct", "l2_hit"),
]
idx = [(hdr.index(k), n) for k, n in want if k in hdr]
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
w = [max(len(l[c]) for l in lines) for c in range(len(lines[0]))]
txt = "\n".join("  ".join(v.ljust(w[c]) for c, v in enumerate(l)) for l in lines)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
