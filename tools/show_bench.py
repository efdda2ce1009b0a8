"""Print the headline fields of a bench.py JSON line: python tools/show_bench.py gpurun_out/bench_x.log"""
import json
import sys

for path in sys.argv[1:]:
    for line in open(path):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        print(path, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]),
              "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4), "clocks", d["clocks"].get("sm_mhz"))
        for k in d.get("kernels") or []:
            print("   ", k)
        print("    sweep", d.get("sweep"))
        print("    cpu", (d.get("cpu_baseline") or {}).get("value"), "sum_us", d.get("kernel_time_sum_us"))
