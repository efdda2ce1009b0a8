"""Per-call device times of one fp32 training step at the baseline shape.  python tools/fp32_probe.py [batch]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import BASELINE_CFG, time_calls  # noqa: E402
from vit_b200 import get_model  # noqa: E402
from vit_b200.step import TrainStep  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
m = get_model(json.loads(json.dumps(BASELINE_CFG)), precision="32", device=dev).train()
st = TrainStep(m, B, lr=1e-3, grad_clip=0.5, use_graph=True, train=True)
x, y = torch.rand(B, 4096, device=dev), torch.rand(B, device=dev)
for _ in range(3):
    st.step(x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    st.step(x, y)
e1.record()
torch.cuda.synchronize()
print(f"fp32 B={B}: step {e0.elapsed_time(e1) / 20 * 1e3:.1f} us, launches {st.kernel_launches()}")
eng = st.eng
tot = {}
for key, prog in eng._progs.items():
    if key[0] not in ("fwd", "bwd"):
        continue
    for name, args, sec in time_calls(eng, [prog], repeats=5, iters=3):
        t = tot.setdefault((key[0], name.replace("vitb200_", "")), [0.0, 0])
        t[0] += sec * 1e6
        t[1] += 1
s = 0.0
for (d, n), (us, k) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"  {d} {n:<28} x{k:<3} {us:8.1f} us total")
    s += us
print("  sum", round(s, 1), "us")
