"""Per-call device times of one training step at a long sequence (stride 8 -> T = 510, stride 2 -> T = 2034).

    python tools/longseq_probe.py [stride] [batch]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import BASELINE_CFG, time_calls  # noqa: E402
from vit_b200 import get_model  # noqa: E402
from vit_b200.step import TrainStep  # noqa: E402

stride = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)
cfg = json.loads(json.dumps(BASELINE_CFG))
cfg["model"]["stride_size"] = stride
m = get_model(cfg, precision="bf16-mixed", device=dev).train()
st = TrainStep(m, B, lr=1e-3, grad_clip=0.5, use_graph=False, train=True)
x, y = torch.rand(B, 4096, device=dev), torch.rand(B, device=dev)
for _ in range(2):
    st.step(x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    st.step(x, y)
e1.record()
torch.cuda.synchronize()
print(f"T={m.config.tokens} B={B}: step {e0.elapsed_time(e1) / 3:.3f} ms")
eng = st.eng
tot = {}
for key, prog in eng._progs.items():
    if key[0] not in ("fwd", "bwd"):
        continue
    for name, args, sec in time_calls(eng, [prog], repeats=2, iters=2):
        t = tot.setdefault((key[0], name.replace("vitb200_", "")), [0.0, 0])
        t[0] += sec * 1e3
        t[1] += 1
for (d, n), (ms, k) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"  {d} {n:<28} x{k:<3} {ms:8.3f} ms total")
