"""Run a few eager (non-graph) training steps of the bench workload: the target of `ncu --set full`.

    ncu --set full --import-source on --clock-control none --launch-skip 66 --launch-count 22 \
        -o gpurun_out/prof python tools/prof_step.py [--batch 64] [--steps 4]

22 kernels per step (3 layers): skip the first 3 steps, capture the 4th.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import BASELINE_CFG  # noqa: E402
from vit_b200 import get_model  # noqa: E402
from vit_b200.step import TrainStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=4)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(42)
m = get_model(json.loads(json.dumps(BASELINE_CFG)), precision="bf16-mixed", device=dev).train()
st = TrainStep(m, a.batch, lr=1e-3, grad_clip=0.5, use_graph=False, train=True)
x = torch.rand(a.batch, 4096, device=dev)
y = torch.rand(a.batch, device=dev)
for _ in range(a.steps):
    st.step(x, y)
torch.cuda.synchronize()
print("launches/step", st.kernel_launches(), "loss", float(st.eng.loss[0]))
