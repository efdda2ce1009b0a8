"""Phase stamps of the optimizer tail kernel (a) as the last launch of a full step and (b) launched alone on the same
gradient partials (L2-warm from its own previous read).  VITB200_TIMELINE=1 python tools/tail_probe.py"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
assert os.environ.get("VITB200_TIMELINE") == "1", "set VITB200_TIMELINE=1"
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import BASELINE_CFG  # noqa: E402
from vit_b200 import _lib, get_model  # noqa: E402
from vit_b200.step import TrainStep  # noqa: E402

B = 64
dev = torch.device("cuda", 0)
m = get_model(json.loads(json.dumps(BASELINE_CFG)), precision="bf16-mixed", device=dev).train()
st = TrainStep(m, B, lr=1e-3, grad_clip=0.5, use_graph=True, train=True)
x, y = torch.rand(B, 4096, device=dev), torch.rand(B, device=dev)
lib = _lib.load()
labels = ["pdl_wait", "reduce partial slots", "block sum + ticket + wait", "norm, coef, bias corrections", "AdamW"]


def show(tag):
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (32 * 512))()
    assert lib.vitb200_tl_tail(buf) == 0
    a = np.frombuffer(buf, dtype=np.int64).reshape(512, 32)
    a = a[a[:, 0] != 0]
    if a[:, 9].any():   # streamed mode: stamp 9 = this block's gradient groups have been signalled
        late = a[:, 9] >= np.sort(a[:, 9])[-20]
        print(tag, "streamed: reduce after the group wait, all blocks mean", (a[:, 2] - a[:, 9]).mean(), "last 20 blocks mean",
              (a[late, 2] - a[late, 9]).mean(), "their wait end -> kernel end", (a[late, 5] - a[late, 9]).mean(),
              "| last wait end -> last end", a[:, 5].max() - a[:, 9].max())
    a = a[:, :len(labels) + 1]
    d = np.diff(a, axis=1)
    print(tag, " ".join(f"{lab.split()[0]}={d[:, i].mean():.0f}/{d[:, i].max()}" for i, lab in enumerate(labels)),
          f"first-start..last-end={a[:, -1].max() - a[:, 1].min()}")


for _ in range(5):
    st.step(x, y)
show("in step     :")
for i in range(3):
    st.eng.optimizer_step(fused_reduce=True)
    show(f"alone #{i}    :")
