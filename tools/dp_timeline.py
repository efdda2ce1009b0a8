"""Phase stamps of the data-parallel optimizer tail (debug build; run under torchrun with VITB200_TIMELINE=1)."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
assert os.environ.get("VITB200_TIMELINE") == "1"
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import BASELINE_CFG  # noqa: E402
from vit_b200 import _lib, dp, get_model  # noqa: E402
from vit_b200.step import TrainStep  # noqa: E402

rank, local, world = dp.init_from_env("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
torch.manual_seed(7)
m = get_model(json.loads(json.dumps(BASELINE_CFG)), precision="bf16-mixed", device=dev).train()
dp.broadcast_parameters(m._arena.data)
step = TrainStep(m, 64, use_graph=True, world_size=world, train=True)
x = torch.rand(64, 4096, device=dev); y = torch.rand(64, device=dev)
for _ in range(50):
    step.step(x, y)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * (32 * 512))()
assert lib.vitb200_tl_tail(buf) == 0
a = np.frombuffer(buf, dtype=np.int64).reshape(512, 32)
a = a[a[:, 0] != 0]
order = [0, 1, 6, 8, 2, 3, 4, 5]
labels = ["pdl_wait", "slot reduce + push to every rank", "poll own buffer + sum in rank order", "warp sums",
          "publish block partial", "grid barrier (two hops) + norm / coef", "AdamW"]
t = a[:, order]
d = np.diff(t, axis=1)
if rank == 0:
    print(f"world {world}: tail kernel, {a.shape[0]} CTAs, total {d.sum(1).mean():.0f} cycles")
    for i, lab in enumerate(labels):
        print(f"   {lab:<34} mean {d[:, i].mean():8.0f}   max {d[:, i].max():8d}")
dist.barrier()
step.close()
dist.destroy_process_group()
