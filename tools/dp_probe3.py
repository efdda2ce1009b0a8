"""2+ GPU check of the in-kernel peer-memory all-reduce against the NCCL path: same parameters after N steps, step time of both."""
import faulthandler
import json
import os
import sys
import time

faulthandler.dump_traceback_later(100, exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import BASELINE_CFG  # noqa: E402
from vit_b200 import dp, get_model  # noqa: E402
from vit_b200.step import TrainStep  # noqa: E402

t0 = time.time()
rank, local, world = dp.init_from_env("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)


def mark(msg):
    print(f"[{rank}] {time.time() - t0:6.1f}s {msg}", flush=True)


B = 64
g = torch.Generator().manual_seed(100 + rank)
xs = [torch.rand(B, 4096, generator=g).to(dev) for _ in range(4)]
ys = [torch.rand(B, generator=g).to(dev) for _ in range(4)]
res = {}
for mode in ("nccl", "peer"):
    torch.manual_seed(7)
    m = get_model(json.loads(json.dumps(BASELINE_CFG)), precision="bf16-mixed", device=dev).train()
    dp.broadcast_parameters(m._arena.data)
    step = TrainStep(m, B, use_graph=True, world_size=world, train=True, peer_allreduce=(mode == "peer"))
    losses = [float(step.step(xs[i % 4], ys[i % 4])) for i in range(6)]
    mark(f"{mode}: losses {losses}")
    flat = m._arena.data.clone()
    ref = flat.clone(); dist.broadcast(ref, src=0)
    assert torch.equal(ref, flat), f"{mode}: replicas diverged"
    res[mode] = flat
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(300):
        step.step(xs[i % 4], ys[i % 4])
    e1.record()
    torch.cuda.synchronize()
    mark(f"{mode}: {e0.elapsed_time(e1) / 300 * 1e3:.1f} us/step")
    step.close()
d = float((res["nccl"] - res["peer"]).abs().max() / res["nccl"].abs().max())
mark(f"peer vs nccl parameter difference after 6 steps: {d:.3e}")
assert d < 1e-3, d   # (the two paths sum the gradient partials in different fixed groupings; Adam amplifies the rounding)
dist.barrier()
dist.destroy_process_group()
mark("done")
