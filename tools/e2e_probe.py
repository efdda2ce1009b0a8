# where does TrainStep.fit_host spend wall-clock time over a SHORT run (the driver's 20-step window)?
import json, os, sys, time, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from bench import BASELINE_CFG
from vit_b200 import get_model
from vit_b200.step import TrainStep
dev = torch.device("cuda:0")
torch.manual_seed(42)
m = get_model(json.loads(json.dumps(BASELINE_CFG)), precision="bf16-mixed", device=dev).train()
B = 64
st = TrainStep(m, B, use_graph=True, train=True)
g = torch.Generator().manual_seed(1)
hx = torch.rand(8, B, 4096, generator=g).pin_memory(); hy = torch.rand(8, B, generator=g).pin_memory()
def batches(n):
    for i in range(n):
        yield hx[i % 8], hy[i % 8]
st.fit_host(batches(5))
torch.cuda.synchronize()
for n in (20, 20, 40, 200):
    stamps = []
    t0 = time.perf_counter()
    st.fit_host(batches(n), on_loss=lambda i, v: stamps.append(time.perf_counter() - t0))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"n={n}: {dt*1e3:.3f} ms total, {dt/n*1e6:.1f} us/step, first losses at", [f"{s*1e6:.0f}" for s in stamps[:4]],
          "last", f"{stamps[-1]*1e6:.0f}", "us")
# device-side time of the same number of steps (events)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
x = torch.rand(B, 4096, device=dev); y = torch.rand(B, device=dev)
for _ in range(5): st.step(x, y)
torch.cuda.synchronize(); e0.record()
for _ in range(20): st.step(x, y)
e1.record(); torch.cuda.synchronize()
print("device step:", e0.elapsed_time(e1) / 20 * 1e3, "us")
