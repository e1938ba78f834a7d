"""Does capturing the rollout in a CUDA graph help? (launch-gap estimate)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

torch.set_grad_enabled(False)

from common import build_ddpm  # noqa: E402

T, B = 4, 256
net, sampler, value, sd, vsd = build_ddpm(T)
noise = torch.randn(T + 1, B, 3, 32, 32, device="cuda")


def run():
    d = sampler.sample(B, device="cuda", noise=noise)
    return value(d["sample"], T)


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
print("eager   ms/step", e0.elapsed_time(e1) / 10)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    run()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = run()
torch.cuda.synchronize()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    g.replay()
e1.record()
torch.cuda.synchronize()
print("graphed ms/step", e0.elapsed_time(e1) / 10, float(out.mean()))
