"""Where does the HOST time of an eager layer-step go?  cProfile + wall-clock split of
SparseConvTensor -> layer(s) -> loss -> backward through the public API (device-resident inputs).
Usage: python tools/host_profile.py [cfg2|encoder] [reps]"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import spconv_b200.pytorch as spconv
from bench_utils import make_encoder6, surface_cloud
from spconv_b200.pytorch import ops

which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda:0")
shape = [41, 1600, 1408]
rng = np.random.default_rng(50051)
inds = torch.from_numpy(surface_cloud(rng, shape, 100_000)).to(dev)
if which == "cfg2":
    C = 64
    layers = [spconv.SubMConv3d(64, 64, 3, padding=1, bias=False).to(dev).half()]
else:
    C = 16
    layers = [m.to(dev).half() for m in make_encoder6(spconv)]
feats = torch.randn(inds.shape[0], C, device=dev).half()


def step():
    x = spconv.SparseConvTensor(feats.detach().requires_grad_(True), inds, shape, 1)
    for m in layers:
        m.weight.grad = None
        x = m(x)
    loss = x.features.square().mean(dtype=torch.float32)
    loss.backward()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
# (a) wall clock with the GPU kept busy-free: host time per step = wall / reps when host-bound
t0 = time.perf_counter()
for _ in range(reps):
    step()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_total = time.perf_counter() - t0
ops.launch_count(reset=True)
step()
launches = ops.launch_count(reset=True)
print(f"[{which}] wall per step: issue {1e3 * t_issue / reps:.3f} ms, incl. drain {1e3 * t_total / reps:.3f} ms; "
      f"library launches per step {launches}")
pr = cProfile.Profile()
pr.enable()
for _ in range(reps):
    step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue())
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30)
print(s.getvalue())
