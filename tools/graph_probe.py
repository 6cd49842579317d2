#!/usr/bin/env python
"""How much of the step is launch gaps?  Captures one training step in a CUDA graph (scalars baked: timing probe only) and
compares replay time with eager launches."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
from unetb200.model import UNet

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
m = UNet(B.CLASSES, B.BATCH, B.CH, learning_rate=3e-5, precision="bf16", seed=0)
xs, ls = B.synthetic_host_batches(1, 0)
x = torch.from_numpy(xs[0]).to(dev)
l = torch.from_numpy(ls[0]).to(dev)
for _ in range(3):
    m.train_step(x, l)
torch.cuda.synchronize()
def timeit(fn, n=10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n
eager = timeit(lambda: m.train_step(x, l))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    m.train_step(x, l)
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    m.train_step(x, l)
graph = timeit(g.replay)
print(json.dumps({"eager_ms": eager, "graph_ms": graph, "launches_per_step": m.launches // 15}))
