#!/usr/bin/env python3
"""Dev tool: host-side profile (cProfile) of the end-to-end call model(i8ie.tensor(pinned)).numpy()."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import i8ie  # noqa: E402
from int8inferenceengine_b200 import workloads as W  # noqa: E402
from int8inferenceengine_b200.runner import build_module  # noqa: E402

model = build_module("alexnet", W.make_weights("alexnet", 0), calib=W.make_images("alexnet", 100, 1))
hs = [torch.from_numpy(W.make_images("alexnet", 100, 2 + i)).pin_memory() for i in range(3)]
for i in range(12):
    model(i8ie.tensor(hs[i % 3])).numpy()
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(30):
    model(i8ie.tensor(hs[i % 3])).numpy()
print(f"step {(time.perf_counter() - t0) / 30 * 1e3:.3f} ms")
# phases
ts = [0.0, 0.0, 0.0]
for i in range(30):
    a = time.perf_counter(); x = i8ie.tensor(hs[i % 3]); b = time.perf_counter(); y = model(x); c = time.perf_counter(); z = y.numpy(); d = time.perf_counter()
    ts[0] += b - a; ts[1] += c - b; ts[2] += d - c
print("host ms per step: tensor() %.3f  model() %.3f  numpy() %.3f" % tuple(t / 30 * 1e3 for t in ts))
pr = cProfile.Profile()
pr.enable()
for i in range(30):
    model(i8ie.tensor(hs[i % 3])).numpy()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
