#!/usr/bin/env python3
"""Dev tool: where the end-to-end step goes — raw pinned H2D of one batch vs the API call."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import i8ie  # noqa: E402
from int8inferenceengine_b200 import backend as B, workloads as W  # noqa: E402
from int8inferenceengine_b200.runner import build_module  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 100
model = build_module("alexnet", W.make_weights("alexnet", 0), calib=W.make_images("alexnet", 100, 1))
hs = [torch.from_numpy(W.make_images("alexnet", batch, 2 + i)).pin_memory() for i in range(3)]
dev = torch.empty_like(hs[0], device="cuda")


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        fn(i) if fn.__code__.co_argcount else fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def raw():
    dev.copy_(hs[0], non_blocking=True)
    torch.cuda.synchronize()


def api(i=0):
    return model(i8ie.tensor(hs[i % 3])).numpy()


print(f"raw pinned H2D of {hs[0].numel() * 4 / 1e6:.1f} MB: {timed(raw):.3f} ms")
for c in ("1", "2", "4", "5", "10"):
    B._h2d_chunks = (lambda rows, nbytes, c=int(c): c if rows % c == 0 else 1)
    for _ in range(4):
        api()
    print(f"api step, {c} chunks: {timed(api):.3f} ms")
