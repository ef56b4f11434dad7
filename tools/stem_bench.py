#!/usr/bin/env python3
"""Dev tool: the AlexNet stem (input quantise + conv1) straight from fp32 images, rotating over
several distinct batches so that the input comes from HBM like in the real forward.
I8IE_NO_STEM_FUSEQ=1 selects the two-kernel path (stem_quantize + stem2)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from int8inferenceengine_b200 import backend as B  # noqa: E402
from gpu_utils import make_layer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=100)
    ap.add_argument("--ring", type=int, default=4)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    a = np.sqrt(6.0 / 363)
    w = rng.uniform(-a, a, size=(96, 3, 11, 11)).astype(np.float32)
    b = rng.uniform(-0.05, 0.05, size=(96,)).astype(np.float32)
    L = make_layer("conv", w, b, (np.float32(0.0518), 116), 4, 2)
    L.fuse_relu = True
    xs = [B.tensor_from_torch(torch.empty(args.batch, 3, 224, 224, device="cuda").uniform_(-2.1, 2.6))
          for _ in range(args.ring)]
    for x in xs:
        y = L.forward_quantize_fused(x, 0.025, 127)
    torch.cuda.synchronize()
    if os.environ.get("I8IE_STEM2_TRACE"):   # eager launches only: the library dumps CTA 0's timeline
        print("trace written to", os.environ["I8IE_STEM2_TRACE"])
        return
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(args.reps):
            for x in xs:
                y = L.forward_quantize_fused(x, 0.025, 127)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(args.iters):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / (args.reps * args.ring) * 1e3)
    print(f"stem (quantise + conv1) batch {args.batch}: {best:.2f} us per batch", flush=True)


if __name__ == "__main__":
    main()
