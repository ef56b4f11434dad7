#!/usr/bin/env python3
"""In-situ kernel timeline of the AlexNet INT8 forward (CUDA-graph replay) via CUPTI
(torch.profiler): per-kernel average duration INSIDE the real step (warm L2, real neighbours)
and the idle gaps between kernels. Dev tool; summaries are copied to profiles/."""
import argparse
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import i8ie  # noqa: E402
from int8inferenceengine_b200 import backend as B, workloads as W  # noqa: E402
from int8inferenceengine_b200.runner import build_module  # noqa: E402


def short(name):
    name = name.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    base = name.split("(")[0]
    return base[:80]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=100)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--topology", default="alexnet")
    args = ap.parse_args()
    topo = args.topology
    model = build_module(topo, W.make_weights(topo, 0), calib=W.make_images(topo, 100, 1))
    xs = [i8ie.Tensor(B.tensor_from_torch(torch.from_numpy(W.make_images(topo, args.batch, 2 + i)).cuda()))
          for i in range(3)]
    for i in range(5):
        model(xs[i % 3])
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(args.steps):
            model(xs[i % 3])
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.name != "Memset (Device)"]
    evs.sort(key=lambda e: e.time_range.start)
    agg = OrderedDict()
    busy = 0.0
    for e in evs:
        d = e.time_range.end - e.time_range.start
        busy += d
        a = agg.setdefault(short(e.name), [0, 0.0])
        a[0] += 1
        a[1] += d
    span = evs[-1].time_range.end - evs[0].time_range.start
    print(f"# {topo} batch {args.batch}: {args.steps} steps, {len(evs)} kernels, span {span / args.steps:.1f} us/step, "
          f"kernel-busy {busy / args.steps:.1f} us/step, idle {100 * (1 - busy / span):.1f} %")
    print("| us/step | launches/step | avg us | kernel |")
    print("|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {t / args.steps:7.2f} | {c / args.steps:.1f} | {t / c:7.2f} | `{k}` |")
    # ordered list of one step
    per = len(evs) // args.steps
    print("\n# kernel order of the last step (start offset us, duration us)")
    last = evs[-per:]
    t0 = last[0].time_range.start
    for e in last:
        print(f"  {e.time_range.start - t0:8.1f} {e.time_range.end - e.time_range.start:7.2f}  {short(e.name)}")


if __name__ == "__main__":
    main()
