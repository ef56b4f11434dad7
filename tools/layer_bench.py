#!/usr/bin/env python3
"""Per-layer kernel timing without Python launch overhead: each AlexNet layer's INT8 kernel is
captured R times into a CUDA graph and replayed; reports us/launch and TOP/s on the
reference's un-padded dims. Dev tool (also used to produce profiles/)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from int8inferenceengine_b200 import backend as B, workloads as W  # noqa: E402
from gpu_utils import make_layer  # noqa: E402


def layer_specs(topology, batch):
    t = W.TOPOLOGIES[topology]
    c, h, w = t["input"]
    out = []
    for op in t["ops"]:
        if op[0] == "conv":
            _, name, cin, cout, k, s, p = op
            oh, ow = W.conv_out_hw(h, w, k, s, p)
            out.append((name, "conv", (batch, cin, h, w), (cout, cin, k, k), s, p, batch * oh * ow * cout * cin * k * k))
            c, h, w = cout, oh, ow
        elif op[0] == "pool":
            h, w = (h - op[1]) // op[2] + 1, (w - op[1]) // op[2] + 1
        elif op[0] == "fc":
            out.append((op[1], "fc", (batch, op[2]), (op[3], op[2]), 1, 0, batch * op[2] * op[3]))
    return out


def run(topology="alexnet", batch=100, layers="", reps=20, iters=10, impl=0, relu=1, cp=0, quiet=False):
    """Time every selected layer; returns [(name, us per launch, TOP/s, impl)]."""
    rng = np.random.default_rng(0)
    want = set(layers.split(",")) if layers else None
    rows = []
    for name, kind, xshape, wshape, s, p, macs in layer_specs(topology, batch):
        if want and name not in want:
            continue
        fan = int(np.prod(wshape[1:]))
        a = np.sqrt(6.0 / fan)
        w = rng.uniform(-a, a, size=wshape).astype(np.float32)
        b = rng.uniform(-0.05, 0.05, size=(wshape[0],)).astype(np.float32)
        L = make_layer(kind, w, b, (np.float32(0.1), 120), s, p)
        L.fuse_relu = bool(relu)
        q = torch.randint(0, 256, xshape, dtype=torch.uint8, device="cuda")
        if kind == "conv":
            n, c, h, ww = xshape
            pitch = cp or B._act_pitch(c) if hasattr(B, "_act_pitch") else (c + 15) // 16 * 16
            buf = torch.full((n, h, ww, pitch), 127, dtype=torch.uint8, device="cuda")
            buf[..., :c] = q.permute(0, 2, 3, 1)
            x = B.TensorU8(B._Storage(buf.reshape(-1)), list(xshape), "nhwc", (n, c, h, ww, pitch), 0.05, 127)
        else:
            m, k = xshape
            x = B.TensorU8(B._Storage(q.reshape(-1)), list(xshape), "nhwc", (m, k, 1, 1, k), 0.05, 127)
        del q
        for _ in range(3):
            y = L._forward_u8(x, impl=impl)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                y = L._forward_u8(x, impl=impl)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(iters):
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps * 1e3)
        tops = 2 * macs / (best * 1e-6) / 1e12
        used = getattr(L, "_last_impl", "-")
        rows.append((name, best, tops, used))
        if not quiet:
            print(f"{name:6s} impl={used} {best:9.2f} us  {tops:8.1f} TOP/s  (MACs {macs / 1e9:.2f} G)", flush=True)
        del g, x, y, L
        torch.cuda.empty_cache()
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=100)
    ap.add_argument("--topology", default="alexnet")
    ap.add_argument("--layers", default="")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--impl", type=int, default=0)
    ap.add_argument("--relu", type=int, default=1)
    ap.add_argument("--cp", type=int, default=0, help="force the input channel pitch (conv)")
    args = ap.parse_args()
    rows = run(args.topology, args.batch, args.layers, args.reps, args.iters, args.impl, args.relu, args.cp)
    print(f"sum {sum(r[1] for r in rows):.1f} us")


if __name__ == "__main__":
    main()
