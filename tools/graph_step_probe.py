#!/usr/bin/env python3
"""Dev probe for the N-GPU step (round-2 item: one enqueue per step). Launch under torchrun:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29517 tools/graph_step_probe.py --batch 1000 --steps 200

Times, per rank and as the max over ranks, (a) the step bench.py runs — `Module.__call__` (its own
CUDA-graph replay) followed by the stream-ordered ResultExchange — and (b) the same forward +
exchange captured into ONE CUDA graph per input buffer (NCCL all-gather inside the capture), so that
a step is a single host enqueue. Verified on ONE GPU only (batch 125: 293.9 us plain, 288.4 us one-graph,
identical logits); the N > 1 run with NCCL inside the capture is a round-2 item. Nothing in the product or in
bench.py depends on it."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import i8ie  # noqa: E402
from int8inferenceengine_b200 import backend as B, sharding, workloads as W  # noqa: E402
from int8inferenceengine_b200.runner import build_module  # noqa: E402


def timed(fn, steps, world):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--ring", type=int, default=4)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    topo = "alexnet"
    lbatch = args.batch // world
    model = build_module(topo, W.make_weights(topo, 0), calib=W.make_images(topo, 100, 1))
    rng = np.random.default_rng(2 + rank)
    lo, hi = W.TOPOLOGIES[topo]["range"]
    xs = [i8ie.Tensor(B.tensor_from_torch(torch.from_numpy(
        rng.uniform(lo, hi, size=(lbatch, 3, 224, 224)).astype(np.float32)).cuda())) for _ in range(args.ring)]
    refs = [torch.zeros(lbatch, dtype=torch.int64, device="cuda") for _ in range(args.ring)]
    ex = sharding.ResultExchange(args.batch, 10, torch.device("cuda", local))

    def plain(i):
        out = model(xs[i % args.ring])
        ex(out.data.buf.view(lbatch, 10), refs[i % args.ring])

    ms_plain = timed(plain, args.steps, world)
    expect = ex.logits_all.clone()          # result of the last plain step (input (steps - 1) % ring)

    # (b) forward + exchange in one graph per input buffer
    model.graph = False                     # eager forward inside our own capture
    graphs = []
    for j in range(args.ring):
        for _ in range(2):
            out = model(xs[j])
            ex(out.data.buf.view(lbatch, 10), refs[j])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = model(xs[j])
            ex(out.data.buf.view(lbatch, 10), refs[j])
        graphs.append(g)

    def graphed(i):
        graphs[i % args.ring].replay()

    ms_graph = timed(graphed, args.steps, world)
    same = bool(torch.equal(ex.logits_all, expect))
    if rank == 0:
        print(f"N={world} batch {args.batch} ({lbatch}/GPU): plain step {ms_plain * 1e3:.1f} us, "
              f"one-graph step {ms_graph * 1e3:.1f} us, gathered logits identical: {same}", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
