#!/usr/bin/env python3
"""Dev tool (torchrun, N >= 2): where the sharded step's exchange term goes. Times, per rank and MAX over ranks,
  a) the forward graph alone            b) exchange alone (pack-push + wait-unpack in a graph)
  c) forward + exchange in one graph (the bench's step)
for AlexNet at global batch 1000."""
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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gb = int(os.environ.get("GB", "1000"))
    lb = gb // world
    dev = torch.device("cuda", local)
    model = build_module("alexnet", W.make_weights("alexnet", 0), calib=W.make_images("alexnet", 100, 1))
    rng = np.random.default_rng(2 + rank)
    xs = [i8ie.Tensor(B.tensor_from_torch(torch.from_numpy(rng.uniform(-2.1, 2.6, size=(lb, 3, 224, 224)).astype(np.float32)).cuda()))
          for _ in range(2)]
    refs = [torch.zeros(lb, dtype=torch.int64, device=dev) for _ in range(2)]
    ex = sharding.make_exchange(gb, 10, dev)
    step = sharding.ShardedStep(model, xs, refs, ex)
    logits = torch.zeros(lb, 10, dtype=torch.float32, device=dev)
    for _ in range(3):
        ex(logits, refs[0])
    torch.cuda.synchronize()
    g_ex = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_ex):
        ex(logits, refs[0])

    def timed(fn, n=200):
        for i in range(10):
            fn(i)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            fn(i)
        b.record()
        torch.cuda.synchronize(); dist.barrier()
        ms = a.elapsed_time(b) / n
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return ms, float(t.item())

    for i in range(6):
        model(xs[i % 2])
    res = {"forward alone": timed(lambda i: model(xs[i % 2])),
           "exchange alone": timed(lambda i: g_ex.replay()),
           "forward + exchange": timed(lambda i: step(i))}
    for k, (mine, mx) in res.items():
        allv = [None] * world
        dist.all_gather_object(allv, round(mine * 1e3, 1))
        if rank == 0:
            print(f"{k:20s} max {mx * 1e3:7.1f} us   per rank {allv}", flush=True)
    ex.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
