"""CPU (gloo, world_size 2 and 3) tests of the multi-GPU host logic: contiguous batch shards,
logits all-gather in rank order, agreement-count all-reduce. The compute leg is the oracle's
CPU forward, so the sharded result can be compared with the unsharded one bit-exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from int8inferenceengine_b200 import workloads as W
from int8inferenceengine_b200.sharding import gather_logits, reduce_count, shard_range


def test_shard_ranges_partition_the_batch():
    for gb in [1, 7, 100, 1000, 1001]:
        for world in [1, 2, 3, 4, 8]:
            spans = [shard_range(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a1 >= a0
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, gb, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import models
    topo = "lenet"
    sd = W.make_weights(topo, 0)
    pm = models.PortModel(topo, sd)
    pm.convert(pm.calibrate_minmax(W.make_images(topo, 100, 1)))
    x = W.make_images(topo, gb, 2)
    lo, hi = shard_range(gb, rank, world)
    local = torch.from_numpy(pm.forward_int8(x[lo:hi]))
    full = gather_logits(local, gb)
    ref_arg = pm.forward_int8(x).argmax(1)
    agree = reduce_count(int((local.numpy().argmax(1) == ref_arg[lo:hi]).sum()))
    if rank == 0:
        np.save(os.path.join(out_dir, "full.npy"), full.numpy())
        np.save(os.path.join(out_dir, "agree.npy"), np.array([agree]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,gb", [(2, 10), (3, 10)])
def test_sharded_forward_equals_unsharded(tmp_path, world, gb):
    mp.spawn(_worker, args=(world, _free_port(), gb, str(tmp_path)), nprocs=world, join=True)
    from oracle import models
    pm = models.PortModel("lenet", W.make_weights("lenet", 0))
    pm.convert(pm.calibrate_minmax(W.make_images("lenet", 100, 1)))
    exp = pm.forward_int8(W.make_images("lenet", gb, 2))
    got = np.load(tmp_path / "full.npy")
    assert np.array_equal(got, exp)
    assert int(np.load(tmp_path / "agree.npy")[0]) == gb
