#!/usr/bin/env python3
"""Bring-up probe for the tcgen05 kernels: runs small structured cases through the C ABI
with impl=2, compares the s32 accumulators with the oracle, and dumps raw results to
gpurun_out/tc_probe.npz so mismatches can be analysed offline. Dev tool, not a test."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from int8inferenceengine_b200 import _lib  # noqa: E402
from oracle import port  # noqa: E402
from gpu_utils import make_layer, u8_tensor_from_nchw  # noqa: E402

L = _lib.load()
dump = {}
summary = []


def report(tag, got, exp):
    bad = got != exp
    nbad = int(bad.sum())
    line = f"{tag}: shape {exp.shape} mismatches {nbad}/{exp.size}"
    if nbad:
        idx = np.argwhere(bad)
        line += f" first {idx[:6].tolist()} rows_bad {np.unique(idx[:, 0]).size} cols_bad {np.unique(idx[:, -1]).size}"
        dump[tag + "_got"] = got
        dump[tag + "_exp"] = exp
    err = L.i8ie_debug_tc_error(1)
    line += f" tc_error={err}"
    print(line, flush=True)
    summary.append(line)
    return nbad == 0 and err == 0


def fc_case(m, k, n, seed=0, impl=2, ident=False):
    rng = np.random.default_rng(seed)
    if ident:
        w = np.zeros((n, k), np.float32)
        for j in range(n):
            w[j, j % k] = 1.0
        w[0, 0] = 1.0
        b = np.zeros(n, np.float32)
        b[0] = -1.0  # forces min/max = -1/1 so scale = 2/127 and qw = +-63
    else:
        w = rng.uniform(-0.2, 0.2, size=(n, k)).astype(np.float32)
        b = rng.uniform(-0.05, 0.05, size=(n,)).astype(np.float32)
    q = rng.integers(0, 256, size=(m, k), dtype=np.uint8)
    layer = make_layer("fc", w, b, (np.float32(0.2), 128))
    qw, qb, ws = port.quantize_weight(w, b)
    exp, exp_acc = port.linear_u8(q, qw, qb, np.float32(0.03), 77, ws, np.float32(0.2), 128, want_acc=True)
    acc = torch.full((m * n,), -777, dtype=torch.int32, device="cuda")
    out = layer._forward_u8(u8_tensor_from_nchw(q, 0.03, 77), acc_out=acc, impl=impl)
    torch.cuda.synchronize()
    ok = report(f"fc_m{m}_k{k}_n{n}{'_id' if ident else ''}_acc", acc.cpu().numpy().reshape(m, n), exp_acc)
    ok &= report(f"fc_m{m}_k{k}_n{n}{'_id' if ident else ''}_u8", out.numpy(), exp)
    return ok


def conv_case(n, c, h, w_, kc, k, s, p, seed=0, impl=2):
    rng = np.random.default_rng(seed)
    a = np.sqrt(6.0 / (c * k * k))
    w = rng.uniform(-a, a, size=(kc, c, k, k)).astype(np.float32)
    b = rng.uniform(-0.05, 0.05, size=(kc,)).astype(np.float32)
    q = rng.integers(0, 256, size=(n, c, h, w_), dtype=np.uint8)
    layer = make_layer("conv", w, b, (np.float32(0.06), 120), s, p)
    qw, qb, ws = port.quantize_weight(w, b)
    exp, exp_acc = port.conv2d_u8(q, qw, qb, s, p, np.float32(0.03), 77, ws, np.float32(0.06), 120, want_acc=True)
    oh, ow = exp.shape[2], exp.shape[3]
    acc = torch.full((n * oh * ow * kc,), -777, dtype=torch.int32, device="cuda")
    out = layer._forward_u8(u8_tensor_from_nchw(q, 0.03, 77), acc_out=acc, impl=impl)
    torch.cuda.synchronize()
    tag = f"conv_n{n}c{c}h{h}w{w_}kc{kc}k{k}s{s}p{p}"
    ok = report(tag + "_acc", acc.cpu().numpy().reshape(n, oh * ow, kc), exp_acc)
    ok &= report(tag + "_u8", out.numpy(), exp)
    return ok


if __name__ == "__main__":
    stage = sys.argv[1] if len(sys.argv) > 1 else "all"
    ok = True
    if stage in ("fc", "all"):
        ok &= fc_case(128, 32, 32, ident=True)     # one MMA, identity-ish weights
        ok &= fc_case(128, 32, 32)                 # one MMA
        ok &= fc_case(128, 128, 32)                # 4 MMAs in one K block (descriptor advance)
        ok &= fc_case(128, 512, 64)                # several K blocks (pipeline wrap: 4 stages)
        ok &= fc_case(100, 784, 10)                # ragged M/K/N (BASELINE config 1)
        ok &= fc_case(300, 1000, 200)              # multiple M tiles, BN=128/256 paths
        ok &= fc_case(130, 4096, 4096)
    if stage in ("conv", "all"):
        ok &= conv_case(2, 32, 8, 8, 32, 1, 1, 0)      # 1x1: im2col walk only
        ok &= conv_case(2, 32, 8, 8, 32, 3, 1, 0)      # taps, no padding
        ok &= conv_case(2, 32, 8, 8, 32, 3, 1, 1)      # zero-fill + border correction
        ok &= conv_case(3, 64, 13, 13, 96, 3, 1, 1)    # BK=64
        ok &= conv_case(2, 96, 27, 27, 256, 5, 1, 2)   # conv2 shape, BK=32
        ok &= conv_case(2, 256, 13, 13, 384, 3, 1, 1)  # conv3 shape, BK=128, BN=192
        ok &= conv_case(2, 128, 9, 11, 64, 3, 2, 1)    # stride 2, non-square
        ok &= conv_case(1, 32, 20, 20, 40, 5, 3, 2)    # stride 3, pad 2
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if dump:
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", "tc_probe.npz"), **dump)
    open(os.path.join(ROOT, "gpurun_out", "tc_probe.txt"), "w").write("\n".join(summary) + "\n")
    print("TC PROBE", "OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
