"""Synthetic, seeded workloads for the INT8 hot path (BASELINE.json configs).

Pure numpy, no torch, no oracle: shared by the product benches, the parity tests
and the golden-vector generator so that the B200 path, the C oracle and the
compiled reference all see byte-identical inputs.

Topologies are the ones the reference's notebooks / tests define:
  fc_mnist     sample/notebooks/Fully_Connected_mnist.ipynb  (MyNet: Linear 784->10)
  simple_conv  sample/notebooks/Simple_Convolution_cifar10.ipynb (3 conv k5 + fc)
  alexnet      sample/notebooks/AlexNet_cifar10_resize224.ipynb
  lenet        unittest/test_quantized_layer.py:26-42
A topology is a list of ops:
  ("conv", name, cin, cout, k, stride, pad) | ("fc", name, cin, cout) |
  ("relu",) | ("pool", k, stride) | ("flatten", features)
"""
from __future__ import annotations

import numpy as np

INPUT_SCALE = 0.025   # i8ie/module.py:20 (hard-coded input quantisation)
INPUT_ZP = 127

TOPOLOGIES = {
    "fc_mnist": {
        "input": (1, 28, 28), "range": (0.0, 1.0),
        "ops": [("flatten", 784), ("fc", "fc", 784, 10)],
    },
    "simple_conv": {
        "input": (3, 32, 32), "range": (-2.1, 2.6),
        "ops": [("conv", "conv1", 3, 20, 5, 1, 0), ("relu",),
                ("conv", "conv2", 20, 50, 5, 1, 0), ("relu",), ("pool", 2, 2),
                ("conv", "conv3", 50, 120, 5, 1, 0), ("relu",),
                ("flatten", 960 * 8), ("fc", "fc", 960 * 8, 10)],
    },
    "alexnet": {
        "input": (3, 224, 224), "range": (-2.1, 2.6),
        "ops": [("conv", "conv1", 3, 96, 11, 4, 2), ("relu",), ("pool", 3, 2),
                ("conv", "conv2", 96, 256, 5, 1, 2), ("relu",), ("pool", 3, 2),
                ("conv", "conv3", 256, 384, 3, 1, 1), ("relu",),
                ("conv", "conv4", 384, 384, 3, 1, 1), ("relu",),
                ("conv", "conv5", 384, 256, 3, 1, 1), ("relu",), ("pool", 3, 2),
                ("flatten", 6 * 6 * 256),
                ("fc", "fc1", 6 * 6 * 256, 4096), ("relu",),
                ("fc", "fc2", 4096, 4096), ("relu",),
                ("fc", "fc3", 4096, 10)],
    },
    "lenet": {
        "input": (1, 28, 28), "range": (-2.0, 2.0),
        "ops": [("conv", "conv1", 1, 20, 5, 1, 0), ("pool", 2, 2),
                ("conv", "conv2", 20, 50, 5, 1, 0), ("pool", 2, 2),
                ("flatten", 800), ("fc", "fc1", 800, 500), ("relu",),
                ("fc", "fc2", 500, 10)],
    },
    # a small AlexNet-shaped net (padding, stride 4, overlapping pools, C=3 stem)
    # that the scalar oracle finishes in well under a second
    "mini_alex": {
        "input": (3, 67, 67), "range": (-2.1, 2.6),
        "ops": [("conv", "conv1", 3, 32, 11, 4, 2), ("relu",), ("pool", 3, 2),
                ("conv", "conv2", 32, 64, 5, 1, 2), ("relu",), ("pool", 3, 2),
                ("conv", "conv3", 64, 96, 3, 1, 1), ("relu",),
                ("flatten", 96 * 3 * 3), ("fc", "fc1", 96 * 3 * 3, 128), ("relu",),
                ("fc", "fc2", 128, 10)],
    },
}

MAC_PER_IMAGE = {  # BASELINE.md §2 (un-padded reference dims)
    "fc_mnist": 7_840, "simple_conv": 25_252_800, "alexnet": 1_131_201_056,
}


def layer_names(topology: str):
    return [op[1] for op in TOPOLOGIES[topology]["ops"] if op[0] in ("conv", "fc")]


def make_weights(topology: str, seed: int = 0):
    """He-uniform weights U(-a,a), a=sqrt(6/fan_in); bias U(-0.05,0.05). Returns a
    state_dict {"name.weight": f32, "name.bias": f32} in the key format
    i8ie.Module.load expects (i8ie/module.py:10-16)."""
    rng = np.random.default_rng(seed)
    sd = {}
    for op in TOPOLOGIES[topology]["ops"]:
        if op[0] == "conv":
            _, name, cin, cout, k, _, _ = op
            a = np.sqrt(6.0 / (cin * k * k))
            sd[f"{name}.weight"] = rng.uniform(-a, a, size=(cout, cin, k, k)).astype(np.float32)
        elif op[0] == "fc":
            _, name, cin, cout = op
            a = np.sqrt(6.0 / cin)
            sd[f"{name}.weight"] = rng.uniform(-a, a, size=(cout, cin)).astype(np.float32)
        else:
            continue
        sd[f"{name}.bias"] = rng.uniform(-0.05, 0.05, size=(cout,)).astype(np.float32)
    return sd


def make_images(topology: str, batch: int, seed: int):
    """Seeded synthetic images in the range the notebooks' normalisation yields;
    inside the representable window of 0.025/127 so the unclamped input
    quantise (quantize_utils.cc:49) never wraps."""
    t = TOPOLOGIES[topology]
    lo, hi = t["range"]
    rng = np.random.default_rng(seed)
    return rng.uniform(lo, hi, size=(batch,) + t["input"]).astype(np.float32)


def conv_out_hw(h, w, k, stride, pad):
    return (h - k + 2 * pad) // stride + 1, (w - k + 2 * pad) // stride + 1
