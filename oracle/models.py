"""Whole-topology runners over the oracle. TEST INFRASTRUCTURE ONLY.

PortModel  runs a topology (int8inferenceengine_b200.workloads.TOPOLOGIES) through the C
           restatement (oracle.port), following i8ie/module.py:18-24 for the
           quantise -> forward -> dequantise wrapping.
RefModel   runs the same topology through the compiled reference's pybind11
           objects (oracle.ref), layer by layer, the way
           unittest/test_quantized_layer.py:63-95 drives it.
Both record, per op, the u8 activation (NCHW) and its (scale, zero_point).
"""
from __future__ import annotations

import numpy as np

from int8inferenceengine_b200.workloads import INPUT_SCALE, INPUT_ZP, TOPOLOGIES

from . import port, ref


class PortModel:
    def __init__(self, topology, state_dict):
        self.ops = TOPOLOGIES[topology]["ops"]
        self.sd = {k: np.ascontiguousarray(v, np.float32) for k, v in state_dict.items()}
        self.q = {}          # name -> (qw, qb, w_scale)
        self.qparams = {}    # name -> (scale f32, zp int)

    # -- fp32 side (calibration feeder; conv2d.cc:63-98, fully_connected.cc:5-21) --
    def forward_fp32(self, x, record=False):
        rec = {}
        x = np.ascontiguousarray(x, np.float32)
        for op in self.ops:
            if op[0] == "conv":
                _, name, _, _, _, s, p = op
                x = port.conv2d_f32(x, self.sd[f"{name}.weight"], self.sd[f"{name}.bias"], s, p)
                rec[name] = x
            elif op[0] == "fc":
                name = op[1]
                x = port.linear_f32(x, self.sd[f"{name}.weight"], self.sd[f"{name}.bias"])
                rec[name] = x
            elif op[0] == "relu":
                x = port.relu_f32(x)
            elif op[0] == "pool":
                x = port.max_pool2d_f32(x, op[1], op[2])
            elif op[0] == "flatten":
                x = x.reshape(-1, op[1])
        return (x, rec) if record else x

    def calibrate_minmax(self, x_cal):
        """Per-layer (scale, zp) from the true min/max of each layer's fp32
        pre-activation output through calibrator.cc:28-35's scalar arithmetic."""
        _, rec = self.forward_fp32(x_cal, record=True)
        return {n: port.get_range_minmax(v.min(), v.max()) for n, v in rec.items()}

    def convert(self, qparams):
        """layer.cc:36-54 with the calibrated ranges supplied (name -> (scale, zp))."""
        self.qparams = {k: (np.float32(v[0]), int(v[1])) for k, v in qparams.items()}
        for op in self.ops:
            if op[0] in ("conv", "fc"):
                n = op[1]
                self.q[n] = port.quantize_weight(self.sd[f"{n}.weight"], self.sd[f"{n}.bias"])

    def forward_int8(self, x, record=False):
        """i8ie/module.py:18-24: quantise(0.025,127) -> layers -> dequantise."""
        recs = []
        q = port.quantize(x, INPUT_SCALE, INPUT_ZP)
        scale, zp = np.float32(INPUT_SCALE), INPUT_ZP
        recs.append(("input", q, scale, zp))
        for op in self.ops:
            if op[0] == "conv":
                _, name, _, _, _, s, p = op
                qw, qb, ws = self.q[name]
                os_, oz = self.qparams[name]
                q = port.conv2d_u8(q, qw, qb, s, p, scale, zp, ws, os_, oz)
                scale, zp = os_, oz
                recs.append((name, q, scale, zp))
            elif op[0] == "fc":
                name = op[1]
                qw, qb, ws = self.q[name]
                os_, oz = self.qparams[name]
                q = port.linear_u8(q, qw, qb, scale, zp, ws, os_, oz)
                scale, zp = os_, oz
                recs.append((name, q, scale, zp))
            elif op[0] == "relu":
                q = port.relu_u8(q, zp)
                recs.append(("relu", q, scale, zp))
            elif op[0] == "pool":
                q = port.max_pool2d_u8(q, op[1], op[2])
                recs.append(("pool", q, scale, zp))
            elif op[0] == "flatten":
                q = q.reshape(-1, op[1])
        logits = port.dequantize(q, scale, zp)
        return (logits, recs) if record else logits


class RefModel:
    """The compiled reference, driven through `_CXX_i8ie` (src/pybind11.cc:37-55)."""

    def __init__(self, topology, state_dict):
        self.m = ref.module()
        self.ops = TOPOLOGIES[topology]["ops"]
        self.layers = {}
        for op in self.ops:
            if op[0] == "conv":
                _, name, cin, cout, k, s, p = op
                L = self.m.Conv2d(cin, cout, k, s, p)
            elif op[0] == "fc":
                _, name, cin, cout = op
                L = self.m.Linear(cin, cout)
            else:
                continue
            L.load_weight(np.ascontiguousarray(state_dict[f"{name}.weight"], np.float32))
            L.load_bias(np.ascontiguousarray(state_dict[f"{name}.bias"], np.float32))
            self.layers[name] = L
        self.is_quant = False

    def _forward(self, t, recs):
        m = self.m
        for op in self.ops:
            if op[0] in ("conv", "fc"):
                t = self.layers[op[1]](t)
                tag = op[1]
            elif op[0] == "relu":
                t = m.relu(t)
                tag = "relu"
            elif op[0] == "pool":
                t = m.max_pool2d(t, op[1], op[2])
                tag = "pool"
            elif op[0] == "flatten":
                t = t.reshape([-1, op[1]])
                continue
            if recs is not None:
                recs.append((tag, np.array(t.numpy(), copy=True), np.float32(t.scale()), int(t.zero_point())))
        return t

    def forward_fp32(self, x, record=False):
        assert not self.is_quant, "reference frees fp32 weights at convert (layer.cc:52-53)"
        recs = [] if record else None
        t = self._forward(self.m.tensor(np.ascontiguousarray(x, np.float32)), recs)
        out = np.array(t.numpy(), copy=True)
        return (out, recs) if record else out

    def calibrate(self, x_cal):
        """notebook protocol: prepare(); model(x_cal); convert()  (i8ie/module.py:26-35)."""
        for L in self.layers.values():
            L.prepare()
        self.forward_fp32(x_cal)
        for L in self.layers.values():
            L.convert()
        self.is_quant = True

    def forward_int8(self, x, record=False):
        assert self.is_quant
        m = self.m
        recs = [] if record else None
        t = m.quantize(m.tensor(np.ascontiguousarray(x, np.float32)), INPUT_SCALE, INPUT_ZP)
        if record:
            recs.append(("input", np.array(t.numpy(), copy=True), np.float32(t.scale()), int(t.zero_point())))
        t = self._forward(t, recs)
        logits = np.array(m.dequantize(t).numpy(), copy=True)
        return (logits, recs) if record else logits

    def qparams(self, x_probe):
        """Per-layer (scale, zp) as the reference calibrated them — read from layer
        outputs (conv2d.cc:112-113, fully_connected.cc:27-28)."""
        _, recs = self.forward_int8(x_probe, record=True)
        names = [op[1] for op in self.ops if op[0] in ("conv", "fc")]
        out = {}
        for tag, _, s, z in recs:
            if tag in names:
                out[tag] = (np.float32(s), int(z))
        return out
