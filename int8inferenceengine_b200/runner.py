"""Builds `i8ie.Module`s for the BASELINE topologies and runs them the way the
reference's notebooks do (prepare -> one calibration batch -> convert -> forward)."""
from __future__ import annotations

import numpy as np

from . import api
from .workloads import TOPOLOGIES


class TopologyModule(api.Module):
    """An i8ie.Module whose forward() is generated from a workloads.TOPOLOGIES op list —
    equivalent to the hand-written MyNet classes of the reference notebooks."""

    def __init__(self, topology):
        super().__init__()
        self._ops = TOPOLOGIES[topology]["ops"]
        for op in self._ops:
            if op[0] == "conv":
                _, name, cin, cout, k, s, p = op
                self.__dict__[name] = api.Conv2d(cin, cout, kernel_size=k, stride=s, padding=p)
            elif op[0] == "fc":
                _, name, cin, cout = op
                self.__dict__[name] = api.Linear(cin, cout)
        self.record = None   # when a list, forward() appends (tag, Tensor) per op

    def forward(self, x):
        rec = self.record
        for op in self._ops:
            if op[0] in ("conv", "fc"):
                x = self.__dict__[op[1]](x)
                tag = op[1]
            elif op[0] == "relu":
                x = api.relu(x)
                tag = "relu"
            elif op[0] == "pool":
                x = api.max_pool2d(x, op[1], op[2])
                tag = "pool"
            else:
                x = x.reshape(-1, op[1])
                continue
            if rec is not None:
                rec.append((tag, x))
        return x

    def layers(self):
        return {op[1]: self.__dict__[op[1]] for op in self._ops if op[0] in ("conv", "fc")}

    def set_qparams(self, qparams):
        """Inject per-layer (scale, zero_point) — e.g. those the reference calibrated."""
        for name, (s, z) in qparams.items():
            self.__dict__[name].layer.set_qparams(s, z)


def build_module(topology, state_dict, qparams=None, calib=None, per_channel=False):
    """Module with weights loaded and converted. Either inject `qparams`
    ({layer: (scale, zp)}) or pass a calibration batch `calib` (numpy NCHW f32).
    per_channel=True (extension, not in the reference): per-output-channel weight scales."""
    m = TopologyModule(topology)
    m.load(state_dict)
    if per_channel:
        for L in m.layers().values():
            L.layer.per_channel = True
    if qparams is not None:
        m.set_qparams(qparams)
    elif calib is not None:
        m.prepare()
        m(api.tensor(np.asarray(calib)))
    m.convert()
    return m
