"""Loader for the compiled reference (oracle/_ref/_CXX_i8ie*.so). TEST INFRASTRUCTURE ONLY.

The .so is the reference's own src/*.cc built unmodified by oracle/Makefile
(`make ref`) against the stand-in mkl.h. We bind its pybind11 module directly
(src/pybind11.cc:37-55) and do NOT import the reference's Python package, so
nothing under /root/reference is needed at run time (it does not exist on the
GPU box; the built .so travels with the snapshot).
"""
from __future__ import annotations

import glob
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_mod = None


def so_path():
    hits = sorted(glob.glob(os.path.join(_HERE, "_ref", "_CXX_i8ie*.so")))
    return hits[0] if hits else None


def available():
    return so_path() is not None


def module():
    """The reference's pybind11 module `_CXX_i8ie` (raises if it was never built)."""
    global _mod
    if _mod is None:
        p = so_path()
        if p is None:
            raise RuntimeError("oracle/_ref is not built (run `make -C oracle ref` where /root/reference exists)")
        spec = importlib.util.spec_from_file_location("_CXX_i8ie", p)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _mod = m
    return _mod
