"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every
symbol include/i8ie_sm100.h declares; the Python surface mirrors the reference's; the
product path refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "i8ie_sm100.h")).read()
    return sorted(set(re.findall(r"I8IE_API[^;(]*?\b(i8ie_\w+)\s*\(", src)))


def test_header_symbols_match_binding_table():
    from int8inferenceengine_b200 import _lib
    assert _declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_loads_and_exports_every_symbol():
    from int8inferenceengine_b200 import _lib
    L = _lib.load()
    for s in _declared_symbols():
        assert hasattr(L, s), s
    assert b"sm_100a" in L.i8ie_version()


def test_host_side_entry_points():
    """The two host functions of the ABI need no GPU: layer.cc:6-26 and calibrator.cc:28-35."""
    import numpy as np
    from int8inferenceengine_b200 import _lib
    from oracle import port
    L = _lib.load()
    rng = np.random.default_rng(0)
    w = rng.uniform(-0.3, 0.3, size=(37, 11)).astype(np.float32)
    b = rng.uniform(-0.05, 0.05, size=(37,)).astype(np.float32)
    qw = np.empty(w.shape, np.int8)
    qb = np.empty(b.shape, np.int8)
    sc = ctypes.c_float()
    assert L.i8ie_quantize_weight_host(w.ctypes.data, w.size, b.ctypes.data, b.size, qw.ctypes.data,
                                       qb.ctypes.data, ctypes.byref(sc)) == 0
    eqw, eqb, es = port.quantize_weight(w, b)
    assert np.array_equal(qw, eqw) and np.array_equal(qb, eqb) and np.float32(sc.value) == es
    for mn, mx in [(-3.0, 5.0), (0.5, 4.0), (-4.0, -0.5), (0.0, 0.0), (-1e-3, 1e-3), (-300.0, 900.0)]:
        s, z = ctypes.c_float(), ctypes.c_uint8()
        assert L.i8ie_range_from_minmax_host(mn, mx, ctypes.byref(s), ctypes.byref(z)) == 0
        assert (np.float32(s.value), int(z.value)) == port.get_range_minmax(mn, mx)


def test_python_surface_matches_reference():
    import i8ie
    # i8ie/__init__.py:6-10 of the reference
    assert sorted(i8ie.__all__) == sorted(["tensor", "argmax", "relu", "max_pool2d", "Linear", "Conv2d",
                                           "Tensor", "quantize", "dequantize"])
    for name in i8ie.__all__ + ["Module"]:
        assert hasattr(i8ie, name)
    for meth in ["load", "prepare", "convert", "__call__"]:
        assert hasattr(i8ie.Module, meth)
    for meth in ["load_weight", "load_bias", "prepare", "convert", "__call__"]:
        assert hasattr(i8ie.Linear, meth) and hasattr(i8ie.Conv2d, meth)
    for attr in ["reshape", "numpy", "sum", "shape", "scale", "zero_point"]:
        assert hasattr(i8ie.Tensor, attr)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import numpy as np
    import i8ie
    from int8inferenceengine_b200._lib import I8ieError
    with pytest.raises(I8ieError):
        i8ie.tensor(np.zeros((2, 2), np.float32))
    with pytest.raises(I8ieError):
        i8ie.Linear(4, 4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "int8inferenceengine_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "liboracle" not in txt, f
