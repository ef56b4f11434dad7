"""int8inferenceengine_b200 — the B200 (sm_100a) backend for the i8ie INT8 inference hot path.

  api        the unchanged `i8ie` user surface (also importable as top-level `i8ie`)
  backend    backend object protocol (stand-in for the reference's `_CXX_i8ie`)
  _lib       ctypes binding of lib/libi8ie_sm100.so (C ABI: include/i8ie_sm100.h)
  build      in-tree nvcc build of that library
  workloads  seeded synthetic workloads / topologies of the BASELINE configs
"""
__version__ = "0.1.0"
