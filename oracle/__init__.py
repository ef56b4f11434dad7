"""Parity oracle for the i8ie INT8 hot path. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product (int8inferenceengine_b200) never does.

  oracle.port  ctypes binding of oracle/liboracle_i8ie.so (C restatement, i8ie_oracle.c)
  oracle.ref   loader + driver for the compiled reference (oracle/_ref/_CXX_i8ie*.so)
  oracle.models  whole-topology runners over either of the above
"""
