#!/usr/bin/env python3
"""Dev tool: one-screen summary of a bench.py JSON line (python tools/bench_brief.py file.json ...)."""
import json
import sys

for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f"{path}: N={d['n_gpus']} {d['config']['workload']}  {d['ms_per_step']:.4f} ms/step  {d['value']:.0f} img/s  "
          f"e2e {d['e2e']['value']:.0f}  pageable {d.get('e2e_pageable', {}).get('value', 0):.0f}  "
          f"launches/step {d['gpu_launches'] // d['steps']}  parity {d['parity'] and d['parity']['bit_identical']}  "
          f"frac {r.get('frac', 0):.3f} share {r.get('share_of_step')}  clocks {d['clocks'].get('sm_mhz')} {d['clocks'].get('reasons')}")
    print("   " + "  ".join(f"{l['layer']} {l['us']:.1f}" for l in d.get("layers", [])))
    if d.get("per_gpu_alone_ms") is not None:
        print(f"   alone {d['per_gpu_alone_ms']:.4f} ms  exchange {d['exchange_ms']:.4f} ms")
