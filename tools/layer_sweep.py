#!/usr/bin/env python3
"""BASELINE config 5: every AlexNet conv / fc layer standalone (u8 activations U{0..255}, s8 weights),
batch 1 .. 4096, as TOP/s on the reference's un-padded dims next to the INT8 tensor roofline
(the measured kind::i8 peak of profiles/i8_peak.json, else 2 x the measured cuBLAS bf16 rate). Writes a markdown table to stdout."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import layer_bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256,512,1024,2048,4096")
    ap.add_argument("--layers", default="")
    args = ap.parse_args()
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    bf16 = json.load(open(pk))["bf16_tflops"] if os.path.exists(pk) else 1590.0
    peak = 2 * bf16
    pi8 = os.path.join(ROOT, "profiles", "i8_peak.json")
    peak_src = f"2 x measured bf16 {bf16} TFLOP/s"
    if os.path.exists(pi8):   # kind::i8 tensor peak measured with the UTCIMMA-only microbenchmark
        peak = json.load(open(pi8))["i8_tops_burst"]
        peak_src = "measured kind::i8 peak, tools/ubench/i8_peak.cu"
    batches = [int(b) for b in args.batches.split(",")]
    table = {}
    names = []
    for b in batches:
        reps = 20 if b <= 256 else (8 if b <= 1024 else 3)
        for name, us, tops, impl in layer_bench.run(batch=b, layers=args.layers, reps=reps, iters=5, quiet=True):
            table[(name, b)] = (us, tops)
            if name not in names:
                names.append(name)
        print(f"# batch {b} done", file=sys.stderr, flush=True)
    print(f"# AlexNet layer sweep on one B200: TOP/s (un-padded 2*M*N*K) / % of {peak:.0f} TOP/s "
          f"({peak_src}) / us per launch")
    print("| batch | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for b in batches:
        cells = []
        for n in names:
            us, tops = table[(n, b)]
            cells.append(f"{tops:.0f} ({100 * tops / peak:.0f} %) {us:.1f} us")
        print(f"| {b} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
