#!/usr/bin/env python3
"""Dev tool: AlexNet conv2-5 at several batch sizes with the pair kernel's tile forced to each candidate
(I8IE_TC_BN = 128 / 192 / 256 / 384 / 512 where 512 = 256 columns x two accumulators) next to the cost
model's own pick — calibrates tc_pick_bn_pair. One subprocess per variant (the override is read at plan creation)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = """
import sys; sys.path.insert(0, %r); sys.path.insert(0, %r + '/tools')
import layer_bench
for b in %s:
    for name, us, tops, impl in layer_bench.run(batch=b, layers=%r, reps=10, iters=5, quiet=True):
        print(b, name, round(us, 2), round(tops, 1), flush=True)
"""


def main():
    batches = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "100,125,250,500").split(",")]
    layers = sys.argv[2] if len(sys.argv) > 2 else "conv2,conv3,conv4,conv5"
    res = {}
    variants = ["auto", "128", "192", "256", "384", "512"]
    for v in variants:
        env = dict(os.environ)
        if v != "auto":
            env["I8IE_TC_BN"] = v
        r = subprocess.run([sys.executable, "-c", CODE % (ROOT, ROOT, batches, layers)], env=env, capture_output=True, text=True,
                           timeout=600)
        for ln in r.stdout.splitlines():
            b, name, us, tops = ln.split()
            res[(int(b), name, v)] = float(us)
        if r.returncode != 0:
            print(f"variant {v} failed: {r.stderr[-400:]}", file=sys.stderr)
    names = layers.split(",")
    print("| batch | layer | " + " | ".join(variants) + " |")
    print("|---|---|" + "---|" * len(variants))
    for b in batches:
        for n in names:
            print(f"| {b} | {n} | " + " | ".join(f"{res.get((b, n, v), float('nan')):.1f}" for v in variants) + " |")


if __name__ == "__main__":
    main()
