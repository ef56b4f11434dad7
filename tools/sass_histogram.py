#!/usr/bin/env python3
"""SASS opcode histogram of libi8ie_sm100.so per kernel (cuobjdump -sass): the evidence that the hot path is
hand-written tcgen05 / TMEM / TMA code (UTC*MMA, LDTM, UTMALDG, UBLKCP ...). Writes a markdown table.
  python tools/sass_histogram.py > profiles/r02_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "int8inferenceengine_b200", "lib", "libi8ie_sm100.so")
KEY = ["UTCIMMA", "UTCIMMA.2CTA", "UTCBAR", "LDTM", "UTMALDG", "UTMAPF", "UBLKCP", "UBLKPF", "SYNCS", "FMUL2", "FFMA2", "FADD2",
       "F2IP", "I2FP", "IDP.4A", "REDUX", "ATOMG", "LDG", "STG", "LDS", "STS", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern = None
    hist = collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = m.group(1)
            hist[kern] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and kern:
            op = m.group(1)
            hist[kern][op] += 1
            hist[kern]["_total"] += 1
    print("# SASS opcode histogram of libi8ie_sm100.so (cuobjdump -sass, sm_100a)\n")
    print("Counts of static instructions per kernel. `UTCIMMA` = tcgen05.mma kind::i8 (`.2CTA` = cta_group::2), `LDTM` = tcgen05.ld,")
    print("`UTMALDG` = cp.async.bulk.tensor (TMA, incl. `.IM2COL`), `UBLKCP` = cp.async.bulk, `SYNCS` = mbarrier ops,")
    print("`FMUL2/FFMA2/FADD2` = packed fp32x2, `F2IP` = saturating float->u8 pack, `IDP.4A` = dp4a.\n")
    names = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
    pretty = {}
    for mangled, d in zip(hist, names):
        d = d.replace("(anonymous namespace)::", "").replace("i8ie::", "")
        d = re.sub(r"^void ", "", d)
        pretty[mangled] = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", d)
    total = collections.Counter()
    rows = []
    for k, c in hist.items():
        agg = collections.Counter()
        for op, n in c.items():
            if op == "_total":
                continue
            base = op
            for key in sorted(KEY, key=len, reverse=True):
                if op == key or op.startswith(key + ".") or (key == "UTCIMMA.2CTA" and "UTCIMMA" in op and "2CTA" in op):
                    base = key
                    break
            else:
                continue
            if base == "UTCIMMA" and "2CTA" in op:
                base = "UTCIMMA.2CTA"
            agg[base] += n
        if any(agg[x] for x in ("UTCIMMA", "UTCIMMA.2CTA", "LDTM", "UTMALDG", "UBLKCP", "IDP.4A")) or c["_total"] > 400:
            rows.append((k, c["_total"], agg))
        total.update(agg)
    cols = [k for k in KEY if total[k]]
    print("| kernel | instrs | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for k, n, agg in sorted(rows, key=lambda r: -r[1]):
        print(f"| `{pretty[k][:90]}` | {n} | " + " | ".join(str(agg[c]) if agg[c] else "" for c in cols) + " |")
    print(f"\n| **whole library** | | " + " | ".join(str(total[c]) for c in cols) + " |")


if __name__ == "__main__":
    sys.exit(main())
