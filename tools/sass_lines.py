#!/usr/bin/env python3
"""Dev tool: per-source-line instruction counts / stall samples of one kernel.

  ncu -i rep.ncu-rep --page source --csv > sass.csv
  cuobjdump -xelf all lib/tc_gemm.o && nvdisasm --print-line-info -c tc_gemm.sm_100a.cubin > dis.txt
  python tools/sass_lines.py sass.csv dis.txt <mangled-kernel-substring> [top]

The ncu SASS page has no line numbers in CSV form; nvdisasm's `//## File "...", line N` markers of
the same cubin are matched to it by instruction order."""
import csv
import re
import sys
from collections import defaultdict


def main():
    sass_csv, dis, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
    rows = list(csv.reader(open(sass_csv)))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = []   # (file:line [inlined chain]) per instruction of the kernel, in order
    inside, cur = False, "?"
    for ln in open(dis):
        if ln.startswith("//---") and ".text." in ln:
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = f"{m.group(1).split('/')[-1]}:{m.group(2)}"
            if "inlined at" in m.group(3):
                m2 = re.search(r'inlined at "([^"]+)", line (\d+)', m.group(3))
                if m2:
                    cur += f" <- {m2.group(1).split('/')[-1]}:{m2.group(2)}"
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines.append(cur)
    if len(lines) != len(data):
        print(f"warning: {len(lines)} disassembled vs {len(data)} profiled instructions", file=sys.stderr)
    agg = defaultdict(lambda: [0, 0])
    tot = 0
    for loc, r in zip(lines, data):
        n = int(r[ix["Instructions Executed"]])
        agg[loc][0] += n
        agg[loc][1] += int(r[ix["# Samples"]])
        tot += n
    print(f"total warp instructions {tot}")
    for loc, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{n:10d} {100 * n / tot:5.1f}%  samples {s:5d}  {loc}")


if __name__ == "__main__":
    main()
