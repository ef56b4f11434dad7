#!/usr/bin/env python3
"""Turn ncu output (gpurun_out/, scratch) into the small text summaries kept under profiles/.

  python tools/ncu_summary.py launches gpurun_out/r01_launches.csv  > profiles/r01_launches.md
  python tools/ncu_summary.py report   gpurun_out/prof.ncu-rep      > profiles/r01_top_kernel.md

`launches` aggregates a `--metrics gpu__time_duration.sum --csv` launch list by kernel name
(count, total, share of all profiled GPU time).  `report` prints, per profiled launch of an
`ncu --set full` capture, the handful of metrics the roofline discussion needs.
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg",
    "smsp__cycles_active.avg",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem",
]


def short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("(anonymous namespace)::", "")
    return name[:110]


def launches(path: str) -> None:
    per = OrderedDict()   # launch ID -> {name, grid, block, ns, dram}
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        e = per.setdefault(r["ID"], {"name": short(r["Kernel Name"]), "grid": r["Grid Size"],
                                     "block": r["Block Size"], "ns": 0.0, "dram": None})
        v = float(r["Metric Value"].replace(",", ""))
        u = (r.get("Metric Unit") or "").lower()
        if r.get("Metric Name") == "gpu__time_duration.sum":
            e["ns"] = v * {"us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(u, 1)
        elif r.get("Metric Name") in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            e["dram"] = (e["dram"] or 0.0) + v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    rows = list(per.values())
    total = sum(r["ns"] for r in rows)
    agg = OrderedDict()
    for r in rows:
        c = agg.setdefault((r["name"], r["grid"], r["block"]), [0, 0.0, 0.0, r["dram"] is not None])
        c[0] += 1
        c[1] += r["ns"]
        c[2] += r["dram"] or 0.0
    print(f"# launch list summary: {path}")
    print(f"# {len(rows)} launches, {total / 1e3:.1f} us of GPU time "
          "(ncu per-launch times are serialised, one kernel at a time: compare shares)")
    print("| share % | launches | avg us | avg DRAM MB (rd+wr) | kernel | grid | block |")
    print("|---|---|---|---|---|---|---|")
    for (n, g, b), (c, v, d, has) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        dram = f"{d / c / 1e6:8.2f}" if has else "-"
        print(f"| {100 * v / total:5.1f} | {c} | {v / c / 1e3:8.2f} | {dram} | `{n}` | {g} | {b} |")


def report(path: str) -> None:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"],
                         check=True, capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rd[0], rd[1], rd[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary: {path}")
    for r in body:
        print(f"\n## launch {r[col['ID']]}: `{short(r[col['Kernel Name']])}`  "
              f"grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        for k in KEEP:
            if k in col:
                print(f"- {k} = {r[col[k]]} {units[col[k]]}")
        rd_b = col.get("dram__bytes_read.sum")
        wr_b = col.get("dram__bytes_write.sum")
        if rd_b is not None and wr_b is not None:
            def tobytes(i):
                v = float(r[i].replace(",", ""))
                u = units[i].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            print(f"- traffic (dram read+write) = {tobytes(rd_b) + tobytes(wr_b):.0f} bytes")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
