#!/usr/bin/env python
"""Summarise an ncu launch list (tools/launch_list.sh): per kernel and per (kernel, grid) time and DRAM bytes of the LAST pass.
usage: python tools/launch_summary.py gpurun_out/<tag>_launches_b32.csv [nlast]"""
import collections
import csv
import sys


def load(path):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    cur = {}
    for r in rd:
        if len(r) < len(hdr) or not r[0].isdigit():
            continue
        k = r[ix["ID"]]
        d = cur.setdefault(k, {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
        try:
            d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            pass
        d["unit_" + r[ix["Metric Name"]]] = r[ix["Metric Unit"]]
    for k in sorted(cur, key=int):
        rows.append(cur[k])
    return rows


def scale(d, m):
    u = d.get("unit_" + m, "")
    v = d.get(m, 0.0)
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


def main():
    rows = load(sys.argv[1])
    # the last pass starts at the last geom_fields launch
    starts = [i for i, r in enumerate(rows) if "geom_fields" in r["name"]]
    rows = rows[starts[-1]:] if starts else rows
    agg = collections.OrderedDict()
    for r in rows:
        nm = r["name"].split("(")[0].replace("void ", "")
        key = nm if len(sys.argv) < 3 else (nm, r["grid"])
        a = agg.setdefault(key, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += scale(r, "gpu__time_duration.sum"); a[2] += scale(r, "dram__bytes_read.sum"); a[3] += scale(r, "dram__bytes_write.sum")
    tot = sum(a[1] for a in agg.values())
    print("kernel,launches,time_ms,time_share,dram_read_GB,dram_write_GB")
    for k, a in agg.items():
        print(f"{k},{a[0]},{a[1]:.3f},{a[1] / tot:.4f},{a[2] / 1e9:.3f},{a[3] / 1e9:.3f}")
    print(f"TOTAL,{sum(a[0] for a in agg.values())},{tot:.3f},1.0,{sum(a[2] for a in agg.values()) / 1e9:.3f},{sum(a[3] for a in agg.values()) / 1e9:.3f}")


if __name__ == "__main__":
    main()
