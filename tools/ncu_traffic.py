#!/usr/bin/env python3
"""Summarises an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`
log of the first-stage launches into the JSON bench.py reports as roofline.traffic.
usage: ncu_traffic.py <launches.csv> <out.json> "<command that was profiled>" [config] """
import csv
import json
import os
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9,
        "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}


def main(src, dst, command, config=3):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    idc, kn, mn, mu, mv = (hdr.index(x) for x in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    per = {}
    for r in rows[1:]:
        d = per.setdefault(r[idc], {"kernel": r[kn].split("(")[0]})
        d[r[mn]] = float(r[mv].replace(",", "")) * UNIT.get(r[mu], 1)
    launches = [d for d in per.values() if "dram__bytes_read.sum" in d]
    n = len(launches)
    rd = sum(d["dram__bytes_read.sum"] for d in launches) / n
    wr = sum(d["dram__bytes_write.sum"] for d in launches) / n
    ns = sum(d.get("gpu__time_duration.sum", 0.0) for d in launches) / n
    out = {"kernel": "sw_u16_kernel", "launches_captured": n, "dram_bytes_read_per_launch": rd, "dram_bytes_written_per_launch": wr,
           "dram_bytes_per_launch": rd + wr, "ncu_ms_per_launch": ns / 1e6,
           "per_launch": [{"kernel": d["kernel"], "read": d["dram__bytes_read.sum"], "written": d["dram__bytes_write.sum"]} for d in launches],
           "config": int(config), "n_gpus": 1,
           "kernel_tag": open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oswald_b200", "build_tag.txt")).read().strip()
           if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oswald_b200", "build_tag.txt")) else None,
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sw_u16_kernel -c 17 " + command}
    json.dump(out, open(dst, "w"), indent=1)
    print("traffic: %d launches, %.1f MB read + %.1f MB written per launch" % (n, rd / 1e6, wr / 1e6))


if __name__ == "__main__":
    main(*sys.argv[1:5])
