"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list.

    python tools/ncu_summary.py gpurun_out/x.csv [--per-launch] [--match ehgr] > profiles/x_summary.csv

Default: one row per kernel (launches, total us, share of the listed time, DRAM MB per launch, achieved DRAM GB/s).
--per-launch: one row per launch in launch order (short kernel name, grid, us, read MB, write MB, GB/s).
"""
from __future__ import annotations

import argparse
import csv
import re
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    m = re.match(r"([A-Za-z0-9_:]+(?:<[^(]{0,80}>)?)", name)
    s = m.group(1) if m else name[:80]
    return s.replace("ehgr::", "").replace("(anonymous namespace)::", "")


def load(path):
    rows = OrderedDict()
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        k = int(r["ID"])
        e = rows.setdefault(k, {"name": short(r["Kernel Name"]), "grid": r["Grid Size"], "block": r["Block Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            e["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3 if unit == "ms" else v)
        elif m.startswith("dram__bytes_read"):
            e["rd"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        elif m.startswith("dram__bytes_write"):
            e["wr"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return list(rows.values())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--per-launch", action="store_true")
    ap.add_argument("--match", default="")
    a = ap.parse_args()
    rows = [r for r in load(a.csv) if a.match in r["name"]]
    w = csv.writer(sys.stdout)
    if a.per_launch:
        w.writerow(["kernel", "grid", "block", "us", "dram_read_MB", "dram_write_MB", "dram_GBps"])
        for r in rows:
            b = r.get("rd", 0) + r.get("wr", 0)
            w.writerow([r["name"], r["grid"], r["block"], f"{r.get('us', 0):.2f}", f"{r.get('rd', 0) / 1e6:.3f}",
                        f"{r.get('wr', 0) / 1e6:.3f}", f"{b / max(r.get('us', 0), 1e-9) / 1e3:.1f}"])
        return
    agg = OrderedDict()
    for r in rows:
        e = agg.setdefault(r["name"], {"n": 0, "us": 0.0, "b": 0.0})
        e["n"] += 1
        e["us"] += r.get("us", 0)
        e["b"] += r.get("rd", 0) + r.get("wr", 0)
    tot = sum(e["us"] for e in agg.values()) or 1.0
    w.writerow(["kernel", "launches", "total_us", "share", "avg_us", "dram_MB_per_launch", "dram_GBps"])
    for k, e in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        w.writerow([k, e["n"], f"{e['us']:.1f}", f"{e['us'] / tot:.4f}", f"{e['us'] / e['n']:.2f}",
                    f"{e['b'] / e['n'] / 1e6:.3f}", f"{e['b'] / max(e['us'], 1e-9) / 1e3:.1f}"])


if __name__ == "__main__":
    main()
