"""Compare two tools/bench_kernels.py outputs: python tools/cmp_bk.py base.txt new.txt [kinds,...]"""
import re
import sys

def load(p):
    d = {}
    for ln in open(p):
        m = re.match(r"(\S+)\s+(\S+)\s+(.*?)\s+([\d.]+) us\s+([\d.]+) GB/s\s+([\d.]+)% of HBM", ln)
        if m:
            d[(m.group(1), m.group(2))] = (m.group(3).strip(), float(m.group(4)), float(m.group(6)))
    return d

a, b = load(sys.argv[1]), load(sys.argv[2])
kinds = set(sys.argv[3].split(",")) if len(sys.argv) > 3 else None
tot = {}
for k in a:
    if k in b and (kinds is None or k[0] in kinds):
        print(f"{k[0]:10s} {k[1]:4s} {a[k][0]:28s} {a[k][1]:8.1f} -> {b[k][1]:8.1f} us  ({b[k][1] / a[k][1] - 1:+6.1%})   {a[k][2]:5.1f}% -> {b[k][2]:5.1f}% of HBM peak")
        t = tot.setdefault(k[0], [0.0, 0.0])
        t[0] += a[k][1]
        t[1] += b[k][1]
for k, (x, y) in tot.items():
    print(f"SUM {k:10s} {x:8.1f} -> {y:8.1f} us ({y / x - 1:+.1%})")
