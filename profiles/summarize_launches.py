"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    order = []
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"^void ", "", name)
        v = float(r[vi].replace(",", ""))
        if r[ui] == "ns":
            v /= 1e3
        elif r[ui] == "ms":
            v *= 1e3
        agg[name][0] += 1
        agg[name][1] += v
        order.append((name, v))
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot:.1f} us over {len(order)} launches")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:10.1f} us  {100 * t / tot:5.1f}%  x{n:<4d} avg {t / n:8.1f} us  {name[:110]}")


if __name__ == "__main__":
    main(sys.argv[1])
