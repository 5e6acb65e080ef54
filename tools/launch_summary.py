"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (developer tool)."""
import collections
import csv
import re
import sys


def main(path, per_launch=False):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    seq = []
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
        name = row["Kernel Name"]
        key = re.sub(r"^void |b2::|\(anonymous namespace\)::|\(.*$", "", name)[:70]
        grid = row.get("Grid Size", "")
        seq.append((key, grid, v))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:4d} avg={v[1] / v[0]:8.1f} us  {k}")
    print(f"total {tot:.1f} us over {len(seq)} launches")
    if per_launch:
        for i, (k, g, v) in enumerate(seq[:per_launch]):
            print(f"{i:4d} {v:8.1f} us grid={g:>14s} {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
