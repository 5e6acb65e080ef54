"""Developer tool: per-launch table (duration, tensor-pipe duty, L2 / DRAM throughput) from an
`ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active...,lts__t_bytes.sum.per_second,dram__bytes.sum.per_second --csv`
log.  usage: python tools/launch_table.py log.csv [first_id]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
L = collections.OrderedDict()
for r in data:
    if len(r) < len(hdr):
        continue
    d = L.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]})
    d[r[ix["Metric Name"]]] = (float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]])
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot = 0.0
for k, d in L.items():
    v, u = d["gpu__time_duration.sum"]
    us = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    tot += us
    if k < first:
        continue
    tc = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", (0, ""))[0]
    l2 = d.get("lts__t_bytes.sum.per_second", (0, ""))[0] / 1e12
    dr = d.get("dram__bytes.sum.per_second", (0, ""))[0] / 1e12
    name = d["name"].split("(")[0].replace("void ", "").replace("b2::", "").replace("<unnamed>::", "")
    print(f"{k:4d} {name[:44]:44s} grid {d['grid']:>14s} {us:9.1f} us  tensor {tc:5.1f} %  L2 {l2:5.2f} TB/s  DRAM {dr:5.2f} TB/s")
print(f"total {tot / 1e3:.2f} ms over {len(L)} launches")
