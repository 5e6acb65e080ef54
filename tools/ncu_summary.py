"""Summarises `ncu -i X.ncu-rep --page raw --csv` exports: one line per captured launch with the roofline-relevant
metrics (developer tool).  usage: python tools/ncu_summary.py gpurun_out/ncu_r2/*.raw.csv > profiles/rN_ncu_summary.txt"""
import csv
import re
import sys

WANT = [("gpu__time_duration.sum", "dur_us", 1e-3), ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
        ("dram__bytes_write.sum", "dram_wr_MB", 1e-6), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
        ("dram__bytes.sum.per_second", "dram_GBps", 1e-9), ("lts__t_bytes.sum.per_second", "l2_GBps", 1e-9),
        ("lts__t_bytes.sum", "l2_MB", 1e-6),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_tc_pct", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_active_pct", 1),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pct", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct", 1),
        ("launch__grid_size", "grid", 1), ("launch__registers_per_thread", "regs", 1)]
UNIT = {"ns": 1.0, "us": 1e3, "ms": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte/s": 1.0,
        "Kbyte/s": 1e3, "Mbyte/s": 1e6, "Gbyte/s": 1e9, "Tbyte/s": 1e12}


def num(v, unit):
    try:
        x = float(v.replace(",", ""))
    except Exception:
        return None
    return x * UNIT.get(unit, 1.0)


for path in sys.argv[1:]:
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    if len(rows) < 3:
        continue
    head, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(head)}
    print(f"# {path}")
    for r in rows[2:]:
        name = re.sub(r"^void |b2::|\(anonymous namespace\)::|<unnamed>::|\(.*$", "", r[col["Kernel Name"]])[:60]
        out = []
        for key, label, scale in WANT:
            if key in col:
                v = num(r[col[key]], units[col[key]])
                if v is not None:
                    out.append(f"{label}={v * scale:.4g}")
        d = dict(o.split("=") for o in out)
        if "dur_us" in d and "dram_rd_MB" in d and "dram_GBps" not in d:
            out.append(f"dram_GBps={(float(d['dram_rd_MB']) + float(d['dram_wr_MB'])) / float(d['dur_us']) * 1e3:.0f}")
        print(f"{name:60s} " + " ".join(out))
