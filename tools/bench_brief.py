"""Prints a one-line digest of bench.py's JSON line (developer tool): python bench.py ... | python tools/bench_brief.py"""
import json
import sys

for ln in sys.stdin.read().strip().splitlines():
    if not ln.startswith("{"):
        continue
    d = json.loads(ln)
    r = d.get("roofline") or {}
    c = d.get("clocks") or {}
    print(f"{d['value']:.2f} steps/s {d['ms_per_step']:.2f} ms/step e2e {d['e2e']['value']:.2f} | step TF {d.get('step_tflops', 0):.0f} "
          f"gemm TF {r.get('achieved', 0):.0f} attn TF {r.get('attention_tflops') or 0:.0f} | "
          f"ms/step {({k: round(v, 2) for k, v in (r.get('by_category_ms_per_step') or {}).items()})} | "
          f"clk {c.get('sm_mhz')} {c.get('reasons')} launches {d.get('gpu_launches')}")
