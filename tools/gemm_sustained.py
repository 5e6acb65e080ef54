"""Developer tool (gpurun): SUSTAINED throughput of the tcgen05 GEMM against cuBLAS on the DiT shapes -- each arm replays
a CUDA graph of 8 launches (distinct weight buffers) for about 1.5 s while nvidia-smi samples the SM clock, so both
arms run under the same power cap.  One JSON line per shape."""
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import b200dit  # noqa: E402


def clocks_during(fn):
    samples, stop = [], threading.Event()

    def poll():
        while not stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                samples.append((float(out[0]), float(out[1])))
            except Exception:
                pass
            time.sleep(0.1)
    th = threading.Thread(target=poll)
    th.start()
    r = fn()
    stop.set()
    th.join()
    samples = samples[len(samples) // 3:] or [(0.0, 0.0)]
    return r, sorted(s[0] for s in samples)[len(samples) // 2], sorted(s[1] for s in samples)[len(samples) // 2]


def sustained(run, seconds=1.5):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            run()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        reps = max(3, int(seconds * 1e3 / max(e0.elapsed_time(e1), 1e-3)))

        def timed():
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        return clocks_during(timed)


M = int(sys.argv[1]) if len(sys.argv) > 1 else 6240
torch.manual_seed(0)
NW = 8
for (N, K) in [(4608, 1536), (8960, 1536), (1536, 8960), (1536, 1536)]:
    a = torch.randn(M, K, device="cuda").half()
    ws = [(torch.randn(N, K, device="cuda") / math.sqrt(K)).half() for _ in range(NW)]
    bias = torch.randn(N, device="cuda")
    bh = bias.half()
    fl = 2.0 * M * N * K * NW
    rec = {"M": M, "N": N, "K": K}
    for name, run in (("ours", lambda: [b200dit.linear(a, w, bias, "f16", 0) for w in ws]),
                      ("cublas", lambda: [torch.nn.functional.linear(a, w, bh) for w in ws])):
        ms, mhz, watts = sustained(run)
        rec[name] = {"tflops": round(fl / ms / 1e9, 1), "us_per_gemm": round(ms * 1e3 / NW, 1), "sm_mhz": mhz, "watts": watts,
                     "frac_of_clock_peak": round(fl / ms / 1e9 / (148 * 8192 * mhz * 1e6 / 1e12), 3) if mhz else None}
    print(json.dumps(rec), flush=True)
