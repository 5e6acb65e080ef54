"""Developer diagnosis (run under gpurun): which resource paces the GEMM main loop?  Times DiT shapes
with B200_GEMM_DBG = 0 (normal), 1 (no epilogue stores), 2 (no operand loads), 4 (no MMAs) and
combinations.  Each setting runs in its own process because the flag is read at library load."""
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    import b200dit
    from sweep_r1b import time_graph
    torch.manual_seed(0)
    NW = 8
    M = int(os.environ.get("DIAG_M", "3120"))
    for (N, K, bn) in [(8960, 1536, 256), (8960, 1536, 1256), (1536, 8960, 256), (1536, 8960, 1256), (1536, 1536, 256),
                       (1536, 1536, 192), (1536, 1536, 144), (4608, 1536, 256)]:
        a = torch.randn(M, K, device="cuda").half()
        ws = [(torch.randn(N, K, device="cuda") / math.sqrt(K)).half() for _ in range(NW)]
        bias = torch.randn(N, device="cuda")

        def run():
            for w in ws:
                b200dit.linear(a, w, bias, "f16", bn)
        ms = time_graph(run) / NW
        print(f"  M={M} N={N} K={K} bn={bn}: {ms*1e3:.1f} us", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        child()
    else:
        for dbg in ("0", "1", "2", "4", "3", "6"):
            print(f"B200_GEMM_DBG={dbg}", flush=True)
            env = dict(os.environ, B200_GEMM_DBG=dbg, B200_GEMM_SK="0")
            subprocess.run(["timeout", "120", sys.executable, os.path.abspath(__file__), "child"], env=env)
