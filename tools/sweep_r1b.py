"""Developer sweep (run under gpurun): tile-width sweep of the tcgen05 GEMM on the DiT shapes with the
weights streaming from HBM (rotating over distinct weight buffers, launches captured in one CUDA
graph so host launch cost does not pace the GPU), and the attention kernel on the DiT shapes.

    python tools/sweep_r1b.py gemm [M ...]
    python tools/sweep_r1b.py attn
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def time_graph(fn, reps=3):
    import torch
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def sweep_gemm(Ms):
    import torch
    import b200dit
    torch.manual_seed(0)
    NW = 8
    for M in Ms:
        for (N, K) in [(1536, 1536), (4608, 1536), (8960, 1536), (1536, 8960)]:
            a = torch.randn(M, K, device="cuda").half()
            ws = [(torch.randn(N, K, device="cuda") / math.sqrt(K)).half() for _ in range(NW)]
            bias = torch.randn(N, device="cuda")
            fl = 2.0 * M * N * K
            line = f"M={M} N={N} K={K}:"
            ref = None
            for bn in (0, 144, 192, 256, 1192, 1224, 1256):
                def run():
                    for w in ws:
                        b200dit.linear(a, w, bias, "f16", bn)
                try:
                    ms = time_graph(run) / NW
                except Exception as ex:  # noqa
                    line += f" bn{bn}=ERR({str(ex)[:60]})"
                    continue
                out = b200dit.linear(a, ws[0], bias, "f16", bn)
                if ref is None:
                    ref = a.float() @ ws[0].float().t() + bias
                err = float((out.float() - ref).norm() / ref.norm())
                line += f" bn{bn}={ms*1e3:.1f}us/{fl/ms/1e9:.0f}T" + ("" if err < 1e-3 else f"(ERR {err:.1e})")

            def run_cublas():
                for w in ws:
                    torch.nn.functional.linear(a, w, bias.half())
            ms = time_graph(run_cublas) / NW
            line += f" | cuBLAS={ms*1e3:.1f}us/{fl/ms/1e9:.0f}T"
            print(line, flush=True)


def sweep_attn():
    import torch
    import b200dit
    from oracle import dit_oracle as O
    torch.manual_seed(1)
    print("attention kernel:", "v1" if os.environ.get("B200_ATTN_V1") else "v2", flush=True)
    for (B, Lq, Lk, H) in [(2, 1560, 1560, 12), (2, 1560, 512, 12), (4, 1560, 1560, 12), (4, 1560, 512, 12), (1, 6240, 6240, 12),
                           (1, 32760, 32760, 12)]:
        q = torch.randn(B, Lq, H, 128, device="cuda").half()
        k = torch.randn(B, Lk, H, 128, device="cuda").half()
        v = torch.randn(B, Lk, H, 128, device="cuda").half()
        for _ in range(2):
            out = b200dit.flash_attention(q, k, v)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        n = 10 if Lq < 10000 else 3
        ev[0].record()
        for _ in range(n):
            b200dit.flash_attention(q, k, v)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / n
        fl = 4.0 * B * H * Lq * Lk * 128
        line = f"attn B={B} Lq={Lq} Lk={Lk}: {ms*1e3:.1f} us = {fl/ms/1e9:.0f} TFLOP/s (incl. V transpose)"
        if Lq <= 1560:
            ref = O.softmax_attention(q[0].cpu().float(), k[0].cpu().float(), v[0].cpu().float(), None)
            line += f" rel-L2 {float((out[0].cpu().float() - ref).norm() / ref.norm()):.2e}"
        try:
            from flash_attn import flash_attn_func
            for _ in range(2):
                flash_attn_func(q, k, v)
            ev[0].record()
            for _ in range(n):
                flash_attn_func(q, k, v)
            ev[1].record()
            torch.cuda.synchronize()
            ms2 = ev[0].elapsed_time(ev[1]) / n
            line += f"; flash-attn2 {ms2*1e3:.1f} us = {fl/ms2/1e9:.0f}"
        except Exception as ex:  # noqa
            line += f"; flash-attn2 unavailable ({type(ex).__name__})"
        print(line, flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "gemm"
    if what == "gemm":
        sweep_gemm([int(x) for x in sys.argv[2:]] or [3120, 1560, 6240])
    else:
        sweep_attn()
