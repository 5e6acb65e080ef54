"""GPU bring-up probes (developer tool, run under gpurun).  Each probe runs in its own subprocess
under `timeout` so a trapped kernel cannot take the rest of the session down; output goes to
gpurun_out/check_*.log.

    python tools/gpu_check.py [gemm attn dit pytest]
"""
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def probe_gemm():
    import torch
    import b200dit
    torch.manual_seed(0)
    for (M, N, K, bn) in [(128, 128, 64, 128), (128, 256, 64, 256), (128, 128, 256, 128), (256, 512, 512, 0),
                          (1560, 1536, 1536, 0), (1560, 4608, 1536, 256), (1560, 8960, 1536, 0),
                          (1560, 1536, 8960, 128), (300, 384, 200, 128), (77, 96, 72, 128), (6240, 8960, 1536, 256)]:
        a = torch.randn(M, K).half().cuda()
        w = (torch.randn(N, K) / math.sqrt(K)).half().cuda()
        bias = torch.randn(N).cuda()
        out = b200dit.linear(a, w, bias, "f32", bn)
        torch.cuda.synchronize()
        ref = a.float() @ w.float().t() + bias
        e = rel(out, ref)
        print(f"gemm M={M} N={N} K={K} bn={bn}: rel-L2 {e:.3e}", flush=True)
        if e > 1e-3 and M * N <= 256 * 512:
            d = (out - ref).abs()
            print("  err by 32-row x 32-col block:\n", (d.view(M // 32, 32, N // 32, 32).amax(dim=(1, 3))).cpu())
    # timing of the big shapes
    for (M, N, K) in [(6240, 8960, 1536), (6240, 1536, 8960), (6240, 4608, 1536), (6240, 1536, 1536), (1560, 8960, 1536),
                      (3120, 8960, 1536), (32760, 8960, 1536)]:
        a = torch.randn(M, K).half().cuda()
        w = (torch.randn(N, K) / math.sqrt(K)).half().cuda()
        for epi in ("f16",):
            for _ in range(3):
                b200dit.linear(a, w, None, epi)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for _ in range(10):
                b200dit.linear(a, w, None, epi)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / 10
            t0 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            t0[0].record()
            for _ in range(10):
                torch.matmul(a, w.t())
            t0[1].record()
            torch.cuda.synchronize()
            ms_ref = t0[0].elapsed_time(t0[1]) / 10
            print(f"gemm time M={M} N={N} K={K} {epi}: {ms*1e3:.1f} us = {2*M*N*K/ms/1e9:.0f} TFLOP/s "
                  f"(cuBLAS {ms_ref*1e3:.1f} us = {2*M*N*K/ms_ref/1e9:.0f})", flush=True)


def probe_attn():
    import torch
    import b200dit
    from oracle import dit_oracle as O
    torch.manual_seed(1)
    for (B, Lq, Lk, H, kl) in [(1, 128, 128, 1, None), (1, 128, 256, 1, None), (1, 256, 384, 2, None),
                               (1, 1560, 1560, 12, None), (2, 300, 512, 3, [77, 512]), (1, 130, 1000, 1, [999])]:
        q, k, v = (torch.randn(B, L, H, 128).half().cuda() for L in (Lq, Lk, Lk))
        out = b200dit.flash_attention(q, k, v, k_lens=torch.tensor(kl) if kl else None)
        torch.cuda.synchronize()
        for b in range(B):
            ref = O.softmax_attention(q[b].cpu().float(), k[b].cpu().float(), v[b].cpu().float(), kl[b] if kl else None)
            print(f"attn B={B} Lq={Lq} Lk={Lk} H={H} item {b}: rel-L2 {rel(out[b].cpu().float(), ref):.3e}", flush=True)
    for (B, L, H) in [(1, 1560, 12), (4, 1560, 12), (1, 6240, 12), (1, 32760, 12)]:
        q, k, v = (torch.randn(B, L, H, 128).half().cuda() for _ in range(3))
        for _ in range(2):
            b200dit.flash_attention(q, k, v)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(5):
            b200dit.flash_attention(q, k, v)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 5
        fl = 4.0 * B * H * L * L * 128
        line = f"attn time B={B} L={L}: {ms*1e3:.1f} us = {fl/ms/1e9:.0f} TFLOP/s (incl. V transpose)"
        try:
            from flash_attn import flash_attn_func
            for _ in range(2):
                flash_attn_func(q, k, v)
            ev[0].record()
            for _ in range(5):
                flash_attn_func(q, k, v)
            ev[1].record()
            torch.cuda.synchronize()
            ms2 = ev[0].elapsed_time(ev[1]) / 5
            line += f"; flash-attn2 {ms2*1e3:.1f} us = {fl/ms2/1e9:.0f}"
        except Exception as ex:  # noqa
            line += f"; flash-attn2 unavailable ({type(ex).__name__})"
        print(line, flush=True)


def probe_dit():
    import time
    import torch
    import b200dit
    from oracle import dit_oracle as O
    import __graft_entry__ as g
    g.smoke()
    for layers in (1, 4):
        sd = O.make_synthetic_weights(1536, 8960, 12, layers, seed=21)
        eng = b200dit.DitEngine.from_state_dict(sd, num_heads=12)
        gen = torch.Generator().manual_seed(5)
        x = [torch.randn(16, 1, 60, 104, generator=gen)]
        ctx = [torch.randn(77, 4096, generator=gen)]
        t = torch.tensor([999.0])
        tap = eng.set_tap(0, 1560)
        out = eng.forward(x, t, ctx, 1560)[0].cpu()
        taps = {0: None}
        t0 = time.time()
        ref = O.dit_forward(sd, x, t, ctx, 1560, taps=taps)[0]
        print(f"dit {layers} layers: out rel-L2 {rel(out, ref):.3e}; block-0 stream rel-L2 "
              f"{rel(tap.cpu(), taps[0][0]):.3e} (oracle {time.time()-t0:.1f}s)", flush=True)
        eng.close()


def main():
    os.makedirs(OUT, exist_ok=True)
    stages = sys.argv[1:] or ["gemm", "attn", "dit", "pytest"]
    if len(stages) == 2 and stages[0] == "--probe":
        {"gemm": probe_gemm, "attn": probe_attn, "dit": probe_dit}[stages[1]]()
        return
    for st in stages:
        log = os.path.join(OUT, f"check_{st}.log")
        if st == "pytest":
            cmd = ["timeout", "900", sys.executable, "-m", "pytest", "tests", "-m", "gpu", "-x", "-q", "--timeout", "600"]
        else:
            cmd = ["timeout", "600", sys.executable, os.path.abspath(__file__), "--probe", st]
        with open(log, "w") as f:
            rc = subprocess.run(cmd, stdout=f, stderr=subprocess.STDOUT, cwd=ROOT).returncode
        tail = open(log).read()[-3000:]
        print(f"===== {st}: exit {rc}\n{tail}", flush=True)


if __name__ == "__main__":
    main()
