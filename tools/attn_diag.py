"""Developer diagnosis (gpurun): which part of the attention step paces it?  B200_ATTN_DBG bits:
1 no exp2, 2 no TMEM reads of S, 4 no P store, 8 softmax warps idle, 16 no Q.K^T, 32 no P.V, 64 Q.K^T with its A
operand from TMEM (no Q re-read from shared memory).  One process per setting; `attn_diag.py 0 64 72` picks settings."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    import b200dit
    torch.manual_seed(1)
    for (B, Lq, Lk) in [(4, 1560, 1560), (1, 12480, 12480)]:
        q = torch.randn(B, Lq, 12, 128, device="cuda").half()
        k = torch.randn(B, Lk, 12, 128, device="cuda").half()
        v = torch.randn(B, Lk, 12, 128, device="cuda").half()
        for _ in range(2):
            b200dit.flash_attention(q, k, v)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(5):
            b200dit.flash_attention(q, k, v)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 5
        print(f"  B={B} L={Lq}: {ms*1e3:.1f} us = {4.0*B*12*Lq*Lk*128/ms/1e9:.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for dbg in (sys.argv[1:] or ("0", "1", "2", "4", "3", "7", "8")):
            print(f"B200_ATTN_DBG={dbg}", flush=True)
            subprocess.run(["timeout", "120", sys.executable, os.path.abspath(__file__), "child"],
                           env=dict(os.environ, B200_ATTN_DBG=dbg))
