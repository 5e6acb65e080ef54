"""Developer tool (gpurun, optionally under ncu): flash_attention on the shapes given as B,Lq,Lk triples
(default: the DiT step's self / cross shapes and a long-key control with the same number of query tiles)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import b200dit  # noqa: E402

shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(4, 1560, 1560), (4, 1560, 512), (4, 1560, 6240)]
torch.manual_seed(0)
for (B, Lq, Lk) in shapes:
    q = torch.randn(B, Lq, 12, 128, device="cuda").half()
    k = torch.randn(B, Lk, 12, 128, device="cuda").half()
    v = torch.randn(B, Lk, 12, 128, device="cuda").half()
    for _ in range(3):
        b200dit.flash_attention(q, k, v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        b200dit.flash_attention(q, k, v)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"B={B} Lq={Lq} Lk={Lk}: {us:.1f} us per call (V transpose + attention + combine) = "
          f"{4.0 * B * 12 * Lq * Lk * 128 / us / 1e6:.0f} TFLOP/s", flush=True)
