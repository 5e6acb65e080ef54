import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, b200dit
def t(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (M, N, K) in [(6240, 8960, 1536), (3120, 1536, 1536), (8192, 8192, 8192)]:
    a = torch.randn(M, K).half().cuda(); w = (torch.randn(N, K) / math.sqrt(K)).half().cuda(); bias = torch.randn(N).cuda()
    line = f"M={M} N={N} K={K}:"
    for epi, b in (("f16", None), ("f16", bias), ("gelu", bias), ("f32", bias)):
        us = t(lambda: b200dit.linear(a, w, b, epi, 256))
        line += f" {epi}{'+b' if b is not None else ''} {us:.0f}us={2*M*N*K/us/1e6:.0f}TF"
    print(line, flush=True)
