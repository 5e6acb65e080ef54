"""Developer stress (gpurun): many attention launches on the DiT shapes; a protocol race shows up as an
mbarrier timeout trap (the kernels bound every wait) or as a result that differs between launches."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import b200dit  # noqa: E402

torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
for (B, Lq, Lk, kl) in [(4, 1560, 1560, None), (4, 1560, 512, [512, 77, 300, 1]), (2, 1560, 257, None)]:
    q = torch.randn(B, Lq, 12, 128, device="cuda").half()
    k = torch.randn(B, Lk, 12, 128, device="cuda").half()
    v = torch.randn(B, Lk, 12, 128, device="cuda").half()
    kt = torch.tensor(kl) if kl else None
    ref = b200dit.flash_attention(q, k, v, k_lens=kt)
    bad = 0
    for i in range(n):
        out = b200dit.flash_attention(q, k, v, k_lens=kt)
        if i % 100 == 99:
            bad += int(not torch.equal(out, ref))
    torch.cuda.synchronize()
    print(f"B={B} Lq={Lq} Lk={Lk}: {n} launches, mismatching samples {bad}", flush=True)
