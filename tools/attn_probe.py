"""Developer probe (gpurun): attention parity by query tile for short / ragged key lengths."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import b200dit  # noqa: E402
from oracle import dit_oracle as O  # noqa: E402

torch.manual_seed(3)
for (B, Lq, Lk, H, kl) in [(1, 1560, 512, 12, [77]), (2, 1560, 512, 12, [512, 77]), (1, 1560, 512, 12, [64]),
                           (1, 1560, 512, 12, [13]), (1, 300, 512, 3, [77]), (1, 1560, 512, 3, [77]),
                           (1, 1560, 512, 12, [129]), (1, 1560, 512, 12, [200])]:
    q = torch.randn(B, Lq, H, 128, device="cuda").half()
    k = torch.randn(B, Lk, H, 128, device="cuda").half()
    v = torch.randn(B, Lk, H, 128, device="cuda").half()
    for rep in range(3):
        out = b200dit.flash_attention(q, k, v, k_lens=torch.tensor(kl))
    torch.cuda.synchronize()
    for b in range(B):
        ref = O.softmax_attention(q[b].cpu().float(), k[b].cpu().float(), v[b].cpu().float(), kl[b])
        d = (out[b].cpu().float() - ref)
        rel = float(d.norm() / ref.norm())
        per_tile = [float(d[t * 128:(t + 1) * 128].abs().max()) for t in range((Lq + 127) // 128)]
        per_head = [float(d[:, h].abs().max()) for h in range(H)]
        print(f"B={B} Lq={Lq} H={H} klen={kl[b]} item {b}: rel {rel:.2e} bad tiles "
              f"{[i for i, e in enumerate(per_tile) if e > 0.02]} bad heads {[i for i, e in enumerate(per_head) if e > 0.02]} "
              f"nan {bool(torch.isnan(out[b]).any())}", flush=True)
