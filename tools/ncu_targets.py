"""Developer tool (run under ncu on the GPU box): launches each hot kernel of the DiT step once on the
benchmark shapes (M = 3120 = cond + uncond of one [16,1,60,104] latent) so that one `ncu --set full`
pass captures them: ffn.0 / ffn.2 / o / qkv GEMMs, self- and cross-attention, LayerNorm, RMSNorm+RoPE."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import b200dit  # noqa: E402

torch.manual_seed(0)
M = int(os.environ.get("TARGET_M", "6240"))           # 6240 = cond + uncond of two [16,1,60,104] samples (bench default)
which = sys.argv[1:] or ["gemm", "attn"]
if "gemm" in which:
    for (N, K, epi) in [(8960, 1536, "gelu"), (1536, 8960, "f32"), (1536, 1536, "f32"), (4608, 1536, "f16")]:
        a = torch.randn(M, K, device="cuda").half()
        w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).half()
        bias = torch.randn(N, device="cuda")
        for _ in range(2):
            b200dit.linear(a, w, bias, epi)
        torch.cuda.synchronize()
        print(f"gemm M={M} N={N} K={K} epi={epi}", flush=True)
if "attn" in which:
    for (B, Lq, Lk) in [(M // 1560, 1560, 1560), (M // 1560, 1560, 512)]:
        q = torch.randn(B, Lq, 12, 128, device="cuda").half()
        k = torch.randn(B, Lk, 12, 128, device="cuda").half()
        v = torch.randn(B, Lk, 12, 128, device="cuda").half()
        for _ in range(2):
            b200dit.flash_attention(q, k, v)
        torch.cuda.synchronize()
print("done")
