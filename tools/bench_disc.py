"""APT discriminator forward at the 1.3B width (developer tool, run under gpurun): backbone with three taps +
heads, timed with CUDA events.  Random-init weights (b200dit.synthetic); taps (10, 20, 30) because the
reference's (16, 26, 36) do not exist on the 30-block 1.3B.

    python tools/bench_disc.py [--items 2] [--frames 1] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200dit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=2)
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    cfg = dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=30, in_dim=16, out_dim=16, text_dim=4096,
               text_len=512, freq_dim=256)
    eng = b200dit.DitEngine(**cfg)
    eng.load_state_dict(b200dit.synthetic.dit_weights(cfg, 0, "cuda"))
    disc = b200dit.AptDiscriminator(eng, b200dit.synthetic.disc_head_weights(1536, 1), tap_blocks=(10, 20, 30))
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(a.items, 16, a.frames, 60, 104, generator=g, device="cuda")
    ctx = [torch.randn(512, 4096, generator=g, device="cuda").half() for _ in range(a.items)]
    t = torch.full((a.items,), 0.7)
    L = a.frames * 30 * 52

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters

    full = timed(lambda: disc(x, t, ctx, L))
    taps = [torch.randn(a.items * L, 1536, device="cuda") for _ in range(3)]
    heads = timed(lambda: disc.heads(taps, a.items, L))
    flops_heads = 3 * a.items * L * (2.0 * 1536 * 1536)          # the K projections (the V projection collapses)
    print(json.dumps({"items": a.items, "tokens_per_item": L, "disc_forward_ms": round(full, 3),
                      "heads_ms": round(heads, 3), "heads_share": round(heads / full, 4),
                      "heads_gemm_tflops": round(flops_heads / heads / 1e9, 1),
                      "reference_heads_gemm_flops_ratio": 2.0}))


if __name__ == "__main__":
    main()
