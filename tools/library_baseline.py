"""Same-box GPU LIBRARY baseline for the bench workload (SURVEY.md 8d "GPU reference baseline").

The reference source cannot travel to the GPU box, so this is the reference's GPU path re-stated with the SAME library
kernels it dispatches to -- nothing of this repo's engine is on the path:
  * every nn.Linear under torch.autocast(float16) (model.py:540) -> cuBLAS / cuBLASLt fp16 GEMMs,
  * attention -> flash_attn 2.8 `flash_attn_func` on fp16 q/k/v (attention.py:111-127 calls FA2's varlen entry point),
  * LayerNorm / RMSNorm / GELU(tanh) / modulation -> PyTorch eager kernels in the dtypes the reference uses (App. A),
  * 3-D RoPE in complex128 per item (model.py:42-69), sinusoid in float64 (model.py:17-27),
  * the per-step text embedding and cross-attention K/V projections the reference recomputes on every forward
    (model.py:532,176-180).
Two call patterns are timed on the bench configuration (Wan2.1-T2V-1.3B, latent [16,1,60,104], 50-step UniPC not
included -- forwards + CFG combine only, which favours the baseline):
  "as called":  cond and uncond forwards as two separate B = 1 calls per sample (text2video.py:238-241),
  "co-batched": one B = 4 call for two samples (the engine's bench batch).
Output: one JSON line per pattern with denoise-steps/s, TFLOP/s and rel-L2 against the engine on the same inputs.
usage: python tools/library_baseline.py [--steps 10] [--layers 30]
"""
import argparse
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def sinusoid(freq_dim, t):
    half = freq_dim // 2
    w = torch.pow(10000.0, -torch.arange(half, dtype=torch.float64, device=t.device) / half)
    ang = t.to(torch.float64)[:, None] * w[None]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=1)


def rope_freqs(grid, device):
    c = 64
    nf, nh, nw = c - 2 * (c // 3), c // 3, c // 3
    f, h, w = grid

    def axis(n, npos):
        inv = 1.0 / torch.pow(10000.0, torch.arange(0, 2 * n, 2, dtype=torch.float64, device=device) / (2 * n))
        return torch.polar(torch.ones(npos, n, dtype=torch.float64, device=device),
                           torch.arange(npos, dtype=torch.float64, device=device)[:, None] * inv[None])

    af, ah, aw = axis(nf, f), axis(nh, h), axis(nw, w)
    return torch.cat([af[:, None, None].expand(f, h, w, nf), ah[None, :, None].expand(f, h, w, nh),
                      aw[None, None, :].expand(f, h, w, nw)], dim=-1).reshape(f * h * w, 1, c)


def rope_apply(x, freqs):
    """model.py:42-69: per item, complex128 multiply, back to fp32.  x [B, L, n, 128]."""
    out = []
    for i in range(x.shape[0]):
        xi = torch.view_as_complex(x[i].to(torch.float64).reshape(x.shape[1], x.shape[2], -1, 2))
        out.append(torch.view_as_real(xi * freqs).flatten(2))
    return torch.stack(out).float()


def rms(x, g, eps=1e-6):
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)).type_as(x) * g


def ln(x, w=None, b=None, eps=1e-6):
    return F.layer_norm(x.float(), (x.shape[-1],), w, b, eps).type_as(x)


def attention(q, k, v):
    from flash_attn import flash_attn_func
    return flash_attn_func(q.half(), k.half(), v.half())


def forward(sd, xs, t, ctx, heads, layers, grid):
    """WanModel.forward (model.py:502-563) for items of one grid; xs [B,16,F,H,W], ctx [B,512,4096]."""
    lin = lambda k, x: F.linear(x, sd[k + ".weight"], sd[k + ".bias"])
    B = xs.shape[0]
    x = F.conv3d(xs, sd["patch_embedding.weight"], sd["patch_embedding.bias"], stride=(1, 2, 2)).flatten(2).transpose(1, 2)
    L, d = x.shape[1], x.shape[2]
    with torch.autocast("cuda", dtype=torch.float32):
        e = lin("time_embedding.2", F.silu(lin("time_embedding.0", sinusoid(256, t).float())))
        e0 = lin("time_projection.1", F.silu(e)).unflatten(1, (6, d))
    context = lin("text_embedding.2", F.gelu(lin("text_embedding.0", ctx), approximate="tanh"))
    freqs = rope_freqs(grid, x.device)
    with torch.autocast("cuda", dtype=torch.float16):
        for i in range(layers):
            p = f"blocks.{i}."
            with torch.autocast("cuda", dtype=torch.float32):
                m = (sd[p + "modulation"] + e0).chunk(6, dim=1)
            u = ln(x).float() * (1 + m[1]) + m[0]
            q = rms(lin(p + "self_attn.q", u), sd[p + "self_attn.norm_q.weight"]).view(B, L, heads, -1)
            k = rms(lin(p + "self_attn.k", u), sd[p + "self_attn.norm_k.weight"]).view(B, L, heads, -1)
            v = lin(p + "self_attn.v", u).view(B, L, heads, -1)
            a = attention(rope_apply(q, freqs), rope_apply(k, freqs), v).flatten(2)
            with torch.autocast("cuda", dtype=torch.float32):
                x = x + lin(p + "self_attn.o", a) * m[2]
            un = ln(x, sd[p + "norm3.weight"], sd[p + "norm3.bias"])
            qc = rms(lin(p + "cross_attn.q", un), sd[p + "cross_attn.norm_q.weight"]).view(B, L, heads, -1)
            kc = rms(lin(p + "cross_attn.k", context), sd[p + "cross_attn.norm_k.weight"]).view(B, -1, heads, 128)
            vc = lin(p + "cross_attn.v", context).view(B, -1, heads, 128)
            x = x + lin(p + "cross_attn.o", attention(qc, kc, vc).flatten(2))
            u2 = ln(x).float() * (1 + m[4]) + m[3]
            y = lin(p + "ffn.2", F.gelu(lin(p + "ffn.0", u2), approximate="tanh"))
            with torch.autocast("cuda", dtype=torch.float32):
                x = x + y * m[5]
    with torch.autocast("cuda", dtype=torch.float32):
        hm = (sd["head.modulation"] + e.unsqueeze(1)).chunk(2, dim=1)
        y = lin("head.head", ln(x) * (1 + hm[1]) + hm[0])
    f, h, w = grid
    y = y.view(B, f, h, w, 1, 2, 2, 16)
    return torch.einsum("bfhwpqrc->bcfphqwr", y).reshape(B, 16, f, h * 2, w * 2).float()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--layers", type=int, default=30)
    a = ap.parse_args()
    import b200dit
    from b200dit import flops, synthetic
    dev = torch.device("cuda:0")
    cfg = dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=a.layers, in_dim=16, out_dim=16, text_dim=4096,
               freq_dim=256)
    sd = synthetic.dit_weights(cfg, 0, dev)
    eng = b200dit.DitEngine.from_state_dict(sd, num_heads=12, device=dev)
    sdg = {k: v.float() for k, v in sd.items()}           # the reference holds fp32 parameters; autocast casts per call
    g = torch.Generator().manual_seed(1)
    grid = (1, 30, 52)
    x = torch.randn(2, 16, 1, 60, 104, generator=g).to(dev)
    ctx = torch.randn(2, 512, 4096, generator=g).to(dev)
    ctx_n = torch.randn(1, 512, 4096, generator=g).to(dev).expand(2, -1, -1).contiguous()
    t = torch.tensor([999.0, 999.0], device=dev)
    guide = 5.0
    step_flop = 2 * 2 * flops.dit_forward_flops(1560, layers=a.layers)       # two samples x (cond + uncond), all work counted

    def as_called():
        outs = []
        for i in range(2):
            c = forward(sdg, x[i:i + 1], t[i:i + 1], ctx[i:i + 1], 12, a.layers, grid)
            u = forward(sdg, x[i:i + 1], t[i:i + 1], ctx_n[i:i + 1], 12, a.layers, grid)
            outs.append(u + guide * (c - u))
        return torch.cat(outs)

    def cobatched():
        o = forward(sdg, torch.cat([x, x]), torch.cat([t, t]), torch.cat([ctx, ctx_n]), 12, a.layers, grid)
        return o[2:] + guide * (o[:2] - o[2:])

    # the caller's autocast: bf16 in WanT2V.generate (text2video.py:202); the blocks re-enter fp16 (model.py:540)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ref = torch.stack(eng.forward_cfg(list(x), t, list(ctx), list(ctx_n), 1560, guide))
        for name, fn in (("as called: 2 samples x (cond, uncond) as four B=1 forwards (text2video.py:238-241)", as_called),
                         ("co-batched: one B=4 forward for 2 samples", cobatched)):
            for _ in range(a.warmup):
                out = fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                out = fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            rel = float((out - ref).norm() / ref.norm())
            print(json.dumps({"impl": "library (cuBLAS fp16 autocast + flash_attn %s + PyTorch eager elementwise)" % __import__("flash_attn").__version__,
                              "pattern": name, "layers": a.layers, "ms_per_2sample_step": ms,
                              "denoise_steps_per_s": 2000.0 / ms, "tflops": step_flop / ms / 1e9,
                              "rel_l2_vs_engine": rel, "steps": a.steps}), flush=True)
        # the engine on the same inputs, same timing harness (graphs on, context cached after the first call)
        for _ in range(a.warmup):
            eng.forward_cfg(list(x), t, list(ctx), list(ctx_n), 1560, guide)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            eng.forward_cfg(list(x), t, list(ctx), list(ctx_n), 1560, guide)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        print(json.dumps({"impl": "engine (libb200dit.so)", "pattern": "forward_cfg, 2 samples", "layers": a.layers,
                          "ms_per_2sample_step": ms, "denoise_steps_per_s": 2000.0 / ms, "steps": a.steps}), flush=True)


if __name__ == "__main__":
    main()
