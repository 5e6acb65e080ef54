"""Algorithmic FLOP counts of the hot path (2 FLOP per MAC, no padding, no recompute).
Used by bench.py and the tools to turn measured time into TFLOP/s; the test suite keeps independent copies
and checks the two against each other."""


def dit_forward_flops(L, Lc=512, dim=1536, ffn=8960, layers=30, text_dim=4096, patch_k=64, freq_dim=256,
                      context_cached=False, extra_kv_tokens=0):
    """One WanModel.forward over L tokens (model.py:502-563).  Per block: self q,k,v,o projections 8 L d^2,
    self attention 4 L^2 d, cross q,o 4 L d^2, cross k,v 4 Lc d^2, cross attention 4 L Lc d, FFN 4 L d f.
    context_cached leaves out the step-invariant context work (text embedding, cross k/v projections), which
    is counted once per prompt, not per step.  extra_kv_tokens = 257 adds the i2v second K/V stream."""
    d = dim
    ctx_block = 4 * Lc * d * d + 4 * extra_kv_tokens * d * d
    ctx_other = 2 * Lc * text_dim * d + 2 * Lc * d * d
    per_block = 8 * L * d * d + 4 * L * L * d + 4 * L * d * d + ctx_block + 4 * L * (Lc + extra_kv_tokens) * d + 4 * L * d * ffn
    other = 2 * L * patch_k * d * 2 + ctx_other + 2 * d * (freq_dim + d + 6 * d)
    if context_cached:
        per_block -= ctx_block
        other -= ctx_other
    return layers * per_block + other


def dit_block_flops(L, Lc=512, dim=1536, ffn=8960):
    """One WanAttentionBlock (model.py:279-330) over one item of L tokens with the step-invariant context work
    cached: q,k,v,o 8 L d^2 + self attention 4 L^2 d + cross q,o 4 L d^2 + cross attention 4 L Lc d + FFN 4 L d f
    (SURVEY 8d's fused-block micro-benchmark counts exactly this)."""
    d = dim
    return 8 * L * d * d + 4 * L * L * d + 4 * L * d * d + 4 * L * Lc * d + 4 * L * d * ffn


def vae_decode_flops(T, h=60, w=104, dim=96):
    """WanVAE decode (vae.py:544-568) as the reference executes it: latent frame 0 alone (its upsample3d stages skip
    time_conv, vae.py:106-108), then T - 1 frames through the full decoder; conv MACs include causal zero padding."""
    dims = [dim * u for u in (4, 4, 4, 2, 1)]
    c0 = dims[0]

    def one_pass(t, first):
        fl, hh, ww, tt = 0.0, h, w, t
        fl += 2 * tt * hh * ww * 16 * c0 * 27                                     # conv1
        fl += 2 * (2 * 2 * tt * hh * ww * c0 * c0 * 27)                           # two middle residual blocks
        fl += tt * (2 * hh * ww * c0 * 3 * c0 + 2 * hh * ww * c0 * c0 + 4 * (hh * ww) ** 2 * c0)   # middle attention
        for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
            if i in (1, 2, 3):
                cin = cin // 2
            for _ in range(3):
                fl += 2 * tt * hh * ww * 27 * (cin * cout + cout * cout)
                if cin != cout:
                    fl += 2 * tt * hh * ww * cin * cout
                cin = cout
            if i != 3:
                if i < 2 and not first:
                    fl += 2 * tt * hh * ww * 3 * cout * 2 * cout                  # time_conv
                    tt *= 2
                hh, ww = 2 * hh, 2 * ww
                fl += 2 * tt * hh * ww * 9 * cout * (cout // 2)
        fl += 2 * tt * hh * ww * 27 * dims[-1] * 3                                # head conv
        return fl

    return one_pass(1, True) + (one_pass(T - 1, False) if T > 1 else 0.0)


def vae_encode_flops(T, H=480, W=832, dim=96, z_dim=16):
    """WanVAE encode (vae.py:516-542, Encoder3d :265-366) as the reference executes it: frame 0 alone (its
    downsample3d stages skip the temporal conv), then chunks of 4 frames; conv MACs incl. zero padding."""
    dims = [dim * u for u in (1, 1, 2, 4, 4)]

    def chunk(t, first):
        fl, hh, ww, tt = 0.0, H, W, t
        fl += 2 * tt * hh * ww * 3 * dims[0] * 27                                 # conv1
        for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
            for _ in range(2):
                fl += 2 * tt * hh * ww * 27 * (cin * cout + cout * cout)
                if cin != cout:
                    fl += 2 * tt * hh * ww * cin * cout
                cin = cout
            if i != 3:
                hh, ww = hh // 2, ww // 2
                fl += 2 * tt * hh * ww * 9 * cout * cout                          # stride-2 Conv2d
                if i > 0 and not first:
                    tt //= 2
                    fl += 2 * tt * hh * ww * 3 * cout * cout                      # stride-2 temporal conv
        c = dims[-1]
        fl += 2 * (2 * 2 * tt * hh * ww * c * c * 27)                             # two middle residual blocks
        fl += tt * (2 * hh * ww * c * 3 * c + 2 * hh * ww * c * c + 4 * (hh * ww) ** 2 * c)   # middle attention
        fl += 2 * tt * hh * ww * 27 * c * 2 * z_dim + 2 * tt * hh * ww * (2 * z_dim) ** 2      # head conv + conv1 (1x1x1)
        return fl

    return chunk(1, True) + ((T - 1) // 4) * chunk(4, False)
