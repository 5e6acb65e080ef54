"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference WanVAE decode.

Not part of the product path (see oracle/dit_oracle.py header for who may import oracle/).
Parity is pinned against the reference's own vae.py executed in the authoring container
(tests/test_oracle_vs_reference.py) and the fixtures it produced (tests/golden/vae_*.pt); the
reference ships no golden vectors of its own.

The reference decodes latent frame by latent frame, carrying a two-frame input cache per causal
conv (vae.py:14,202-220,423-472,544-568).  Here that mechanism is written as an explicit stream:
every causal conv owns a `history` of the most recent <= 2 input frames; a chunk of new frames is
convolved against [zeros | history | chunk] and the history rolls forward.  `chunks` selects how the
T latent frames are fed: the reference uses [1, 1, 1, ...]; any split whose first chunk is the single
frame 0 is equivalent -- the engine uses [1, T-1].

State-dict keys are the reference's (`conv2.*`, `decoder.conv1.*`, `decoder.middle.{0,2}.residual.
{0,3}.gamma / {2,6}.{weight,bias}`, `decoder.middle.1.{norm.gamma,to_qkv.*,proj.*}`,
`decoder.upsamples.N.{residual...,shortcut.*,resample.1.*,time_conv.*}`, `decoder.head.{0.gamma,2.*}`).
"""
import math

import torch
import torch.nn.functional as F

VAE_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
            0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]   # vae.py:629-632
VAE_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
           3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]        # vae.py:633-636

# decoder plan for dim=96, dim_mult=[1,2,4,4], num_res_blocks=2, temporal upsample [T,T,F]
# (vae.py:388-416,597-605): (kind, in, out)
def decoder_plan(dim=96, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temporal_upsample=(True, True, False)):
    dims = [dim * u for u in [dim_mult[-1]] + list(dim_mult[::-1])]
    plan = []
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if i in (1, 2, 3):
            cin = cin // 2
        for _ in range(num_res_blocks + 1):
            plan.append(("res", cin, cout))
            cin = cout
        if i != len(dim_mult) - 1:
            plan.append(("up3d" if temporal_upsample[i] else "up2d", cout, cout // 2))
    return dims[0], plan


class _Stream:
    """Two-frame rolling input history of one causal conv (vae.py:14,28-36,207-217)."""

    def __init__(self):
        self.hist = {}

    def causal_conv(self, name, x, w, b):
        kt = w.shape[2]
        if kt == 1:                                     # 1x1x1 shortcut / conv2: no temporal support
            return F.conv3d(x, w, b)
        h = self.hist.get(name)
        real = x if h is None else torch.cat([h, x], dim=2)
        have = 0 if h is None else h.shape[2]
        if have < kt - 1:                                 # left zero padding = 2 - cached frames (:30-34)
            zeros = x.new_zeros(x.shape[0], x.shape[1], (kt - 1) - have, x.shape[3], x.shape[4])
            full = torch.cat([zeros, real], dim=2)
        else:
            full = real
        self.hist[name] = real[:, :, -2:].clone()         # only real frames are remembered, never padding
        ph, pw = w.shape[3] // 2, w.shape[4] // 2
        return F.conv3d(F.pad(full, (pw, pw, ph, ph, 0, 0)), w, b)


def rms_norm_c(x, gamma):
    """vae.py:51-54 -- F.normalize over channels * sqrt(C) * gamma."""
    c = x.shape[1]
    return F.normalize(x, dim=1) * math.sqrt(c) * gamma.view(1, c, *([1] * (x.dim() - 2)))


def _res_block(sd, p, st, x, cin, cout):
    """vae.py:186-220."""
    h = F.conv3d(x, sd[p + "shortcut.weight"], sd[p + "shortcut.bias"]) if cin != cout else x
    y = F.silu(rms_norm_c(x, sd[p + "residual.0.gamma"]))
    y = st.causal_conv(p + "residual.2", y, sd[p + "residual.2.weight"], sd[p + "residual.2.bias"])
    y = F.silu(rms_norm_c(y, sd[p + "residual.3.gamma"]))
    y = st.causal_conv(p + "residual.6", y, sd[p + "residual.6.weight"], sd[p + "residual.6.bias"])
    return y + h


def _attn_block(sd, p, x):
    """vae.py:223-262 -- per-frame single-head attention over h*w positions."""
    b, c, t, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = rms_norm_c(y, sd[p + "norm.gamma"])
    qkv = F.conv2d(y, sd[p + "to_qkv.weight"], sd[p + "to_qkv.bias"]).reshape(b * t, 3, c, h * w)
    q, k, v = (qkv[:, i].transpose(1, 2) for i in range(3))           # [bt, hw, c]
    a = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(c), dim=-1) @ v
    a = a.transpose(1, 2).reshape(b * t, c, h, w)
    a = F.conv2d(a, sd[p + "proj.weight"], sd[p + "proj.bias"])
    return a.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4) + x


def _upsample(sd, p, st, x, kind, first_chunk):
    """vae.py:101-141 (upsample2d / upsample3d)."""
    b, c, t, h, w = x.shape
    if kind == "up3d" and not first_chunk:          # frame 0 skips time_conv ('Rep' sentinel, :106-108)
        y = st.causal_conv(p + "time_conv", x, sd[p + "time_conv.weight"], sd[p + "time_conv.bias"])
        y = y.reshape(b, 2, c, t, h, w)              # channels [0:C] -> frame 2i, [C:2C] -> frame 2i+1
        x = torch.stack((y[:, 0], y[:, 1]), dim=3).reshape(b, c, 2 * t, h, w)
        t = 2 * t
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = F.interpolate(y.float(), scale_factor=(2.0, 2.0), mode="nearest-exact")
    y = F.conv2d(y, sd[p + "resample.1.weight"], sd[p + "resample.1.bias"], padding=1)
    return y.reshape(b, t, c // 2, 2 * h, 2 * w).permute(0, 2, 1, 3, 4)


def decoder_chunk(sd, st, x, first_chunk, plan=None, taps=None):
    """Decoder3d.forward with feat_cache (vae.py:423-472) on one chunk of latent frames."""
    c0, plan = plan or decoder_plan()
    x = st.causal_conv("decoder.conv1", x, sd["decoder.conv1.weight"], sd["decoder.conv1.bias"])
    x = _res_block(sd, "decoder.middle.0.", st, x, c0, c0)
    x = _attn_block(sd, "decoder.middle.1.", x)
    x = _res_block(sd, "decoder.middle.2.", st, x, c0, c0)
    if taps is not None:
        taps.setdefault("middle", []).append(x)
    for i, (kind, cin, cout) in enumerate(plan):
        p = f"decoder.upsamples.{i}."
        x = _res_block(sd, p, st, x, cin, cout) if kind == "res" else _upsample(sd, p, st, x, kind, first_chunk)
        if taps is not None:
            taps.setdefault(f"up{i}", []).append(x)
    x = F.silu(rms_norm_c(x, sd["decoder.head.0.gamma"]))
    return st.causal_conv("decoder.head.2", x, sd["decoder.head.2.weight"], sd["decoder.head.2.bias"])


def vae_decode(sd, z, chunks=None, clamp=True, taps=None):
    """WanVAE.decode for one latent (vae.py:657-663 -> 544-568). z [16,T,h,w] -> [3,1+4(T-1),8h,8w]."""
    zdim, T = z.shape[0], z.shape[1]
    mean = torch.tensor(VAE_MEAN[:zdim]).view(1, zdim, 1, 1, 1)
    std = torch.tensor(VAE_STD[:zdim]).view(1, zdim, 1, 1, 1)
    x = z[None].float() / (1.0 / std) + mean                         # :547-551 with scale=[mean, 1/std]
    x = F.conv3d(x, sd["conv2.weight"], sd["conv2.bias"])            # :553
    chunks = chunks or [1] * T
    assert sum(chunks) == T and chunks[0] == 1
    st, outs, pos = _Stream(), [], 0
    for n in chunks:
        outs.append(decoder_chunk(sd, st, x[:, :, pos:pos + n], first_chunk=(pos == 0), taps=taps))
        pos += n
    out = torch.cat(outs, dim=2)[0].float()
    return out.clamp_(-1, 1) if clamp else out


# ------------------------------------------------------------------------------------------------
# Encoder (vae.py:265-366 Encoder3d, :516-542 WanVAE_.encode): frames are fed as chunks [1, 4, 4, ...]
# with the same per-conv two-frame input history; downsample3d keeps the last frame it saw and, except
# on the first chunk, runs a stride-2 temporal conv over [last frame | chunk] (vae.py:143-160).
def encoder_plan(dim=96, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temporal_downsample=(False, True, True)):
    dims = [dim * u for u in [1] + list(dim_mult)]
    plan = []
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(num_res_blocks):
            plan.append(("res", cin, cout))
            cin = cout
        if i != len(dim_mult) - 1:
            plan.append(("down3d" if temporal_downsample[i] else "down2d", cout, cout))
    return dims[-1], plan


def _downsample(sd, p, st, x, kind):
    """vae.py:93-99,138-160 (downsample2d / downsample3d)."""
    b, c, t, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = F.conv2d(F.pad(y, (0, 1, 0, 1)), sd[p + "resample.1.weight"], sd[p + "resample.1.bias"], stride=2)
    x = y.reshape(b, t, c, y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)
    if kind == "down3d":
        last = st.hist.get(p + "time_conv")
        if last is None:                                   # first chunk: remembered, not convolved (:146-148)
            st.hist[p + "time_conv"] = x.clone()
        else:
            keep = x[:, :, -1:].clone()
            x = F.conv3d(torch.cat([last[:, :, -1:], x], dim=2), sd[p + "time_conv.weight"],
                         sd[p + "time_conv.bias"], stride=(2, 1, 1))
            st.hist[p + "time_conv"] = keep
    return x


def encoder_chunk(sd, st, x, plan=None):
    """Encoder3d.forward with feat_cache (vae.py:316-366) on one chunk of video frames."""
    c_last, plan = plan or encoder_plan()
    x = st.causal_conv("encoder.conv1", x, sd["encoder.conv1.weight"], sd["encoder.conv1.bias"])
    for i, (kind, cin, cout) in enumerate(plan):
        p = f"encoder.downsamples.{i}."
        x = _res_block(sd, p, st, x, cin, cout) if kind == "res" else _downsample(sd, p, st, x, kind)
    x = _res_block(sd, "encoder.middle.0.", st, x, c_last, c_last)
    x = _attn_block(sd, "encoder.middle.1.", x)
    x = _res_block(sd, "encoder.middle.2.", st, x, c_last, c_last)
    x = F.silu(rms_norm_c(x, sd["encoder.head.0.gamma"]))
    return st.causal_conv("encoder.head.2", x, sd["encoder.head.2.weight"], sd["encoder.head.2.bias"])


def vae_encode(sd, video, dim=96):
    """WanVAE.encode for one video (vae.py:641-648 -> 516-542). video [3,T,H,W] in [-1,1], T = 1 + 4k
    -> mu [16, 1 + (T-1)/4, H/8, W/8], normalised with the latent mean / std."""
    T = video.shape[1]
    zdim = sd["conv1.weight"].shape[0] // 2
    plan = encoder_plan(dim)
    st, outs = _Stream(), []
    x = video[None].float()
    for i in range(1 + (T - 1) // 4):
        chunk = x[:, :, :1] if i == 0 else x[:, :, 1 + 4 * (i - 1):1 + 4 * i]
        outs.append(encoder_chunk(sd, st, chunk, plan))
    out = torch.cat(outs, dim=2)
    mu = F.conv3d(out, sd["conv1.weight"], sd["conv1.bias"])[:, :zdim]                 # :533 (mu, log_var chunk)
    mean = torch.tensor(VAE_MEAN[:zdim]).view(1, zdim, 1, 1, 1)
    std = torch.tensor(VAE_STD[:zdim]).view(1, zdim, 1, 1, 1)
    return ((mu - mean) * (1.0 / std))[0]                                              # :534-539, scale = [mean, 1/std]


def make_synthetic_vae_weights(dim=96, z_dim=16, seed=0, round_to=torch.float16, encoder=False):
    """Decoder-only synthetic state dict (kaiming-ish conv init, gamma perturbed away from 1, the
    zero-initialised attention `proj` re-randomised so every path is exercised -- SURVEY 8c)."""
    g = torch.Generator().manual_seed(seed)
    c0, plan = decoder_plan(dim)
    sd = {}

    def conv(name, cout, cin, *k, gain=1.0):
        fan_in = cin * math.prod(k)
        sd[name + ".weight"] = torch.randn(cout, cin, *k, generator=g) * (gain / math.sqrt(fan_in))
        sd[name + ".bias"] = torch.randn(cout, generator=g) * 0.02

    def gamma(name, c, nd):
        sd[name] = (1.0 + 0.1 * torch.randn(c, generator=g)).view(c, *([1] * nd))

    def res(p, cin, cout):
        gamma(p + "residual.0.gamma", cin, 3)
        conv(p + "residual.2", cout, cin, 3, 3, 3)
        gamma(p + "residual.3.gamma", cout, 3)
        conv(p + "residual.6", cout, cout, 3, 3, 3, gain=0.5)
        if cin != cout:
            conv(p + "shortcut", cout, cin, 1, 1, 1)

    conv("conv2", z_dim, z_dim, 1, 1, 1)
    conv("decoder.conv1", c0, z_dim, 3, 3, 3)
    res("decoder.middle.0.", c0, c0)
    gamma("decoder.middle.1.norm.gamma", c0, 2)
    conv("decoder.middle.1.to_qkv", 3 * c0, c0, 1, 1)
    conv("decoder.middle.1.proj", c0, c0, 1, 1, gain=0.5)
    res("decoder.middle.2.", c0, c0)
    for i, (kind, cin, cout) in enumerate(plan):
        p = f"decoder.upsamples.{i}."
        if kind == "res":
            res(p, cin, cout)
        else:
            conv(p + "resample.1", cout, cin, 3, 3)
            if kind == "up3d":
                conv(p + "time_conv", 2 * cin, cin, 3, 1, 1)
    gamma("decoder.head.0.gamma", plan[-1][2], 3)
    conv("decoder.head.2", 3, plan[-1][2], 3, 3, 3)
    if encoder:
        c_last, eplan = encoder_plan(dim)
        conv("conv1", 2 * z_dim, 2 * z_dim, 1, 1, 1)
        conv("encoder.conv1", eplan[0][1], 3, 3, 3, 3)
        for i, (kind, cin, cout) in enumerate(eplan):
            p = f"encoder.downsamples.{i}."
            if kind == "res":
                res(p, cin, cout)
            else:
                conv(p + "resample.1", cout, cin, 3, 3)
                if kind == "down3d":
                    conv(p + "time_conv", cout, cin, 3, 1, 1)
        res("encoder.middle.0.", c_last, c_last)
        gamma("encoder.middle.1.norm.gamma", c_last, 2)
        conv("encoder.middle.1.to_qkv", 3 * c_last, c_last, 1, 1)
        conv("encoder.middle.1.proj", c_last, c_last, 1, 1, gain=0.5)
        res("encoder.middle.2.", c_last, c_last)
        gamma("encoder.head.0.gamma", c_last, 3)
        conv("encoder.head.2", 2 * z_dim, c_last, 3, 3, 3)
    if round_to is not None:
        sd = {k: v.to(round_to).float() for k, v in sd.items()}
    return sd


def vae_decode_flops(T, h=60, w=104, dim=96):
    """Conv MACs*2 as executed by the reference (causal zero-pad included), SURVEY App. C.2."""
    c0, plan = decoder_plan(dim)

    def pass_flops(t, first):
        fl, hh, ww, tt = 0.0, h, w, t
        fl += 2 * tt * hh * ww * 16 * c0 * 27
        fl += 2 * (2 * 2 * tt * hh * ww * c0 * c0 * 27)                        # two middle res blocks
        fl += tt * (2 * hh * ww * c0 * 3 * c0 + 2 * hh * ww * c0 * c0 + 4 * (hh * ww) ** 2 * c0)
        for kind, cin, cout in plan:
            if kind == "res":
                fl += 2 * tt * hh * ww * 27 * (cin * cout + cout * cout)
                if cin != cout:
                    fl += 2 * tt * hh * ww * cin * cout
            else:
                if kind == "up3d" and not first:
                    fl += 2 * tt * hh * ww * 3 * cin * 2 * cin
                    tt *= 2
                hh, ww = 2 * hh, 2 * ww
                fl += 2 * tt * hh * ww * 9 * cin * cout
        fl += 2 * tt * hh * ww * 27 * plan[-1][2] * 3
        return fl

    return pass_flops(1, True) + (pass_flops(T - 1, False) if T > 1 else 0.0)
