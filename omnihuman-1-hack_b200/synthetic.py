"""Random-initialised weights of the reference architectures, for benchmarks and tools (no checkpoint is
reachable offline).  Key names and shapes are the reference state_dict's (model.py:463-498, vae.py:388-421);
the DiT init follows WanModel.init_weights (model.py:590-612: xavier-uniform Linears, N(0, .02) embeddings,
modulation ~ N(0,1)/sqrt(dim)) with the zero-initialised head re-randomised so outputs are not constant."""
import math

import torch


def dit_weights(cfg, seed, device, i2v=False):
    """Reference init (model.py:590-612; head re-initialised as SURVEY 8c), generated on the GPU in fp16."""
    g = torch.Generator(device=device).manual_seed(seed)
    d, f = cfg["dim"], cfg["ffn_dim"]

    def xavier(o, i):
        a = math.sqrt(6.0 / (i + o))
        return ((torch.rand(o, i, generator=g, device=device) * 2 - 1) * a).half()

    def normal(*s, std=1.0):
        return (torch.randn(*s, generator=g, device=device) * std).half()

    sd = {"patch_embedding.weight": xavier(d, cfg["in_dim"] * 4).view(d, cfg["in_dim"], 1, 2, 2),
          "patch_embedding.bias": normal(d, std=0.02),
          "text_embedding.0.weight": normal(d, cfg["text_dim"], std=0.02), "text_embedding.0.bias": normal(d, std=0.02),
          "text_embedding.2.weight": normal(d, d, std=0.02), "text_embedding.2.bias": normal(d, std=0.02),
          "time_embedding.0.weight": normal(d, cfg["freq_dim"], std=0.02), "time_embedding.0.bias": normal(d, std=0.02),
          "time_embedding.2.weight": normal(d, d, std=0.02), "time_embedding.2.bias": normal(d, std=0.02),
          "time_projection.1.weight": xavier(6 * d, d), "time_projection.1.bias": normal(6 * d, std=0.02),
          "head.modulation": normal(1, 2, d) / math.sqrt(d), "head.head.weight": normal(64, d, std=0.02),
          "head.head.bias": normal(64, std=0.02)}
    for i in range(cfg["num_layers"]):
        p = f"blocks.{i}."
        sd[p + "modulation"] = normal(1, 6, d) / math.sqrt(d)
        sd[p + "norm3.weight"] = 1.0 + normal(d, std=0.05)
        sd[p + "norm3.bias"] = normal(d, std=0.02)
        for att in ("self_attn", "cross_attn"):
            for n in "qkvo":
                sd[p + f"{att}.{n}.weight"] = xavier(d, d)
                sd[p + f"{att}.{n}.bias"] = normal(d, std=0.02)
            sd[p + f"{att}.norm_q.weight"] = 1.0 + normal(d, std=0.05)
            sd[p + f"{att}.norm_k.weight"] = 1.0 + normal(d, std=0.05)
        sd[p + "ffn.0.weight"] = xavier(f, d); sd[p + "ffn.0.bias"] = normal(f, std=0.02)
        sd[p + "ffn.2.weight"] = xavier(d, f); sd[p + "ffn.2.bias"] = normal(d, std=0.02)
        if i2v:                                            # second K/V stream (model.py:189-230)
            for n in ("k_img", "v_img"):
                sd[p + f"cross_attn.{n}.weight"] = xavier(d, d)
                sd[p + f"cross_attn.{n}.bias"] = normal(d, std=0.02)
            sd[p + "cross_attn.norm_k_img.weight"] = 1.0 + normal(d, std=0.05)
    if i2v:                                                # MLPProj (model.py:362-374)
        sd["img_emb.proj.0.weight"] = 1.0 + normal(1280, std=0.05); sd["img_emb.proj.0.bias"] = normal(1280, std=0.02)
        sd["img_emb.proj.1.weight"] = xavier(1280, 1280); sd["img_emb.proj.1.bias"] = normal(1280, std=0.02)
        sd["img_emb.proj.3.weight"] = xavier(d, 1280); sd["img_emb.proj.3.bias"] = normal(d, std=0.02)
        sd["img_emb.proj.4.weight"] = 1.0 + normal(d, std=0.05); sd["img_emb.proj.4.bias"] = normal(d, std=0.02)
    return sd


def disc_head_weights(dim, seed, device="cpu", names=(16, 26, 36)):
    """Random weights with the reference discriminator's key names (seaweed_apt/model.py:97-121 + :19-42):
    three single-query heads `cross_attn_<n>.*` and `final_proj.{0,1}.*`.  fp32, CPU generator by default so the
    same seed gives the same values on every machine (the golden fixture stores only the seed)."""
    g = torch.Generator(device=device).manual_seed(seed)
    rn = lambda *s, std=1.0: torch.randn(*s, generator=g, device=device) * std
    sd = {}
    for n in names:
        p = f"cross_attn_{n}."
        sd[p + "query_token"] = rn(1, 1, dim) / math.sqrt(dim)
        for nm in ("norm", "q_norm", "k_norm"):
            sd[p + nm + ".weight"] = 1.0 + rn(dim, std=0.1)
            sd[p + nm + ".bias"] = rn(dim, std=0.05)
        for nm in ("q_proj", "k_proj", "v_proj", "o_proj"):
            sd[p + nm + ".weight"] = rn(dim, dim) / math.sqrt(dim)
            sd[p + nm + ".bias"] = rn(dim, std=0.05)
    sd["final_proj.0.weight"] = 1.0 + rn(3 * dim, std=0.1)
    sd["final_proj.0.bias"] = rn(3 * dim, std=0.05)
    sd["final_proj.1.weight"] = rn(1, 3 * dim) / math.sqrt(3 * dim)
    sd["final_proj.1.bias"] = rn(1, std=0.05)
    return sd


def vae_decoder_weights(dim=96, z_dim=16, seed=0, device="cpu", encoder=False):
    """Decoder + conv2 of WanVAE_ (vae.py:369-421, 505-507) with dim_mult [1,2,4,4], 2 res blocks per stage,
    temporal upsampling in the first two stages (vae.py:597-605); encoder=True adds Encoder3d + conv1
    (vae.py:265-314, 504)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, cout, cin, *k, gain=1.0):
        sd[name + ".weight"] = torch.randn(cout, cin, *k, generator=g) * (gain / math.sqrt(cin * math.prod(k)))
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

    dims = [dim * u for u in (4, 4, 4, 2, 1)]
    conv("conv2", z_dim, z_dim, 1, 1, 1)
    conv("decoder.conv1", dims[0], z_dim, 3, 3, 3)
    res("decoder.middle.0.", dims[0], dims[0])
    gamma("decoder.middle.1.norm.gamma", dims[0], 2)
    conv("decoder.middle.1.to_qkv", 3 * dims[0], dims[0], 1, 1)
    conv("decoder.middle.1.proj", dims[0], dims[0], 1, 1, gain=0.5)
    res("decoder.middle.2.", dims[0], dims[0])
    idx = 0
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if i in (1, 2, 3):
            cin = cin // 2
        for _ in range(3):
            res(f"decoder.upsamples.{idx}.", cin, cout)
            cin = cout
            idx += 1
        if i != 3:
            conv(f"decoder.upsamples.{idx}.resample.1", cout // 2, cout, 3, 3)
            if i < 2:                                              # upsample3d stages carry a time_conv (vae.py:84-85)
                conv(f"decoder.upsamples.{idx}.time_conv", 2 * cout, cout, 3, 1, 1)
            idx += 1
    gamma("decoder.head.0.gamma", dims[-1], 3)
    conv("decoder.head.2", 3, dims[-1], 3, 3, 3)
    if encoder:
        edims = [dim * u for u in (1, 1, 2, 4, 4)]
        conv("conv1", 2 * z_dim, 2 * z_dim, 1, 1, 1)
        conv("encoder.conv1", edims[0], 3, 3, 3, 3)
        idx = 0
        for i, (cin, cout) in enumerate(zip(edims[:-1], edims[1:])):
            for _ in range(2):
                res(f"encoder.downsamples.{idx}.", cin, cout)
                cin = cout
                idx += 1
            if i != 3:
                conv(f"encoder.downsamples.{idx}.resample.1", cout, cout, 3, 3)
                if i > 0:                                          # downsample3d stages (vae.py:94-99)
                    conv(f"encoder.downsamples.{idx}.time_conv", cout, cout, 3, 1, 1)
                idx += 1
        res("encoder.middle.0.", edims[-1], edims[-1])
        gamma("encoder.middle.1.norm.gamma", edims[-1], 2)
        conv("encoder.middle.1.to_qkv", 3 * edims[-1], edims[-1], 1, 1)
        conv("encoder.middle.1.proj", edims[-1], edims[-1], 1, 1, gain=0.5)
        res("encoder.middle.2.", edims[-1], edims[-1])
        gamma("encoder.head.0.gamma", edims[-1], 3)
        conv("encoder.head.2", 2 * z_dim, edims[-1], 3, 3, 3)
    return {k: v.half().float().to(device) for k, v in sd.items()}
