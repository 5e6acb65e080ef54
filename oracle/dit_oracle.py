"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference DiT forward.

Not part of the product path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this file. The shipped engine never calls it.

Parity status: the reference holds NO golden vectors or tests for this path,
so parity is pinned the only way available: this restatement is checked against the reference's own
modules executed in the authoring container (oracle/ref_loader.py, tests/test_oracle_vs_reference.py)
and against the fixtures those modules produced (tests/golden/*.pt, made by oracle/make_golden.py).

Everything is a pure function of a flat `sd` dict that uses the reference state_dict key names
(model.py:463-498):  patch_embedding.{weight,bias}, text_embedding.{0,2}.*, time_embedding.{0,2}.*,
time_projection.1.*, blocks.N.{modulation, norm3.*, self_attn.{q,k,v,o,norm_q,norm_k}.*,
cross_attn.{q,k,v,o,norm_q,norm_k[,k_img,v_img,norm_k_img]}.*, ffn.{0,2}.*}, head.{modulation,head.*},
optional img_emb.proj.{0,1,3,4}.*.

Math follows the reference line by line (cited per function); evaluation is fp32 with the float64
pieces the reference itself uses (sinusoid model.py:17-27, RoPE model.py:42-69).
"""
import math

import torch
import torch.nn.functional as F

EPS = 1e-6

# Yardstick mode (SURVEY section 7 hard part 1): what the reference's OWN GPU path rounds.  Its blocks run under
# autocast(float16) (model.py:540): every nn.Linear takes fp16 operands and returns fp16, WanRMSNorm casts the
# normalised value back to fp16 before the weight (model.py:85-88), flash_attention runs q/k/v (and FA2's
# probabilities) in fp16 and returns fp16 (attention.py:60-76,129).  With the flag on, block_forward applies those
# roundings to the fp32 restatement, so `rel_l2(autocast-emulated, fp32)` is the reference GPU path's own
# distance from exact arithmetic -- the scale the engine's own deviation is judged against.
_AUTOCAST = {"on": False}


class emulate_autocast:
    def __enter__(self):
        self._old, _AUTOCAST["on"] = _AUTOCAST["on"], True

    def __exit__(self, *a):
        _AUTOCAST["on"] = self._old


def _r16(x):
    return x.half().float() if _AUTOCAST["on"] else x


def sinusoid_256(freq_dim, t):
    """model.py:17-27 -- [cos(t*w_i) | sin(t*w_i)], w_i = 10000^(-i/half), float64."""
    half = freq_dim // 2
    pos = t.to(torch.float64)
    w = torch.pow(torch.tensor(10000.0, dtype=torch.float64),
                  -torch.arange(half, dtype=torch.float64) / half)
    ang = pos[:, None] * w[None, :]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=1)


def rope_angles(head_dim, grid):
    """model.py:31-38,46-61,487-492 -- per-token angle table [F*H*W, head_dim/2] in float64.

    64 complex pairs per head split (22, 21, 21) over (frame, row, col); pair j of an axis whose
    sub-dimension is D rotates by pos * 10000^(-2j/D).
    """
    c = head_dim // 2
    n_f, n_h, n_w = c - 2 * (c // 3), c // 3, c // 3
    f, h, w = grid

    def axis(npair, npos):
        dim_axis = 2 * npair
        inv = 1.0 / torch.pow(torch.tensor(10000.0, dtype=torch.float64),
                              torch.arange(0, dim_axis, 2, dtype=torch.float64) / dim_axis)
        return torch.arange(npos, dtype=torch.float64)[:, None] * inv[None, :]

    af, ah, aw = axis(n_f, f), axis(n_h, h), axis(n_w, w)
    ang = torch.cat([
        af[:, None, None, :].expand(f, h, w, n_f),
        ah[None, :, None, :].expand(f, h, w, n_h),
        aw[None, None, :, :].expand(f, h, w, n_w)], dim=-1)
    return ang.reshape(f * h * w, c)


def rope_rotate(x, grid):
    """model.py:42-69 -- x [L, heads, head_dim]; lanes (2j,2j+1) form complex pair j; fp64 -> fp32."""
    L, n, d = x.shape
    ang = rope_angles(d, grid)
    nt = ang.shape[0]
    xr = x[:nt].to(torch.float64).reshape(nt, n, d // 2, 2)
    cos, sin = torch.cos(ang)[:, None, :], torch.sin(ang)[:, None, :]
    a, b = xr[..., 0], xr[..., 1]
    rot = torch.stack([a * cos - b * sin, a * sin + b * cos], dim=-1).reshape(nt, n, d)
    return torch.cat([rot, x[nt:].to(torch.float64)], dim=0).float()


def layer_norm(x, weight=None, bias=None):
    """model.py:91-104 -- fp32 LayerNorm over the last dim, eps 1e-6, optional affine."""
    return F.layer_norm(x.float(), (x.shape[-1],), weight, bias, EPS)


def rms_norm(x, gamma):
    """model.py:80-88 -- over the full channel dim (all heads), fp32, eps 1e-6 (model.py:129-130)."""
    xf = x.float()
    return _r16(xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + EPS)) * gamma


def linear(sd, key, x):
    return F.linear(x, sd[key + ".weight"], sd.get(key + ".bias"))


def block_linear(sd, key, x):
    """nn.Linear inside the autocast(float16) region (model.py:540): exact here, fp16 in / fp16 out in yardstick mode."""
    if not _AUTOCAST["on"]:
        return linear(sd, key, x)
    b = sd.get(key + ".bias")
    return _r16(F.linear(_r16(x), _r16(sd[key + ".weight"]), None if b is None else _r16(b)))


def softmax_attention(q, k, v, k_len):
    """attention.py:24-130 semantics: non-causal, scale 1/sqrt(d), keys j >= k_len masked.
    q [Lq, n, d], k/v [Lk, n, d] -> [Lq, n, d] fp32."""
    d = q.shape[-1]
    if q.shape[0] * k.shape[0] * q.shape[1] > (1 << 28):       # long sequences (T = 21: L = 32 760): query chunks
        step = max(1, (1 << 28) // (k.shape[0] * q.shape[1]))
        return torch.cat([softmax_attention(q[i:i + step], k, v, k_len) for i in range(0, q.shape[0], step)])
    qh, kh, vh = (_r16(t.float()).permute(1, 0, 2) for t in (q, k, v))
    s = torch.matmul(qh, kh.transpose(1, 2)) / math.sqrt(d)
    lk = k.shape[0]
    if k_len is not None and k_len < lk:
        s[:, :, k_len:] = float("-inf")
    if _AUTOCAST["on"]:                     # FA2: un-normalised fp16 probabilities feed P.V, fp32 row sums
        e = torch.exp(s - s.amax(dim=-1, keepdim=True))
        return _r16(torch.matmul(_r16(e), vh) / e.sum(dim=-1, keepdim=True)).permute(1, 0, 2).contiguous()
    return torch.matmul(torch.softmax(s, dim=-1), vh).permute(1, 0, 2).contiguous()


def patch_embed(sd, latent, patch=(1, 2, 2)):
    """model.py:463-464,515-518 -- non-overlapping patch GEMM; token order (f, h, w)."""
    w = sd["patch_embedding.weight"]
    y = F.conv3d(latent[None].float(), w, sd["patch_embedding.bias"], stride=patch)
    grid = tuple(y.shape[2:])
    return y.flatten(2).transpose(1, 2)[0], grid


def time_embed(sd, t, freq_dim=256):
    """model.py:526-528 -- e [B, dim], e0 [B, 6, dim] (fp32)."""
    s = sinusoid_256(freq_dim, t).float()
    e = linear(sd, "time_embedding.2", F.silu(linear(sd, "time_embedding.0", s)))
    e0 = linear(sd, "time_projection.1", F.silu(e))
    return e, e0.unflatten(1, (6, e.shape[1]))


def text_embed(sd, ctx, text_len=512):
    """model.py:531-532 -- zero-pad to text_len rows, Linear, GELU(tanh), Linear."""
    pad = torch.cat([ctx.float(), ctx.new_zeros(text_len - ctx.shape[0], ctx.shape[1]).float()])
    return linear(sd, "text_embedding.2", F.gelu(linear(sd, "text_embedding.0", pad), approximate="tanh"))


def img_embed(sd, clip_fea):
    """model.py:362-374 MLPProj: LN -> Linear -> GELU(erf) -> Linear -> LN (default eps 1e-5)."""
    p = "img_emb.proj."
    x = F.layer_norm(clip_fea.float(), (clip_fea.shape[-1],), sd[p + "0.weight"], sd[p + "0.bias"], 1e-5)
    x = F.gelu(F.linear(x, sd[p + "1.weight"], sd[p + "1.bias"]))
    x = F.linear(x, sd[p + "3.weight"], sd[p + "3.bias"])
    return F.layer_norm(x, (x.shape[-1],), sd[p + "4.weight"], sd[p + "4.bias"], 1e-5)


def block_forward(sd, i, x, e0, grid, ctx, ctx_len, num_heads, n_img=0, ffn_no_grad=False):
    """model.py:279-330 for one item. x [L, dim] fp32, e0 [6, dim], ctx [n_img + text_len, dim].

    n_img > 0 selects the i2v cross-attention (model.py:204-230): the first n_img context rows form
    an un-masked second K/V stream whose attention output is summed before the `o` projection.
    ctx_len is the text key length handed to flash_attention (already incremented by n_img as the
    reference does at model.py:537 -- and clamped by the packed key count).
    ffn_no_grad: the reference evaluates the FFN of blocks with block_idx > 10 under torch.no_grad()
    (model.py:318-325; `y + 0 * ffn_input` keeps the graph connected with a zero gradient): same values,
    the FFN is a constant of the backward (its weights get no gradient, the gate e[5] still does).
    """
    p = f"blocks.{i}."
    L, dim = x.shape
    hd = dim // num_heads
    m = sd[p + "modulation"][0] + e0                       # model.py:289
    sh1, sc1, g1, sh2, sc2, g2 = m.unbind(0)

    u = layer_norm(x) * (1 + sc1) + sh1                      # :292-293
    q = rms_norm(block_linear(sd, p + "self_attn.q", u), sd[p + "self_attn.norm_q.weight"]).view(L, num_heads, hd)
    k = rms_norm(block_linear(sd, p + "self_attn.k", u), sd[p + "self_attn.norm_k.weight"]).view(L, num_heads, hd)
    v = block_linear(sd, p + "self_attn.v", u).view(L, num_heads, hd)
    ntok = grid[0] * grid[1] * grid[2]
    a = softmax_attention(rope_rotate(q, grid), rope_rotate(k, grid), v, ntok)  # :151-156 (k_lens = seq_lens)
    x = x + block_linear(sd, p + "self_attn.o", a.reshape(L, dim)) * g1              # :159-160,296

    un = layer_norm(x, sd[p + "norm3.weight"], sd[p + "norm3.bias"])           # :313 (affine)
    qc = rms_norm(block_linear(sd, p + "cross_attn.q", un), sd[p + "cross_attn.norm_q.weight"]).view(L, num_heads, hd)
    ctx_txt = ctx[n_img:]
    kc = rms_norm(block_linear(sd, p + "cross_attn.k", ctx_txt), sd[p + "cross_attn.norm_k.weight"]).view(-1, num_heads, hd)
    vc = block_linear(sd, p + "cross_attn.v", ctx_txt).view(-1, num_heads, hd)
    ca = softmax_attention(qc, kc, vc, min(ctx_len, ctx_txt.shape[0]))
    if n_img:
        ctx_img = ctx[:n_img]
        ki = rms_norm(block_linear(sd, p + "cross_attn.k_img", ctx_img),
                      sd[p + "cross_attn.norm_k_img.weight"]).view(-1, num_heads, hd)
        vi = block_linear(sd, p + "cross_attn.v_img", ctx_img).view(-1, num_heads, hd)
        ca = ca + softmax_attention(qc, ki, vi, None)                            # :221,228
    x = x + block_linear(sd, p + "cross_attn.o", ca.reshape(L, dim))                 # no gate

    u2 = layer_norm(x) * (1 + sc2) + sh2                                         # :314-315
    ffn = lambda v: block_linear(sd, p + "ffn.2", _r16(F.gelu(block_linear(sd, p + "ffn.0", v), approximate="tanh")))
    if ffn_no_grad:
        with torch.no_grad():
            y = ffn(u2)
        y = y + 0 * u2                                                           # :325
    else:
        y = ffn(u2)
    return x + y * g2                                                            # :328


def head_unpatchify(sd, x, e, grid, out_dim=16, patch=(1, 2, 2)):
    """model.py:349-359 + 565-588. x [L, dim], e [dim] (note: e, not e0)."""
    m = sd["head.modulation"][0] + e[None, :]
    y = linear(sd, "head.head", layer_norm(x) * (1 + m[1]) + m[0])
    f, h, w = grid
    ntok = f * h * w
    y = y[:ntok].view(f, h, w, *patch, out_dim)
    y = torch.einsum("fhwpqrc->cfphqwr", y)
    return y.reshape(out_dim, f * patch[0], h * patch[1], w * patch[2]).float()


def dit_forward(sd, x, t, context, seq_len=None, clip_fea=None, y=None, num_heads=12,
                num_layers=None, text_len=512, freq_dim=256, out_dim=16, taps=None, pad_rows=False,
                ffn_no_grad_from=None):
    """WanModel.forward (model.py:502-563): x list of [C,F,H,W]; t [B]; context list of [Lc,text_dim].

    Returns list of fp32 [out_dim, F, H, W]. `taps` (optional dict) receives the fp32 residual
    stream after selected blocks: taps = {k: None} -> filled with a list over items.
    pad_rows: carry the reference's seq_len - L zero rows per item through the blocks (model.py:522: the padded rows
    are queries like any other -- un-rotated, model.py:66 -- but never keys, model.py:155); the outputs do not
    depend on them, the taps (block outputs [seq_len, dim], read by the APT discriminator) do.
    ffn_no_grad_from: first block index whose FFN is evaluated under no_grad (the reference: 11, model.py:318;
    only matters to autograd).
    """
    if num_layers is None:
        num_layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    t = torch.as_tensor(t).reshape(-1)
    B = len(x)
    if y is not None:                                            # :511-512
        x = [torch.cat([u, v], dim=0) for u, v in zip(x, y)]
    e_all, e0_all = time_embed(sd, t, freq_dim)
    outs = []
    for b in range(B):
        xb, grid = patch_embed(sd, x[b])
        L = xb.shape[0]
        if seq_len is not None:
            assert L <= seq_len, f"Max seq len {L} exceeds limit {seq_len}"   # :521
            if pad_rows and seq_len > L:
                xb = torch.cat([xb, xb.new_zeros(seq_len - L, xb.shape[1])])        # :522
        ctx = text_embed(sd, context[b], text_len)
        ctx_len = context[b].shape[0]
        n_img = 0
        if clip_fea is not None:                                 # :534-537
            img = img_embed(sd, clip_fea[b])
            ctx = torch.cat([img, ctx], dim=0)
            n_img = img.shape[0]
            ctx_len = ctx_len + n_img
        for i in range(num_layers):
            xb = block_forward(sd, i, xb, e0_all[b], grid, ctx, ctx_len, num_heads, n_img,
                               ffn_no_grad=ffn_no_grad_from is not None and i >= ffn_no_grad_from)
            if taps is not None and i in taps:
                taps[i] = (taps[i] or []) + [xb.clone()]
        outs.append(head_unpatchify(sd, xb, e_all[b], grid, out_dim))
    return outs


def cfg_combine(cond, uncond, scale):
    """text2video.py:243-244 / generate.py:229."""
    return uncond + scale * (cond - uncond)


# ---------------------------------------------------------------------------
# Synthetic weights: the reference's own init (model.py:590-612)
# with the zeroed head re-initialised, every tensor rounded to fp16-representable values so that
# the engine's fp16 weight packing is lossless and parity measures arithmetic, not quantisation.
# ---------------------------------------------------------------------------
def make_synthetic_weights(dim=1536, ffn_dim=8960, num_heads=12, num_layers=30, in_dim=16, out_dim=16,
                           text_dim=4096, freq_dim=256, seed=0, i2v=False, patch=(1, 2, 2),
                           round_to=torch.float16):
    g = torch.Generator().manual_seed(seed)

    def xavier(o, i):
        a = math.sqrt(6.0 / (i + o))
        return (torch.rand(o, i, generator=g) * 2 - 1) * a

    def normal(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    sd = {}
    pk = patch[0] * patch[1] * patch[2]
    sd["patch_embedding.weight"] = xavier(dim, in_dim * pk).view(dim, in_dim, *patch)
    sd["patch_embedding.bias"] = normal(dim, std=0.02)
    sd["text_embedding.0.weight"] = normal(dim, text_dim, std=0.02)
    sd["text_embedding.0.bias"] = normal(dim, std=0.02)
    sd["text_embedding.2.weight"] = normal(dim, dim, std=0.02)
    sd["text_embedding.2.bias"] = normal(dim, std=0.02)
    sd["time_embedding.0.weight"] = normal(dim, freq_dim, std=0.02)
    sd["time_embedding.0.bias"] = normal(dim, std=0.02)
    sd["time_embedding.2.weight"] = normal(dim, dim, std=0.02)
    sd["time_embedding.2.bias"] = normal(dim, std=0.02)
    sd["time_projection.1.weight"] = xavier(6 * dim, dim)
    sd["time_projection.1.bias"] = normal(6 * dim, std=0.02)
    for i in range(num_layers):
        p = f"blocks.{i}."
        sd[p + "modulation"] = normal(1, 6, dim) / math.sqrt(dim)
        sd[p + "norm3.weight"] = 1.0 + normal(dim, std=0.05)
        sd[p + "norm3.bias"] = normal(dim, std=0.02)
        for att in ("self_attn", "cross_attn"):
            names = ["q", "k", "v", "o"] + (["k_img", "v_img"] if (i2v and att == "cross_attn") else [])
            for n in names:
                sd[p + f"{att}.{n}.weight"] = xavier(dim, dim)
                sd[p + f"{att}.{n}.bias"] = normal(dim, std=0.02)
            norms = ["norm_q", "norm_k"] + (["norm_k_img"] if (i2v and att == "cross_attn") else [])
            for n in norms:
                sd[p + f"{att}.{n}.weight"] = 1.0 + normal(dim, std=0.05)
        sd[p + "ffn.0.weight"] = xavier(ffn_dim, dim)
        sd[p + "ffn.0.bias"] = normal(ffn_dim, std=0.02)
        sd[p + "ffn.2.weight"] = xavier(dim, ffn_dim)
        sd[p + "ffn.2.bias"] = normal(dim, std=0.02)
    sd["head.modulation"] = normal(1, 2, dim) / math.sqrt(dim)
    sd["head.head.weight"] = normal(out_dim * pk, dim, std=0.02)
    sd["head.head.bias"] = normal(out_dim * pk, std=0.02)
    if i2v:
        sd["img_emb.proj.0.weight"] = 1.0 + normal(1280, std=0.05)
        sd["img_emb.proj.0.bias"] = normal(1280, std=0.02)
        sd["img_emb.proj.1.weight"] = xavier(1280, 1280)
        sd["img_emb.proj.1.bias"] = normal(1280, std=0.02)
        sd["img_emb.proj.3.weight"] = xavier(dim, 1280)
        sd["img_emb.proj.3.bias"] = normal(dim, std=0.02)
        sd["img_emb.proj.4.weight"] = 1.0 + normal(dim, std=0.05)
        sd["img_emb.proj.4.bias"] = normal(dim, std=0.02)
    if round_to is not None:
        sd = {k: v.to(round_to).float() for k, v in sd.items()}
    return sd


def dit_flops(L, Lc=512, dim=1536, ffn=8960, layers=30, text_dim=4096, patch_k=64, freq_dim=256,
              context_cached=False):
    """Algorithmic FLOPs of one forward, 2 FLOP/MAC, no padding.
    context_cached: leave out the step-invariant context work (text embedding, cross-attention k/v
    projections), which SURVEY 8d counts once per prompt rather than once per step."""
    d = dim
    ctx_block = 4 * Lc * d * d
    ctx_other = 2 * Lc * text_dim * d + 2 * Lc * d * d
    per_block = 8 * L * d * d + 4 * L * L * d + 4 * L * d * d + ctx_block + 4 * L * Lc * d + 4 * L * d * ffn
    other = 2 * L * patch_k * d * 2 + ctx_other + 2 * d * (freq_dim + d + 6 * d)
    if context_cached:
        per_block -= ctx_block
        other -= ctx_other
    return layers * per_block + other
