"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the APT discriminator's forward.

Not part of the product path (only tests/ and bench.py's CPU legs may import it).  Restates, in fp32:

  * disc_head      seaweed_apt/model.py:19-83   WanCrossAttentionDiscriminatorBlock.forward: one learned query
                   token, LayerNorm of the block output, q/k/v projections, LayerNorm of q and k over all
                   `dim` channels (qk_norm), per-head softmax over the L tokens, o projection -> [B, 1, dim]
  * timestep_shift seaweed_apt/model.py:158-159 t' = s t / (1 + (s - 1) t), s = 1 for single-frame latents,
                   12 for videos
  * disc_forward   seaweed_apt/model.py:123-186 WanAPTDiscriminator.forward: backbone forward (no grad) at the
                   shifted timestep, block outputs 16 / 26 / 36 (1-based: blocks[15], [25], [35]), three heads,
                   concat -> LayerNorm(3 dim) -> Linear(3 dim, 1) (`:117-121`)

The reference hard-codes blocks 16/26/36 and therefore only runs on a backbone of >= 36 blocks
(`blocks[35]` raises IndexError on the 30-block 1.3B); `tap_blocks` makes the choice explicit.

Parity status: pinned against the UNMODIFIED reference classes executed in the authoring container
(oracle/ref_loader.load_reference_apt; tests/test_cpu_oracle.py::test_disc_oracle_vs_live_reference) and
against the logits / feature tokens they produced on a 36-block tiny backbone (tests/golden/disc_tiny.pt, made
by `oracle/make_golden.py disc`).
"""
import math

import torch
import torch.nn.functional as F

from . import dit_oracle

REFERENCE_TAPS = (16, 26, 36)            # 1-based block numbers, model.py:150-155


def timestep_shift(t, frames):
    """model.py:158-159."""
    s = 1.0 if frames == 1 else 12.0
    return s * t / (1.0 + (s - 1.0) * t)


def disc_head(x, w, prefix, num_heads, qk_norm=True, eps=1e-6):
    """model.py:44-83.  x [B, L, dim] fp32 -> [B, 1, dim]."""
    B, L, dim = x.shape
    g = lambda n: w[prefix + n].float()
    xn = F.layer_norm(x, (dim,), g("norm.weight"), g("norm.bias"), eps)
    query = g("query_token").expand(B, -1, -1)
    q = F.linear(query, g("q_proj.weight"), g("q_proj.bias"))
    k = F.linear(xn, g("k_proj.weight"), g("k_proj.bias"))
    v = F.linear(xn, g("v_proj.weight"), g("v_proj.bias"))
    if qk_norm:
        q = F.layer_norm(q, (dim,), g("q_norm.weight"), g("q_norm.bias"), eps)
        k = F.layer_norm(k, (dim,), g("k_norm.weight"), g("k_norm.bias"), eps)
    hd = dim // num_heads
    q = q.view(B, 1, num_heads, hd).transpose(1, 2)
    k = k.view(B, L, num_heads, hd).transpose(1, 2)
    v = v.view(B, L, num_heads, hd).transpose(1, 2)
    p = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(hd), dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, 1, dim)
    return F.linear(o, g("o_proj.weight"), g("o_proj.bias"))


def disc_logit(feats, w):
    """model.py:174-181 + :117-121.  feats: three [B, 1, dim] tokens."""
    cat = torch.cat([f.squeeze(1) for f in feats], dim=-1)
    n = cat.shape[-1]
    y = F.layer_norm(cat, (n,), w["final_proj.0.weight"].float(), w["final_proj.0.bias"].float(), 1e-5)
    return F.linear(y, w["final_proj.1.weight"].float(), w["final_proj.1.bias"].float())


def disc_heads(taps, w, num_heads, qk_norm=True, eps=1e-6, names=REFERENCE_TAPS):
    """The part after the backbone: taps = three [B, L, dim] block outputs."""
    feats = [disc_head(x, w, f"cross_attn_{n}.", num_heads, qk_norm, eps) for x, n in zip(taps, names)]
    return disc_logit(feats, w), feats


def disc_forward(backbone_sd, w, x, t, context, seq_len, num_heads, tap_blocks=REFERENCE_TAPS, qk_norm=True,
                 eps=1e-6):
    """model.py:123-186.  x [B, C, T, H, W]; t [B]; context list of [rows, text_dim].
    tap_blocks are 1-based block numbers whose outputs feed cross_attn_16 / _26 / _36 in that order."""
    ts = timestep_shift(t, x.shape[2])
    taps = {b - 1: None for b in tap_blocks}
    dit_oracle.dit_forward(backbone_sd, [u for u in x], ts, context, seq_len, num_heads=num_heads, taps=taps,
                           pad_rows=True)
    # the reference's block outputs are [B, seq_len, dim]: when seq_len exceeds the token count the heads also
    # attend over the padded rows (model.py:162-171), which the blocks have turned into non-zero rows
    stacked = [torch.stack(taps[b - 1]) for b in tap_blocks]
    assert all(s.shape[1] == seq_len for s in stacked)
    return disc_heads(stacked, w, num_heads, qk_norm, eps)
