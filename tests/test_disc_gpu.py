"""APT discriminator forward on the B200 through the C ABI (b200disc_*), against the
UNMODIFIED reference's outputs (tests/golden/disc_tiny.pt) and the CPU oracle (oracle/disc_oracle.py)."""
import os

import pytest
import torch

from conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu

TOL_FEAT = 2e-3          # rel-L2 of a head token: fp16 operands in LayerNorm output and K projection, fp32 elsewhere
TOL_LOGIT = 5e-3         # absolute, logits are O(1)


def _golden():
    import b200dit
    g = torch.load(os.path.join(GOLDEN, "disc_tiny.pt"))
    sd = {k: v.float() for k, v in b200dit.synthetic.dit_weights(g["cfg"], g["seed_backbone"], "cpu").items()}
    hw = b200dit.synthetic.disc_head_weights(g["cfg"]["dim"], g["seed_heads"])
    return g, sd, hw


def test_golden_discriminator_forward():
    """WanAPTDiscriminator.forward end to end: backbone at the shifted timestep with the reference's taps
    (blocks 16 / 26 / 36 of a 36-block tiny Wan), heads, final_proj; one- and three-frame latents, and a case whose
    seq_len exceeds the token count (the heads attend over the padded rows of the block outputs too)."""
    import b200dit
    g, sd, hw = _golden()
    eng = b200dit.DitEngine.from_state_dict(sd, num_heads=g["cfg"]["num_heads"])
    disc = b200dit.AptDiscriminator(eng, hw)
    for case in g["cases"]:
        logit, feats = disc(case["x"].cuda(), case["t"], case["context"], case["seq_len"], return_features=True)
        assert logit.shape == case["logit"].shape and feats[0].shape == case["feats"][0].shape
        assert (logit.cpu() - case["logit"]).abs().max() < TOL_LOGIT
        for a, r in zip(feats, case["feats"]):
            assert rel_l2(a.cpu(), r) < TOL_FEAT
        only = disc(case["x"].cuda(), case["t"], case["context"], case["seq_len"])
        assert torch.equal(only, logit)
    assert eng.nonfinite_rows() == 0


def test_reference_error_behaviour():
    """Default taps on a backbone with fewer than 36 blocks: the reference's hook registration raises IndexError
    (seaweed_apt/model.py:154); a seq_len below the token count trips the backbone's assert (wan model.py:521)."""
    import b200dit
    g = torch.load(os.path.join(GOLDEN, "dit_t2v_tiny.pt"))
    eng = b200dit.DitEngine.from_state_dict({k: v.float() for k, v in g["sd"].items()}, num_heads=g["cfg"]["num_heads"])
    hw = b200dit.synthetic.disc_head_weights(g["cfg"]["dim"], 1)
    with pytest.raises(IndexError):
        b200dit.AptDiscriminator(eng, hw)
    disc = b200dit.AptDiscriminator(eng, hw, tap_blocks=(1, 2, 2))
    x = torch.randn(2, 16, 1, 8, 8)
    ctx = [torch.randn(5, g["cfg"]["text_dim"]) for _ in range(2)]
    with pytest.raises(AssertionError):
        disc(x, torch.tensor([0.5, 0.5]), ctx, 8)
    bad = dict(hw); del bad["final_proj.1.bias"]
    with pytest.raises(RuntimeError):
        b200dit.AptDiscriminator(eng, bad, tap_blocks=(1, 2, 2))


def test_small_backbone_explicit_taps_and_odd_token_count():
    """2-block backbone with explicit taps, against the oracle; 15 tokens per item (not a multiple of 8) makes the
    backbone run one item per call, so the heads run per item too."""
    import b200dit
    from oracle import disc_oracle as DO
    g = torch.load(os.path.join(GOLDEN, "dit_t2v_tiny.pt"))
    sd = {k: v.float() for k, v in g["sd"].items()}
    heads = g["cfg"]["num_heads"]
    eng = b200dit.DitEngine.from_state_dict(sd, num_heads=heads)
    hw = b200dit.synthetic.disc_head_weights(g["cfg"]["dim"], 7)
    disc = b200dit.AptDiscriminator(eng, hw, tap_blocks=(1, 2, 1))
    gen = torch.Generator().manual_seed(3)
    for shape in ((16, 1, 6, 10), (16, 2, 8, 8)):
        x = torch.randn(3, *shape, generator=gen)
        L = shape[1] * (shape[2] // 2) * (shape[3] // 2)
        ctx = [torch.randn(9, g["cfg"]["text_dim"], generator=gen) for _ in range(3)]
        t = torch.tensor([0.2, 0.5, 0.8])
        logit, feats = disc(x, t, ctx, L, return_features=True)
        r_logit, r_feats = DO.disc_forward(sd, hw, x, t, ctx, L, heads, tap_blocks=(1, 2, 1))
        assert (logit.cpu() - r_logit).abs().max() < TOL_LOGIT
        for a, r in zip(feats, r_feats):
            assert rel_l2(a.cpu(), r) < TOL_FEAT


@pytest.mark.parametrize("qk_norm", [True, False])
def test_heads_at_1p3b_width_vs_oracle(qk_norm):
    """The head kernels at the 1.3B width (dim 1536, 12 heads) on a full 480p frame's 1560 tokens plus a ragged
    count (not a multiple of the 128-token pooling chunk), given block outputs with per-channel offsets and a few
    dominant tokens so the softmax is far from uniform."""
    import b200dit
    from oracle import disc_oracle as DO
    dim, heads = 1536, 12
    eng = b200dit.DitEngine(dim=dim, ffn_dim=256, num_heads=heads, num_layers=1, text_dim=32)   # weights never used
    hw = b200dit.synthetic.disc_head_weights(dim, 11)
    if not qk_norm:
        hw = {k: v for k, v in hw.items() if "q_norm" not in k and "k_norm" not in k}
    disc = b200dit.AptDiscriminator(eng, hw, tap_blocks=(1, 1, 1), qk_norm=qk_norm)
    gen = torch.Generator().manual_seed(21)
    for B, L in ((2, 1560), (1, 333), (1, 7800)):           # 7800 = a 5-latent-frame video: 61 pooling chunks
        taps = []
        for _ in range(3):
            x = torch.randn(B, L, dim, generator=gen) * (0.5 + torch.rand(dim, generator=gen)) + torch.randn(dim, generator=gen)
            x[:, :: max(L // 7, 1)] *= 4.0
            taps.append(x)
        logit, feats = disc.heads([u.reshape(B * L, dim).cuda() for u in taps], B, L, return_features=True)
        r_logit, r_feats = DO.disc_heads(taps, hw, heads, qk_norm=qk_norm)
        assert (logit.cpu() - r_logit).abs().max() < TOL_LOGIT
        for a, r in zip(feats, r_feats):
            assert rel_l2(a.cpu(), r) < TOL_FEAT
        again = disc.heads([u.reshape(B * L, dim).cuda() for u in taps], B, L)
        assert torch.equal(again, logit)                                     # deterministic reduction order


def test_heads_at_14b_width_vs_oracle():
    """dim 5120 / 40 heads (the reference's 14B widths): three groups of <= 16 heads in the pooling kernel."""
    import b200dit
    from oracle import disc_oracle as DO
    dim, heads = 5120, 40
    eng = b200dit.DitEngine(dim=dim, ffn_dim=256, num_heads=heads, num_layers=1, text_dim=32)   # weights never used
    hw = b200dit.synthetic.disc_head_weights(dim, 12)
    disc = b200dit.AptDiscriminator(eng, hw, tap_blocks=(1, 1, 1))
    gen = torch.Generator().manual_seed(22)
    B, L = 2, 200
    taps = [torch.randn(B, L, dim, generator=gen) * 1.5 + 0.2 for _ in range(3)]
    logit, feats = disc.heads([u.reshape(B * L, dim).cuda() for u in taps], B, L, return_features=True)
    r_logit, r_feats = DO.disc_heads(taps, hw, heads)
    assert (logit.cpu() - r_logit).abs().max() < TOL_LOGIT
    for a, r in zip(feats, r_feats):
        assert rel_l2(a.cpu(), r) < TOL_FEAT
