"""CPU suite: the oracle against the committed golden vectors (produced by the unmodified reference
modules, oracle/make_golden.py) and, when /root/reference is mounted, against the reference itself."""
import os

import pytest
import torch

from conftest import GOLDEN, rel_l2
from oracle import dit_oracle as O, ref_loader, vae_oracle as VO


def _load(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=True)


@pytest.mark.parametrize("name", ["dit_t2v_tiny.pt", "dit_i2v_tiny.pt"])
def test_dit_oracle_vs_golden(name):
    g = _load(name)
    sd = {k: v.float() for k, v in g["sd"].items()}
    out = O.dit_forward(sd, g["x"], g["t"], g["context"], g["seq_len"], clip_fea=g["clip_fea"], y=g["y"],
                        num_heads=g["cfg"]["num_heads"])
    for o, r in zip(out, g["out"]):
        assert o.shape == r.shape and o.dtype == torch.float32
        assert rel_l2(o, r) < 1e-5


def test_dit_oracle_block_1p3b_vs_golden():
    g = _load("dit_block_1p3b.pt")
    sd = O.make_synthetic_weights(1536, 8960, 12, 1, seed=g["cfg"]["seed"])
    out = O.dit_forward(sd, [g["x"][0].float()], g["t"], [g["context"][0].float()], g["seq_len"])
    assert rel_l2(out[0], g["out"][0]) < 1e-5


def test_vae_oracle_vs_golden_and_two_pass():
    g = _load("vae_tiny.pt")
    sd = {k: v.float() for k, v in g["sd"].items()}
    a = VO.vae_decode(sd, g["z"])
    assert a.shape == g["out"].shape
    assert float((a - g["out"]).abs().max()) < 1e-5
    b = VO.vae_decode(sd, g["z"], chunks=[1, g["z"].shape[1] - 1])     # the engine's two-pass schedule
    assert float((b - g["out"]).abs().max()) < 1e-5


def test_seq_len_overflow_raises():
    g = _load("dit_t2v_tiny.pt")
    sd = {k: v.float() for k, v in g["sd"].items()}
    with pytest.raises(AssertionError):
        O.dit_forward(sd, g["x"][:1], g["t"][:1], g["context"][:1], seq_len=10, num_heads=1)


def test_rope_and_sinusoid_layouts():
    """SURVEY App. E identities: pair split (22,21,21), [cos | sin] layout."""
    ang = O.rope_angles(128, (3, 4, 5))
    assert ang.shape == (60, 64)
    tok = (2 * 4 + 3) * 5 + 1                            # (f,h,w) = (2,3,1)
    assert abs(float(ang[tok, 1]) - 2 * 10000 ** (-2 / 44)) < 1e-12
    assert abs(float(ang[tok, 22 + 2]) - 3 * 10000 ** (-4 / 42)) < 1e-12
    assert abs(float(ang[tok, 43 + 5]) - 1 * 10000 ** (-10 / 42)) < 1e-12
    s = O.sinusoid_256(256, torch.tensor([10.0]))
    assert abs(float(s[0, 0]) - float(torch.cos(torch.tensor(10.0, dtype=torch.float64)))) < 1e-12
    assert abs(float(s[0, 128]) - float(torch.sin(torch.tensor(10.0, dtype=torch.float64)))) < 1e-12


def test_flop_formulas_match_survey():
    assert abs(O.dit_flops(1560) / 1e12 - 4.652) < 0.01
    assert abs(O.dit_flops(32760) / 1e12 - 283.0) < 0.5
    assert abs(VO.vae_decode_flops(1) / 1e12 - 4.35) < 0.05


@pytest.mark.skipif(ref_loader.find_reference() is None, reason="reference tree not mounted (container-only check)")
def test_oracle_vs_live_reference():
    M, V = ref_loader.load_reference_modules()
    sd = O.make_synthetic_weights(256, 512, 2, 2, text_dim=64, seed=9)
    m = M.WanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, use_checkpoint=False).eval()
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(2)
    x = [torch.randn(16, 2, 8, 12, generator=g), torch.randn(16, 1, 6, 10, generator=g)]
    ctx = [torch.randn(100, 64, generator=g), torch.randn(37, 64, generator=g)]
    t = torch.tensor([999.0, 3.0])
    with torch.no_grad():
        ref = m(x, t, ctx, seq_len=60)
    mine = O.dit_forward(sd, x, t, ctx, seq_len=60, num_heads=2)
    for a, b in zip(mine, ref):
        assert rel_l2(a, b) < 1e-5
    vsd = VO.make_synthetic_vae_weights(dim=8, seed=4)
    vae = V.WanVAE_(dim=8, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
                    temperal_downsample=[False, True, True]).eval()
    vae.load_state_dict(vsd, strict=False)
    z = torch.randn(16, 3, 4, 6, generator=g)
    with torch.no_grad():
        pix = vae.decode(z[None], [torch.tensor(VO.VAE_MEAN), 1.0 / torch.tensor(VO.VAE_STD)])[0].clamp(-1, 1)
    assert float((VO.vae_decode(vsd, z) - pix).abs().max()) < 1e-5


def test_package_flop_counters_match_oracle():
    """bench.py / tools use omnihuman-1-hack_b200/flops.py; the oracle keeps independent copies (SURVEY 8d formulas)."""
    import b200dit
    from oracle import dit_oracle as O, vae_oracle as VO
    for L in (1560, 6240, 32760):
        for cached in (False, True):
            assert b200dit.flops.dit_forward_flops(L, context_cached=cached) == O.dit_flops(L, context_cached=cached)
    for T in (1, 2, 21):
        assert abs(b200dit.flops.vae_decode_flops(T) / VO.vae_decode_flops(T) - 1) < 1e-9
    assert abs(b200dit.flops.dit_forward_flops(1560) / 4.652e12 - 1) < 1e-3          # 4.652 TFLOP / forward


def test_vae_encode_oracle_vs_golden():
    """Encoder restatement (oracle/vae_oracle.py:vae_encode) vs the unmodified reference WanVAE_.encode fixture."""
    g = _load("vae_enc_tiny.pt")
    sd = {k: v.float() for k, v in g["sd"].items()}
    mu = VO.vae_encode(sd, g["video"].float(), dim=g["dim"])
    assert mu.shape == g["out"].shape
    assert float((mu - g["out"]).abs().max()) < 1e-5
    import b200dit
    syn = b200dit.synthetic.vae_decoder_weights(dim=8, seed=0, encoder=True)
    ref = VO.make_synthetic_vae_weights(dim=8, seed=0, encoder=True)
    assert set(syn) == set(ref) and all(syn[k].shape == ref[k].shape for k in syn)


def _disc_weights(g):
    import b200dit
    sd = {k: v.float() for k, v in b200dit.synthetic.dit_weights(g["cfg"], g["seed_backbone"], "cpu").items()}
    return sd, b200dit.synthetic.disc_head_weights(g["cfg"]["dim"], g["seed_heads"])


def test_disc_oracle_vs_golden():
    """The discriminator restatement (backbone at the shifted timestep, taps 16/26/36, three heads, final_proj)
    against the UNMODIFIED WanAPTDiscriminator's logits and tokens (oracle/make_golden.py disc)."""
    from oracle import disc_oracle as DO
    g = torch.load(os.path.join(GOLDEN, "disc_tiny.pt"))
    sd, hw = _disc_weights(g)
    for case in g["cases"]:
        logit, feats = DO.disc_forward(sd, hw, case["x"], case["t"], case["context"], case["seq_len"],
                                       g["cfg"]["num_heads"])
        assert torch.allclose(logit, case["logit"], atol=2e-5)
        for a, r in zip(feats, case["feats"]):
            assert rel_l2(a, r) < 1e-5
    assert float(DO.timestep_shift(torch.tensor(0.5), 1)) == 0.5
    assert abs(float(DO.timestep_shift(torch.tensor(0.5), 3)) - 6.0 / 6.5) < 1e-7


@pytest.mark.skipif(ref_loader.find_reference() is None, reason="reference tree not mounted (container-only check)")
@pytest.mark.parametrize("qk_norm", [True, False])
def test_disc_head_oracle_vs_live_reference(qk_norm):
    """One head at several heads / ragged token counts against the reference class executed in place."""
    from oracle import disc_oracle as DO
    A = ref_loader.load_reference_apt()
    torch.manual_seed(5)
    blk = A.WanCrossAttentionDiscriminatorBlock(256, 2, qk_norm=qk_norm, eps=1e-6).eval()
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(0.05 * torch.randn_like(p))
    x = torch.randn(3, 37, 256) * 2.0 + 0.3
    w = {"h." + k: v for k, v in blk.state_dict().items()}
    with torch.no_grad():
        ref = blk(x)
    out = DO.disc_head(x, w, "h.", 2, qk_norm=qk_norm)
    assert out.shape == ref.shape == (3, 1, 256)
    assert rel_l2(out, ref) < 1e-6


def test_dit_oracle_gradients_vs_golden():
    """Parity target of the next scope row (backward of the student forward): the oracle's autograd
    through the APT stage-1 loss (student forward at t = 1000, MSE against v_teacher; distilled_trainer.py:262-289)
    against the gradients of the UNMODIFIED WanModel (oracle/make_golden.py grads)."""
    g = torch.load(os.path.join(GOLDEN, "dit_t2v_tiny.pt"))
    r = torch.load(os.path.join(GOLDEN, "dit_grad_tiny.pt"))
    sd = {k: v.float().clone().requires_grad_(True) for k, v in g["sd"].items() if k != "freqs"}
    x = [u.clone().requires_grad_(True) for u in g["x"]]
    out = O.dit_forward(sd, x, r["t"], g["context"], g["seq_len"], num_heads=g["cfg"]["num_heads"])
    loss = sum(torch.nn.functional.mse_loss(o, v) for o, v in zip(out, r["v_teacher"]))
    loss.backward()
    assert abs(float(loss) - float(r["loss"])) < 1e-5
    for a, b in zip(x, r["dx"]):
        assert rel_l2(a.grad, b) < 1e-4
    for k, ref in r["grads"].items():
        assert rel_l2(sd[k].grad, ref) < 1e-4, k
    for k, n in r["grad_norms"].items():
        assert abs(float(sd[k].grad.norm()) - n) <= 1e-4 * max(n, 1e-6) + 1e-9, k


def test_dit_oracle_gradients_deep_model_detached_ffns():
    """13 blocks: the shipped WanModel runs the FFN of blocks 11 and 12 under no_grad (model.py:318-325).  The oracle's
    `ffn_no_grad_from=11` against the gradients of the UNMODIFIED reference (oracle/make_golden.py grads_deep)."""
    r = torch.load(os.path.join(GOLDEN, "dit_grad_deep.pt"))
    c = r["cfg"]
    sd = O.make_synthetic_weights(c["dim"], c["ffn_dim"], c["num_heads"], c["num_layers"], in_dim=c["in_dim"],
                                  text_dim=c["text_dim"], seed=c["seed"])
    sd = {k: v.float().clone().requires_grad_(True) for k, v in sd.items() if k != "freqs"}
    x = [u.clone().requires_grad_(True) for u in r["x"]]
    out = O.dit_forward(sd, x, r["t"], r["context"], r["seq_len"], num_heads=1, ffn_no_grad_from=11)
    loss = sum(torch.nn.functional.mse_loss(o, v) for o, v in zip(out, r["v_teacher"]))
    loss.backward()
    assert rel_l2(out[0].detach(), r["out"][0]) < 1e-5 and abs(float(loss) - float(r["loss"])) < 1e-5
    assert rel_l2(x[0].grad, r["dx"][0]) < 1e-4
    for k, ref in r["grads"].items():
        assert rel_l2(sd[k].grad, ref) < 1e-4, k
    assert sorted(k for k, v in sd.items() if v.grad is None) == r["no_grad"]
    for k, n in r["grad_norms"].items():
        assert abs(float(sd[k].grad.norm()) - n) <= 1e-4 * max(n, 1e-6) + 1e-9, k


def test_omni_audio_oracle_vs_golden_and_live_reference():
    """oracle/omni_oracle.process_audio vs the fixture made by the unmodified OmniConditionsModule
    (Omnihuman/omnihuman_wan_t2v.py:13-60) and, where the tree is mounted, vs the live class."""
    from oracle import omni_oracle as OO, ref_loader
    g = torch.load(os.path.join(GOLDEN, "omni_audio_tiny.pt"), map_location="cpu", weights_only=True)
    sd = {k: v.float() for k, v in g["sd"].items()}
    for c in g["cases"]:
        out = OO.process_audio(sd, c["feats"])
        assert out.shape == c["out"].shape
        assert float((out - c["out"]).abs().max()) < 1e-5
    if ref_loader.find_reference() is None:
        return
    R = ref_loader.load_reference_omni()
    m = R.OmniConditionsModule(model_dim=64, num_frames=3, audio_dim=16, pose_keypoints=2).eval()
    feats = torch.randn(2, 3, 16, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        ref = m.process_audio(feats)
    assert float((OO.process_audio(m.audio_processor.state_dict(), feats) - ref).abs().max()) < 1e-5
