"""Caller loops over the engine (omnihuman-1-hack_b200/pipelines.py) against the same loops written over the CPU
oracle: WanT2V.generate's CFG + UniPC loop (text2video.py:231-252), the OmniHuman loop shape (DPM++, linear CFG
annealing, omnihuman_wan_t2v.py:395-444) on the i2v hooks (config 3), and the APT stage-1 item (generate.py:227-229,
distilled_trainer.py:262-289; config 4).  Tolerances: a K-step trajectory compounds the per-forward error (<= 1e-3,
amplified by the guidance scale in the CFG difference term): rel-L2 <= 2e-2 end to end; single evaluations 4e-3."""
import pytest
import torch

from conftest import rel_l2
from oracle import dit_oracle as O, solver_oracle as SO

pytestmark = pytest.mark.gpu

CFG = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64)


def _oracle_sample(sd, x0, ctx, ctx0, steps, shift, guide, solver, seq_len, heads, anneal=False, clip=None, y=None):
    if solver == "unipc":
        s = SO.UniPCOracle(shift=1.0)
        ts = s.set_timesteps(steps, shift=shift)
    else:
        s = SO.DPMppOracle(shift=1.0)
        ts = s.set_timesteps(sigmas=SO.get_sampling_sigmas(steps, shift))
    x = x0.clone()
    for i, t in enumerate(ts):
        tt = torch.tensor([float(t)])
        kw = dict(num_heads=heads, clip_fea=clip, y=[y] if y is not None else None)
        c = O.dit_forward(sd, [x], tt, [ctx], seq_len, **kw)[0]
        u = O.dit_forward(sd, [x], tt, [ctx0], seq_len, **kw)[0]
        g = guide * (1.0 - i / len(ts)) + 1.0 * (i / len(ts)) if anneal else guide
        v = O.cfg_combine(c, u, g)
        x = s.step(v[None], t, x[None])[0]
    return x


def test_sample_unipc_vs_oracle():
    import b200dit
    from b200dit import pipelines as P
    sd = O.make_synthetic_weights(**CFG, seed=3)
    eng = b200dit.DitEngine.from_state_dict(sd, num_heads=2)
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(16, 2, 8, 12, generator=g)
    ctx, ctx0 = torch.randn(40, 64, generator=g), torch.randn(13, 64, generator=g)
    out = P.sample(eng, [x0], [ctx], [ctx0], steps=6, shift=5.0, guide_scale=5.0, solver="unipc")[0].cpu()
    ref = _oracle_sample(sd, x0, ctx, ctx0, 6, 5.0, 5.0, "unipc", 48, 2)
    assert rel_l2(out, ref) < 2e-2
    out2 = P.sample(eng, [x0], [ctx], [ctx0], steps=6, shift=5.0, guide_scale=5.0, solver="unipc")[0].cpu()
    assert torch.equal(out, out2)                       # deterministic: graph replay + context cache change nothing


def test_sample_omni_loop_on_i2v_hooks():
    """config 3 shape: DPM++ order 2, shift 1.0, cfg 7.5 annealed to 1, clip_fea (audio stand-in) + y (pose stack)."""
    import b200dit
    from b200dit import pipelines as P
    sd = O.make_synthetic_weights(**CFG, in_dim=32, i2v=True, seed=4)
    eng = b200dit.DitEngine.from_state_dict(sd, num_heads=2)
    g = torch.Generator().manual_seed(1)
    x0, y = torch.randn(16, 2, 8, 12, generator=g), torch.randn(16, 2, 8, 12, generator=g)
    ctx, ctx0 = torch.randn(30, 64, generator=g), torch.randn(9, 64, generator=g)
    clip = torch.randn(1, 257, 1280, generator=g)
    out = P.sample(eng, [x0], [ctx], [ctx0], steps=5, shift=1.0, guide_scale=7.5, solver="dpm++", cfg_anneal=True,
                   clip_fea=clip, y=[y])[0].cpu()
    ref = _oracle_sample(sd, x0, ctx, ctx0, 5, 1.0, 7.5, "dpm++", 48, 2, anneal=True, clip=clip, y=y)
    assert rel_l2(out, ref) < 2e-2


def test_teacher_student_item_vs_oracle():
    import b200dit
    from b200dit import pipelines as P
    sd = O.make_synthetic_weights(**CFG, seed=5)
    eng = b200dit.DitEngine.from_state_dict(sd, num_heads=2)
    g = torch.Generator().manual_seed(2)
    noise = torch.randn(16, 1, 8, 12, generator=g)
    ctx, ctx0 = torch.randn(64, 64, generator=g), torch.randn(64, 64, generator=g)
    vt, vs, loss = P.teacher_student_item(eng, noise, ctx, ctx0, guide_scale=7.5, seq_len=24)
    c = O.dit_forward(sd, [noise], torch.tensor([999.0]), [ctx], 24, num_heads=2)[0]
    u = O.dit_forward(sd, [noise], torch.tensor([999.0]), [ctx0], 24, num_heads=2)[0]
    s = O.dit_forward(sd, [noise], torch.tensor([1000.0]), [ctx], 24, num_heads=2)[0]
    ref_t = O.cfg_combine(c, u, 7.5)
    assert rel_l2(vt.cpu(), ref_t) < 6e-3               # 7.5 x the difference term
    assert rel_l2(vs.cpu(), s) < 1e-3
    ref_loss = float(torch.mean((s - ref_t) ** 2))
    assert abs(float(loss) - ref_loss) <= 2e-2 * abs(ref_loss) + 1e-6


class _StandInT2V:
    """The attributes `WanT2V.generate` reads from `self` (text2video.py:63-80,101,108), around tiny random-weight
    modules; the text encoder is a deterministic stub (T5 is outside the path)."""

    class _Mod:
        def __init__(self, sd, **attrs):
            self._sd = sd
            self.__dict__.update(attrs)

        def state_dict(self):
            return self._sd

    def __init__(self, dit_sd, vae_sd):
        import types
        self.device, self.rank, self.t5_cpu, self.sp_size = torch.device("cuda"), 0, True, 1
        self.vae_stride, self.patch_size, self.sample_neg_prompt = (4, 8, 8), (1, 2, 2), "bad"
        self.model = self._Mod(dit_sd, dim=CFG["dim"], ffn_dim=CFG["ffn_dim"], num_heads=CFG["num_heads"],
                               num_layers=CFG["num_layers"], in_dim=16, out_dim=16, text_dim=CFG["text_dim"],
                               text_len=512, freq_dim=256, model_type="t2v", eps=1e-6)
        self.vae = types.SimpleNamespace(model=self._Mod(vae_sd, z_dim=16))
        self.prompts = []

    def text_encoder(self, texts, device):
        self.prompts.append(texts[0])
        g = torch.Generator().manual_seed(sum(map(ord, texts[0])))
        return [torch.randn(5 + len(texts[0]), CFG["text_dim"], generator=g).to(device)]

    def generate(self, *a, **k):
        raise AssertionError("the reference loop must not run")


def test_install_t2v_generate_signature_and_result():
    """`install_t2v` keeps WanT2V.generate's signature (text2video.py:111-121): same seeded noise, CFG + solver loop,
    decode on rank 0; result against the oracle loop + the VAE oracle on the same noise."""
    import b200dit
    from oracle import vae_oracle as VO
    sd = O.make_synthetic_weights(**CFG, seed=5)
    vsd = VO.make_synthetic_vae_weights(dim=8, seed=2)
    pipe = _StandInT2V(sd, vsd)
    eng, vae = b200dit.install_t2v(pipe)
    video = pipe.generate("a cat", size=(64, 48), frame_num=5, shift=3.0, sample_solver="dpm++", sampling_steps=4,
                          guide_scale=4.0, seed=11)
    assert video.shape == (3, 5, 48, 64) and pipe.prompts == ["a cat", "bad"]
    again = pipe.generate("a cat", size=(64, 48), frame_num=5, shift=3.0, sample_solver="dpm++", sampling_steps=4,
                          guide_scale=4.0, seed=11)
    assert torch.equal(video, again)
    other = pipe.generate("a cat", size=(64, 48), frame_num=5, shift=3.0, sample_solver="dpm++", sampling_steps=4,
                          guide_scale=4.0, seed=12, n_prompt="worse")
    assert not torch.equal(video, other) and pipe.prompts[-1] == "worse"
    with pytest.raises(NotImplementedError):
        pipe.generate("a cat", sample_solver="euler")
    # oracle: same noise (device generator, text2video.py:167-169,186-195), same loop, fp32 CPU
    gen = torch.Generator(device="cuda"); gen.manual_seed(11)
    noise = torch.randn(16, 2, 6, 8, dtype=torch.float32, device="cuda", generator=gen).cpu()
    ctx, ctx0 = pipe.text_encoder(["a cat"], "cpu")[0], pipe.text_encoder(["bad"], "cpu")[0]
    x0 = _oracle_sample(sd, noise, ctx, ctx0, 4, 3.0, 4.0, "dpm++", 2 * 3 * 4, CFG["num_heads"])
    ref = VO.vae_decode(vsd, x0)
    err, rel = float((video.cpu() - ref).abs().max()), rel_l2(video.cpu(), ref)
    print(f"install_t2v: video max-abs {err:.3e} rel-L2 {rel:.3e}")
    assert rel < 1e-2 and err < 3e-2                     # measured 8e-4 / 2.8e-3 on B200
    b200dit.uninstall(pipe)
    with pytest.raises(AssertionError):
        pipe.generate("a cat")
