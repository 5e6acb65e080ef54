"""Generates tests/golden/*.pt by running the UNMODIFIED reference modules (oracle/ref_loader.py)
on seeded synthetic weights/inputs.  Container-only (needs /root/reference); the outputs are the
committed known-answer vectors that pin oracle/ and the CUDA engine on the GPU box.

    python oracle/make_golden.py

Fixtures (all small enough to commit; weights stored as fp16 = exactly the values used):
  dit_t2v_tiny.pt   WanModel t2v, dim 128 / 1 head / ffn 256 / 2 layers / text_dim 32, two items with
                    different grids, t and context lengths, seq_len padding  (model.py:502-563)
  dit_i2v_tiny.pt   same with model_type='i2v', in_dim 32 (y channel stack) and clip_fea (config 3 hooks)
  dit_block_1p3b.pt outputs only (weights by seed): ONE 1.3B-shaped block + head on [16,1,60,104]
                    (BASELINE.json configs[0]); weights regenerate from make_synthetic_weights(seed)
  vae_tiny.pt       WanVAE_ decoder, dim 8, z [16,3,6,8] -> [3,9,48,64]  (vae.py:544-568)
  vae_enc_tiny.pt   WanVAE_ encoder, dim 8, video [3,9,32,48] -> mu [16,3,4,6]  (vae.py:516-542)
                    (`python oracle/make_golden.py encode` regenerates only this file)
  solver_traj.pt    FlowUniPC / FlowDPMSolver++ trajectories (fm_solvers_unipc.py, fm_solvers.py) on a toy
                    velocity field: per case the timesteps and the latent after every step
                    (`python oracle/make_golden.py solvers` regenerates only this file)
  disc_tiny.pt      WanAPTDiscriminator.forward (seaweed_apt/model.py:123-186) over a 36-block tiny Wan: inputs,
                    logits and the three head tokens; weights regenerate from b200dit.synthetic by stored seed
                    (`python oracle/make_golden.py disc`)
  omni_audio_tiny.pt  OmniConditionsModule.process_audio (Omnihuman/omnihuman_wan_t2v.py:13-60): audio pre-net weights,
                    wav2vec-shaped features for T = 5 and T = 1, the reference's tokens (`python oracle/make_golden.py omni`)
  dit_grad_tiny.pt  gradients of the APT stage-1 loss through WanModel on dit_t2v_tiny's weights and inputs: the
                    parity target of the backward row, SURVEY 8f F1 (`python oracle/make_golden.py grads`)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dit_oracle as O, ref_loader, vae_oracle as VO  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def half(sd):
    return {k: v.half() for k, v in sd.items()}


def run_ref_dit(M, cfg, sd, x, t, ctx, seq_len, clip=None, y=None):
    m = M.WanModel(model_type="i2v" if cfg["i2v"] else "t2v", in_dim=cfg["in_dim"], dim=cfg["dim"],
                   ffn_dim=cfg["ffn_dim"], num_heads=cfg["num_heads"], num_layers=cfg["num_layers"],
                   text_dim=cfg["text_dim"], use_checkpoint=False).eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        return [o.clone() for o in m(x, t, ctx, seq_len=seq_len, clip_fea=clip, y=y)]


SOLVER_CASES = [("unipc", 10, 5.0), ("unipc", 50, 5.0), ("unipc", 4, 1.0), ("unipc", 1, 5.0),
                ("dpm++", 10, 5.0), ("dpm++", 50, 1.0), ("dpm++", 20, 3.0), ("dpm++", 2, 5.0)]


def make_solver_golden():
    """The UNMODIFIED reference schedulers, driven the way text2video.py:204-252 drives them."""
    from oracle import solver_oracle as SO
    U, D = ref_loader.load_reference_solvers()
    x0 = torch.randn(1, 16, 2, 6, 8, generator=torch.Generator().manual_seed(77))
    cases = []
    for kind, steps, shift in SOLVER_CASES:
        if kind == "unipc":
            s = U.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
            s.set_timesteps(steps, device="cpu", shift=shift)
            ts = s.timesteps
        else:
            s = D.FlowDPMSolverMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
            ts, _ = D.retrieve_timesteps(s, device="cpu", sigmas=D.get_sampling_sigmas(steps, shift))
        x, traj = x0.clone(), []
        for t in ts:
            x = s.step(SO.toy_velocity(x, t), t, x, return_dict=False)[0]
            traj.append(x.clone())
        cases.append(dict(kind=kind, steps=steps, shift=shift, timesteps=ts.clone(), sigmas=s.sigmas.clone(),
                          traj=torch.stack(traj)))
        print("solver", kind, steps, shift, float(x.std()))
    torch.save(dict(x0=x0, cases=cases), os.path.join(OUT, "solver_traj.pt"))


def make_encode_golden():
    """The UNMODIFIED reference WanVAE_.encode on seeded synthetic weights (vae.py:516-542)."""
    _, V = ref_loader.load_reference_modules()
    vsd = VO.make_synthetic_vae_weights(dim=8, seed=6, encoder=True)
    vae = V.WanVAE_(dim=8, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
                    temperal_downsample=[False, True, True]).eval()
    vae.load_state_dict(vsd, strict=True)
    video = (torch.rand(3, 9, 32, 48, generator=torch.Generator().manual_seed(8)) * 2 - 1).half().float()
    mean, std = torch.tensor(VO.VAE_MEAN), torch.tensor(VO.VAE_STD)
    with torch.no_grad():
        mu = vae.encode(video[None], [mean, 1.0 / std]).float()[0]
    torch.save(dict(dim=8, sd=half(vsd), video=video.half(), out=mu), os.path.join(OUT, "vae_enc_tiny.pt"))
    print("vae_enc_tiny", tuple(mu.shape), float(mu.std()))


DISC_CFG = dict(dim=128, ffn_dim=128, num_heads=1, num_layers=36, text_dim=32, in_dim=16, i2v=False, freq_dim=256)


def make_disc_golden():
    """The UNMODIFIED WanAPTDiscriminator (seaweed_apt/model.py:86-186) over a 36-block tiny backbone (the
    reference hooks blocks[15], [25], [35]).  Weights come from b200dit.synthetic with the stored seeds, so the
    fixture holds only inputs and the reference's outputs.  Two cases: one latent frame (timestep shift s = 1)
    and three (s = 12)."""
    import importlib
    syn = importlib.import_module("omnihuman-1-hack_b200.synthetic")
    from oracle import disc_oracle as DO
    M, _ = ref_loader.load_reference_modules()
    A = ref_loader.load_reference_apt()
    cfg = DISC_CFG
    seed_backbone, seed_heads = 4242, 4343
    sd = {k: v.float() for k, v in syn.dit_weights(cfg, seed_backbone, "cpu").items()}
    hw = syn.disc_head_weights(cfg["dim"], seed_heads)
    wan = M.WanModel(model_type="t2v", in_dim=cfg["in_dim"], dim=cfg["dim"], ffn_dim=cfg["ffn_dim"],
                     num_heads=cfg["num_heads"], num_layers=cfg["num_layers"], text_dim=cfg["text_dim"],
                     use_checkpoint=False).eval()
    wan.load_state_dict(sd, strict=True)
    disc = A.WanAPTDiscriminator(wan).eval()
    missing, unexpected = disc.load_state_dict(hw, strict=False)
    assert not unexpected and all(k.startswith("backbone.") for k in missing), (missing, unexpected)
    g = torch.Generator().manual_seed(99)
    cases = []
    for frames, t, pad in ((1, [0.3, 0.9], 0), (3, [0.25, 0.6], 0), (1, [0.7, 0.45], 8)):
        # third case: seq_len = tokens + 8 -- the block outputs carry 8 padded rows per item, which the heads see
        x = torch.randn(2, 16, frames, 8, 8, generator=g)
        ctx = [torch.randn(12, cfg["text_dim"], generator=g), torch.randn(7, cfg["text_dim"], generator=g)]
        tt = torch.tensor(t)
        seq_len = frames * 16 + pad
        with torch.no_grad():
            logit, feats = disc(x, tt, ctx, seq_len, return_features=True)
        o_logit, o_feats = DO.disc_forward(sd, hw, x, tt, ctx, seq_len, cfg["num_heads"])
        err = max(float((a - b).abs().max()) for a, b in zip([logit] + feats, [o_logit] + o_feats))
        print(f"disc golden frames={frames}: logit {logit.flatten().tolist()} oracle max abs err {err:.2e}")
        assert err < 1e-4
        cases.append(dict(x=x, t=tt, context=ctx, seq_len=seq_len, logit=logit.clone(),
                          feats=[f.clone() for f in feats]))
    torch.save(dict(cfg=cfg, seed_backbone=seed_backbone, seed_heads=seed_heads, cases=cases),
               os.path.join(OUT, "disc_tiny.pt"))


def make_omni_golden():
    """The UNMODIFIED OmniConditionsModule (Omnihuman/omnihuman_wan_t2v.py:13-94): its own constructor builds the
    audio pre-net, `process_audio` produces the tokens.  Weights rounded to fp16-representable
    values (the engine packs GEMM weights as fp16).  Cases: T = 5 frames (adjacent-frame concat) and T = 1."""
    from oracle import omni_oracle as OO
    R = ref_loader.load_reference_omni()
    torch.manual_seed(1234)
    m = R.OmniConditionsModule(model_dim=128, num_frames=5, audio_dim=32, pose_keypoints=4).eval()
    with torch.no_grad():
        for p in m.audio_processor.parameters():
            p.copy_(p.half().float())
    sd = {k: v.clone() for k, v in m.audio_processor.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    cases = []
    for B, T in ((2, 5), (3, 1)):
        feats = torch.randn(B, T, 32, generator=g)
        with torch.no_grad():
            out = m.process_audio(feats)          # (the module's forward() itself raises as shipped: `locals()` at :81
                                                  # includes `self`; process_audio is the part that executes)
        err = float((OO.process_audio(sd, feats) - out).abs().max())
        print(f"omni golden B={B} T={T}: out {tuple(out.shape)} oracle max abs err {err:.2e}")
        assert err < 1e-5
        cases.append(dict(feats=feats, out=out.clone()))
    torch.save(dict(sd=half(sd), model_dim=128, audio_dim=32, cases=cases), os.path.join(OUT, "omni_audio_tiny.pt"))


GRAD_KEYS = ["patch_embedding.weight", "text_embedding.0.weight", "time_projection.1.weight", "blocks.0.modulation",
             "blocks.0.self_attn.q.weight", "blocks.0.self_attn.norm_k.weight", "blocks.0.cross_attn.k.weight",
             "blocks.0.norm3.weight", "blocks.1.ffn.0.weight", "blocks.1.ffn.2.bias", "blocks.1.self_attn.o.weight",
             "head.modulation", "head.head.weight"]


def make_grad_golden():
    """Parity target for the NEXT scope row (backward of the student forward): gradients of the APT
    stage-1 loss (distilled_trainer.py:262-289: student forward at t = 1000, MSE against v_teacher) through the
    UNMODIFIED WanModel on the tiny t2v fixture's weights and inputs -- d loss / d x for both items, the gradients
    of GRAD_KEYS in full, and the gradient norm of every parameter."""
    g = torch.load(os.path.join(OUT, "dit_t2v_tiny.pt"))
    M, _ = ref_loader.load_reference_modules()
    cfg = dict(g["cfg"])
    sd = {k: v.float() for k, v in g["sd"].items()}
    m = M.WanModel(model_type="t2v", in_dim=cfg["in_dim"], dim=cfg["dim"], ffn_dim=cfg["ffn_dim"],
                   num_heads=cfg["num_heads"], num_layers=cfg["num_layers"], text_dim=cfg["text_dim"],
                   use_checkpoint=False).train()
    m.load_state_dict(sd, strict=True)
    gen = torch.Generator().manual_seed(123)
    x = [u.clone().requires_grad_(True) for u in g["x"]]
    t = torch.full((len(x),), 1000.0)
    v_teacher = [torch.randn(u.shape, generator=gen) for u in g["x"]]
    out = m(x, t, g["context"], seq_len=g["seq_len"])
    loss = sum(torch.nn.functional.mse_loss(o.float(), v) for o, v in zip(out, v_teacher))
    loss.backward()
    params = dict(m.named_parameters())
    rec = dict(t=t, v_teacher=v_teacher, loss=loss.detach().clone(), dx=[u.grad.clone() for u in x],
               grads={k: params[k].grad.clone() for k in GRAD_KEYS},
               grad_norms={k: float(p.grad.norm()) for k, p in params.items() if p.grad is not None})
    # the oracle's autograd on the same problem
    sd_o = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "freqs"}
    x_o = [u.clone().requires_grad_(True) for u in g["x"]]
    out_o = O.dit_forward(sd_o, x_o, t, g["context"], g["seq_len"], num_heads=cfg["num_heads"])
    loss_o = sum(torch.nn.functional.mse_loss(o, v) for o, v in zip(out_o, v_teacher))
    loss_o.backward()
    worst = max(float((sd_o[k].grad - rec["grads"][k]).norm() / rec["grads"][k].norm()) for k in GRAD_KEYS)
    worst_x = max(float((a.grad - b).norm() / b.norm()) for a, b in zip(x_o, rec["dx"]))
    print(f"grad golden: loss {float(loss):.6f} (oracle {float(loss_o):.6f}), worst rel-L2 weights {worst:.2e}, inputs {worst_x:.2e}")
    assert worst < 1e-4 and worst_x < 1e-4
    torch.save(rec, os.path.join(OUT, "dit_grad_tiny.pt"))


DEEP = dict(dim=128, ffn_dim=256, num_heads=1, num_layers=13, text_dim=32, in_dim=16, seed=21)


def make_grad_deep_golden():
    """The shipped WanModel evaluates the FFN of every block with block_idx > 10 under torch.no_grad() on the CPU
    (model.py:318-325): 13 blocks are the smallest model in which that shows.  Gradients of the same APT stage-1 loss
    through the UNMODIFIED WanModel (13 blocks, dim 128; weights regenerate from the seed): d loss / d x, the gradient
    norm of every parameter (None for the detached FFNs) and a few tensors in full."""
    M, _ = ref_loader.load_reference_modules()
    cfg = dict(DEEP)
    sd = O.make_synthetic_weights(cfg["dim"], cfg["ffn_dim"], cfg["num_heads"], cfg["num_layers"], in_dim=cfg["in_dim"],
                                  text_dim=cfg["text_dim"], seed=cfg["seed"])
    m = M.WanModel(model_type="t2v", in_dim=cfg["in_dim"], dim=cfg["dim"], ffn_dim=cfg["ffn_dim"],
                   num_heads=cfg["num_heads"], num_layers=cfg["num_layers"], text_dim=cfg["text_dim"],
                   use_checkpoint=False).train()
    m.load_state_dict({k: v.float() for k, v in sd.items()}, strict=True)
    gen = torch.Generator().manual_seed(77)
    x0 = [torch.randn(16, 1, 8, 16, generator=gen)]
    ctx = [torch.randn(29, 32, generator=gen)]
    vt = [torch.randn(16, 1, 8, 16, generator=gen)]
    t = torch.tensor([1000.0])
    x = [u.clone().requires_grad_(True) for u in x0]
    out = m(x, t, ctx, seq_len=32)
    loss = sum(torch.nn.functional.mse_loss(o.float(), v) for o, v in zip(out, vt))
    loss.backward()
    params = dict(m.named_parameters())
    keys = ["blocks.12.modulation", "blocks.12.self_attn.q.weight", "blocks.11.cross_attn.o.weight",
            "blocks.10.ffn.0.weight", "blocks.3.ffn.2.weight", "blocks.0.self_attn.v.weight", "time_projection.1.bias"]
    no_grad = sorted(k for k, p in params.items() if p.grad is None)
    assert no_grad == sorted(f"blocks.{i}.ffn.{j}.{w}" for i in (11, 12) for j in (0, 2) for w in ("weight", "bias")), no_grad
    rec = dict(cfg=cfg, x=x0, context=ctx, v_teacher=vt, t=t, seq_len=32, loss=loss.detach().clone(),
               out=[o.detach().clone() for o in out], dx=[u.grad.clone() for u in x],
               grads={k: params[k].grad.clone() for k in keys}, no_grad=no_grad,
               grad_norms={k: float(p.grad.norm()) for k, p in params.items() if p.grad is not None})
    sd_o = {k: v.float().clone().requires_grad_(True) for k, v in sd.items() if k != "freqs"}
    x_o = [u.clone().requires_grad_(True) for u in x0]
    out_o = O.dit_forward(sd_o, x_o, t, ctx, 32, num_heads=1, ffn_no_grad_from=11)
    loss_o = sum(torch.nn.functional.mse_loss(o, v) for o, v in zip(out_o, vt))
    loss_o.backward()
    worst = max(float((sd_o[k].grad - rec["grads"][k]).norm() / rec["grads"][k].norm()) for k in keys)
    assert all(sd_o[k].grad is None for k in no_grad)
    print(f"deep grad golden: loss {float(loss):.6f} (oracle {float(loss_o):.6f}), worst rel-L2 {worst:.2e}, "
          f"dx {float((x_o[0].grad - rec['dx'][0]).norm() / rec['dx'][0].norm()):.2e}, no grad: {len(no_grad)} tensors")
    assert worst < 1e-4
    torch.save(rec, os.path.join(OUT, "dit_grad_deep.pt"))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "grads":
        return make_grad_golden()
    if len(sys.argv) > 1 and sys.argv[1] == "grads_deep":
        return make_grad_deep_golden()
    if len(sys.argv) > 1 and sys.argv[1] == "omni":
        return make_omni_golden()
    if len(sys.argv) > 1 and sys.argv[1] == "disc":
        return make_disc_golden()
    if len(sys.argv) > 1 and sys.argv[1] == "solvers":
        return make_solver_golden()
    if len(sys.argv) > 1 and sys.argv[1] == "encode":
        return make_encode_golden()
    make_encode_golden()
    make_solver_golden()
    M, V = ref_loader.load_reference_modules()
    os.makedirs(OUT, exist_ok=True)
    g = torch.Generator().manual_seed(1234)
    rn = lambda *s: torch.randn(*s, generator=g)

    for name, i2v in (("dit_t2v_tiny", False), ("dit_i2v_tiny", True)):
        cfg = dict(dim=128, ffn_dim=256, num_heads=1, num_layers=2, text_dim=32, in_dim=32 if i2v else 16, i2v=i2v)
        sd = O.make_synthetic_weights(cfg["dim"], cfg["ffn_dim"], cfg["num_heads"], cfg["num_layers"],
                                      in_dim=cfg["in_dim"], text_dim=cfg["text_dim"], i2v=i2v, seed=7)
        x = [rn(16, 2, 8, 12), rn(16, 1, 6, 10)]
        y = [rn(16, 2, 8, 12), rn(16, 1, 6, 10)] if i2v else None
        clip = rn(2, 257, 1280) if i2v else None
        t = torch.tensor([999.0, 417.0])
        ctx = [rn(100, 32), rn(37, 32)]
        out = run_ref_dit(M, cfg, sd, x, t, ctx, 60, clip, y)
        torch.save(dict(cfg=cfg, sd=half(sd), x=x, y=y, clip_fea=clip, t=t, context=ctx, seq_len=60, out=out),
                   os.path.join(OUT, name + ".pt"))
        print(name, [tuple(o.shape) for o in out], float(out[0].std()))

    # one 1.3B-shaped block (configs[0]); weights regenerate from the seed, only I/O is stored
    cfg = dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=1, text_dim=4096, in_dim=16, i2v=False, seed=11)
    sd = O.make_synthetic_weights(1536, 8960, 12, 1, seed=11)
    x = [rn(16, 1, 60, 104)]
    t = torch.tensor([999.0])
    ctx = [rn(77, 4096)]
    # inputs are stored as fp16, so the reference is run on the rounded inputs: the vector is exact
    xh, ch = x[0].half(), ctx[0].half()
    out = run_ref_dit(M, cfg, sd, [xh.float()], t, [ch.float()], 1560)
    torch.save(dict(cfg=cfg, x=[xh], t=t, context=[ch], seq_len=1560, out=out),
               os.path.join(OUT, "dit_block_1p3b.pt"))
    print("dit_block_1p3b", tuple(out[0].shape), float(out[0].std()))

    vsd = VO.make_synthetic_vae_weights(dim=8, seed=5)
    vae = V.WanVAE_(dim=8, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
                    temperal_downsample=[False, True, True]).eval()
    vae.load_state_dict(vsd, strict=False)
    z = rn(16, 3, 6, 8)
    mean, std = torch.tensor(VO.VAE_MEAN), torch.tensor(VO.VAE_STD)
    with torch.no_grad():
        pix = vae.decode(z[None], [mean, 1.0 / std]).float().clamp_(-1, 1)[0]
    torch.save(dict(dim=8, sd=half(vsd), z=z, out=pix), os.path.join(OUT, "vae_tiny.pt"))
    print("vae_tiny", tuple(pix.shape), float(pix.std()))


if __name__ == "__main__":
    main()
