"""Caller-side loops of the reference, restated over the B200 engine (SURVEY.md section 8 "next" rows:
the callers either side of the DiT forward).  Host logic only: every tensor operation is an engine call
(`DitEngine.forward` / `forward_cfg`, `VaeEngine.decode`) or one fused scheduler launch (solvers.py).

  sample()               WanT2V.generate's denoise loop (seaweed_apt/wan/text2video.py:202-252): cond + uncond
                         forward, CFG, scheduler.step; with `cfg_anneal=True` the OmniHuman loop's linear CFG
                         annealing (Omnihuman/omnihuman_wan_t2v.py:395-444: cfg_i = s (1 - i/N) + i/N) and its
                         DPM++ scheduler (:171-176); `clip_fea` / `y` feed the i2v hooks (model.py:511-512, 534-537)
                         that stand in for the audio-token and pose-stack streams
  teacher_student_item() APT stage-1 item (seaweed_apt/generate.py:205-229 + distilled_trainer.py:262-289): teacher
                         cond + uncond at t = 999, v_teacher = u + 7.5 (c - u), student at t = 1000, MSE
  student_step()         the APT stage-1 training step (distilled_trainer.py:241-301): student forward at t = 1000,
                         MSE against the cached v_teacher, backward -- forward AND backward on the engine
  generate_video()       sample() + WanVAE.decode (text2video.py:258-259)
"""
import torch

from . import parallel
from .solvers import (FlowDPMSolverMultistepScheduler, FlowUniPCMultistepScheduler, get_sampling_sigmas,
                      retrieve_timesteps)


def make_scheduler(solver, steps, shift, device):
    """text2video.py:204-221."""
    if solver == "unipc":
        sch = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
        sch.set_timesteps(steps, device=device, shift=shift)
        return sch, sch.timesteps
    if solver == "dpm++":
        sch = FlowDPMSolverMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
        ts, _ = retrieve_timesteps(sch, device=device, sigmas=get_sampling_sigmas(steps, shift))
        return sch, ts
    raise NotImplementedError("Unsupported solver.")


@torch.no_grad()
def sample(engine, noise, context, context_null, steps=50, shift=5.0, guide_scale=5.0, solver="unipc",
           seq_len=None, clip_fea=None, y=None, cfg_anneal=False, callback=None, check_overflow=True):
    """Denoises `noise` (list of [16,T,h,w] fp32, all the same shape) and returns the list of x0 latents.
    Every sample carries its own scheduler state; all samples are co-batched in one engine call per step.
    `check_overflow`: one read of the engine's overflow guard after the last step (the engine computes with fp16
    operands, like the reference's blocks, model.py:540); raises FloatingPointError instead of returning NaN latents."""
    xs = [n.to(engine.device, torch.float32) for n in noise]
    n = len(xs)
    if seq_len is None:
        _, T, h, w = xs[0].shape
        seq_len = T * (h // 2) * (w // 2)                                        # text2video.py:161-164
    scheds = [make_scheduler(solver, steps, shift, engine.device) for _ in range(n)]
    ts_host = scheds[0][1].cpu().tolist()                                        # one read per trajectory
    for i, t in enumerate(ts_host):
        s = guide_scale * (1.0 - i / len(ts_host)) + 1.0 * (i / len(ts_host)) if cfg_anneal else guide_scale
        tt = torch.full((n,), float(t), device=engine.device)
        v = engine.forward_cfg(xs, tt, context, context_null, seq_len, s, clip_fea=clip_fea, y=y)
        xs = [sch.step(vi.unsqueeze(0), t, xi.unsqueeze(0), return_dict=False)[0].squeeze(0)
              for (sch, _), xi, vi in zip(scheds, xs, v)]
        if callback is not None:
            callback(i, t, xs)
    if check_overflow:
        bad = engine.nonfinite_rows()
        if bad:
            raise FloatingPointError(f"{bad} residual-stream rows overflowed the fp16 operand range during sampling")
    return xs


@torch.no_grad()
def sample_cfg_parallel(engine, noise, context, context_null, steps=50, shift=5.0, guide_scale=5.0, solver="unipc",
                        seq_len=None, cfg_anneal=False, check_overflow=True):
    """sample() with the cond / uncond pair of every step split over a rank pair (2k, 2k+1) -- the north star's
    "cond/uncond pair as independent batch items across GPUs", for single-video latency: at T = 21 one forward
    already fills a B200, so co-batching the pair buys nothing and splitting it halves the step time.
    Rank 2k runs the conditional forward, rank 2k+1 the unconditional one (text2video.py:238-241); ONE all_gather
    of the predictions inside the pair per step (8.4 MB at T = 21), then both ranks form u + s (c - u) (:243-244)
    and take the same scheduler step, so the latents stay bit-identical on the two ranks and nothing else is
    exchanged.  Every rank of a pair passes the same `noise` / contexts; returns x0 on both."""
    from .solvers import _lincomb
    role = parallel.rank() % 2
    xs = [n.to(engine.device, torch.float32) for n in noise]
    n = len(xs)
    if seq_len is None:
        _, T, h, w = xs[0].shape
        seq_len = T * (h // 2) * (w // 2)
    ctx = list(context) if role == 0 else list(context_null)
    if len(ctx) == 1 and n > 1:
        ctx = ctx * n
    scheds = [make_scheduler(solver, steps, shift, engine.device) for _ in range(n)]
    ts_host = scheds[0][1].cpu().tolist()
    for i, t in enumerate(ts_host):
        s = guide_scale * (1.0 - i / len(ts_host)) + 1.0 * (i / len(ts_host)) if cfg_anneal else guide_scale
        tt = torch.full((n,), float(t), device=engine.device)
        mine = torch.stack(engine.forward(xs, tt, ctx, seq_len))
        cond, uncond = parallel.exchange_pair(mine)
        v = _lincomb([cond, uncond], [[float(s), 1.0 - float(s)]], cond)[0]              # u + s (c - u)
        xs = [sch.step(v[j].unsqueeze(0), t, xs[j].unsqueeze(0), return_dict=False)[0].squeeze(0)
              for j, (sch, _) in enumerate(scheds)]
    if check_overflow:
        bad = engine.nonfinite_rows()
        if bad:
            raise FloatingPointError(f"{bad} residual-stream rows overflowed the fp16 operand range during sampling")
    return xs


@torch.no_grad()
def teacher_student_item(engine, noise, context, context_null, guide_scale=7.5, t_teacher=999.0, t_student=1000.0,
                         seq_len=1560, student_engine=None):
    """One distillation item.  With a single weight replica (teacher == student at step 0,
    distilled_trainer.py:64) the three forwards run as ONE co-batched call of 3 items."""
    x = noise.to(engine.device, torch.float32)
    if student_engine is None or student_engine is engine:
        t = torch.tensor([t_teacher, t_teacher, t_student], device=engine.device)
        c, u, s = engine.forward([x, x, x], t, [context, context_null, context], seq_len)
    else:
        t = torch.tensor([t_teacher], device=engine.device)
        v = engine.forward_cfg([x], t, [context], [context_null], seq_len, guide_scale)[0]
        s = student_engine.forward([x], torch.tensor([t_student], device=engine.device), [context], seq_len)[0]
        return v, s, torch.mean((s - v) ** 2)
    from .solvers import _lincomb
    g = float(guide_scale)                                                       # one fused launch: u + g (c - u) and s - v
    v_teacher, diff = _lincomb([c, u, s], [[g, 1.0 - g, 0.0], [-g, g - 1.0, 1.0]], c)      # generate.py:229
    loss = torch.mean(diff * diff)                                               # distilled_trainer.py:289
    return v_teacher, s, loss


def student_step(engine, noises, contexts, v_teachers, t_student=1000.0, seq_len=1560, ffn_grad_blocks=11,
                 grad_accum=1):
    """`training_step` of distilled_trainer.py:241-301 on the engine: v = student(noise, t = num_train_timesteps,
    context) (:265-278), loss = mse(v, v_teacher) / gradient_accumulation_steps (:289), loss.backward() (:301).
    The items (one latent grid) are co-batched; d loss / d v = 2 (v - v_teacher) / numel is one fused launch.
    Parameter gradients ACCUMULATE in the engine (`engine.read_grad`, `engine.zero_grad`).  Returns the per-item
    losses (device tensor, un-divided like the `loss_value` the reference logs at :304)."""
    from .solvers import _lincomb
    xs = [u.to(engine.device, torch.float32) for u in noises]
    n = len(xs)
    t = torch.full((n,), float(t_student), device=engine.device)
    outs = engine.train_forward(xs, t, contexts, seq_len)
    numel = outs[0].numel()
    c = 2.0 / (numel * n * grad_accum)            # F.mse_loss over the [n, 16, T, h, w] batch: mean over all elements
    douts, losses = [], []
    for o, v in zip(outs, v_teachers):
        d = _lincomb([o, v.to(engine.device, torch.float32)], [[c, -c]], o)[0]
        douts.append(d)
        losses.append((d * d).sum() * (numel * n * grad_accum / 2.0) ** 2 / numel)
    engine.backward(douts, ffn_grad_blocks=ffn_grad_blocks, want_dx=False)
    return torch.stack(losses)


@torch.no_grad()
def teacher_student_sweep(engine, noises, contexts, context_null, **kw):
    """generate.py:209-232 over independent seeds, items sharded i % world == rank, one all_gather of
    (v_teacher, v_student) and of the losses (default mode)."""
    idx = parallel.shard_indices(len(noises))
    vt, vs, ls = [], [], []
    for i in idx:
        a, b, l = teacher_student_item(engine, noises[i], contexts[i], context_null, **kw)
        vt.append(a); vs.append(b); ls.append(l.reshape(1))
    n = len(noises)
    return (parallel.gather_items(vt, n), parallel.gather_items(vs, n), parallel.gather_items(ls, n))


@torch.no_grad()
def teacher_student_pair_split(engine, noises, contexts, context_null, guide_scale=7.5, t_teacher=999.0,
                               t_student=1000.0, seq_len=1560):
    """The other sharding of SURVEY 8d config 4 (the north star's literal reading): the teacher's cond / uncond
    pair of ONE item is split over a rank pair.  Rank 2k runs teacher-cond and the student co-batched, rank 2k+1
    teacher-uncond and sends its [16,T,h,w] prediction (0.4 MB at 480p) to rank 2k, which forms
    v_teacher = u + s (c - u) (generate.py:229) and the MSE (distilled_trainer.py:289).  Pair k owns items
    i % (world / 2) == k; one all_gather of (v_teacher, v_student, loss) at the end, as in the default mode.
    Returns the same triple as teacher_student_sweep.  Slower than the default mode (two forwards on one rank
    against one on the other); kept because it is the sharding the north star names."""
    r, w = parallel.rank(), parallel.world_size()
    if w % 2:
        raise ValueError("pair-split needs an even number of ranks")
    pairs, k, role = w // 2, r // 2, r % 2
    n = len(noises)
    mine = list(range(k, n, pairs))
    vt, vs, ls = [], [], []
    for i in mine:
        x = noises[i].to(engine.device, torch.float32)
        if role == 0:
            t = torch.tensor([t_teacher, t_student], device=engine.device)
            c, s = engine.forward([x, x], t, [contexts[i], contexts[i]], seq_len)
            u = parallel.recv_from(c, r + 1)
            v = u + guide_scale * (c - u)
            vt.append(v); vs.append(s); ls.append(torch.mean((s - v) ** 2).reshape(1))
        else:
            u = engine.forward([x], torch.tensor([t_teacher], device=engine.device), [context_null], seq_len)[0]
            parallel.send_to(u, r - 1)
    counts = [len(range(q // 2, n, pairs)) if q % 2 == 0 else 0 for q in range(w)]
    out = []
    for local in (vt, vs, ls):
        per_rank = parallel.gather_from_ranks(local, counts)
        res = [None] * n
        for q in range(0, w, 2):
            for j, i in enumerate(range(q // 2, n, pairs)):
                res[i] = per_rank[q][j]
        out.append(res)
    return tuple(out)


@torch.no_grad()
def generate_video(engine, vae, noise, context, context_null, **kw):
    """Denoise + decode (text2video.py:231-259): returns (latents, list of [3, 1+4(T-1), 8h, 8w] videos)."""
    x0 = sample(engine, noise, context, context_null, **kw)
    return x0, vae.decode(x0)
