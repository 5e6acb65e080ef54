"""BASELINE.json configs 3, 4 and 5 at full size on the B200 engine (synthetic weights / inputs of the
reference's shapes).  One JSON line per config; launch under torchrun for N > 1.

    python tools/run_configs.py 3 [--steps 50]     OmniHuman omni-conditions loop shape on the i2v hooks, T = 21
    torchrun ... tools/run_configs.py 4            APT stage-1 items sharded over the ranks (2 GPUs in BASELINE)
    torchrun ... tools/run_configs.py 5            81-frame denoise + WanVAE decode per rank, one gather of latents
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import b200dit  # noqa: E402
from b200dit import parallel, pipelines as P, synthetic  # noqa: E402
from bench import CFG_13B, make_device_weights  # noqa: E402


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, parallel.max_over_ranks(e0.elapsed_time(e1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", type=int, choices=[3, 4, 5])
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--frames", type=int, default=21)
    ap.add_argument("--items", type=int, default=16)
    ap.add_argument("--layers", type=int, default=30)
    ap.add_argument("--cfg-parallel", action="store_true",
                    help="config 5: one video per rank PAIR, cond / uncond forwards on the two ranks, one exchange per step")
    ap.add_argument("--pipe-decode", action="store_true",
                    help="config 5 with --cfg-parallel: the rank pair decodes its video together (time-chunked VAE decode)")
    ap.add_argument("--train", action="store_true",
                    help="config 4: the student training step (forward + backward on the engine, distilled_trainer.py:241-301)")
    ap.add_argument("--batch", type=int, default=1, help="config 4 --train: items per training step (reference: 1)")
    ap.add_argument("--shim", action="store_true",
                    help="config 4 --train: the whole trainer-visible step through wan_shim.install on a module with live "
                         "parameters: forward, loss.backward() (engine), AdamW step, weight reload on the next call")
    ap.add_argument("--pair-split", action="store_true",
                    help="config 4: teacher cond / uncond of one item on a rank pair (one send per item)")
    a = ap.parse_args()
    rank, world = parallel.init()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    g = torch.Generator().manual_seed(1000 + rank)
    rn = lambda *s: torch.randn(*s, generator=g)
    cfg = dict(CFG_13B, num_layers=a.layers)
    T = a.frames
    L = 1560 * T

    if a.config == 3:
        cfg3 = dict(cfg, in_dim=32)
        eng = b200dit.DitEngine(**cfg3, i2v=True, device=dev)
        eng.load_state_dict(make_device_weights(cfg3, 0, dev, i2v=True))
        x0, y = rn(16, T, 60, 104).to(dev), rn(16, T, 60, 104).to(dev)
        ctx, ctx0 = rn(200, 4096).bfloat16().to(dev), rn(120, 4096).bfloat16().to(dev)
        clip = rn(1, 257, 1280).to(dev)
        run = lambda n: P.sample(eng, [x0], [ctx], [ctx0], steps=n, shift=1.0, guide_scale=7.5, solver="dpm++",
                                 cfg_anneal=True, clip_fea=clip, y=[y])
        run(3)                                                # warm-up: eager + capture + replay
        out, ms = timed(lambda: run(a.steps))
        fl = 2 * a.steps * eng.last_flops / 2                 # last_flops counts the co-batched pair
        line = {"config": 3, "workload": f"i2v-hook DiT (in_dim 32, clip_fea 257 tokens, y stack), latent [16,{T},60,104], "
                f"{a.steps} DPM++ steps, cfg 7.5 annealed", "n_gpus": world, "ms": ms,
                "denoise_steps_per_s": a.steps / (ms / 1e3), "tflops": 2 * a.steps * (eng.last_flops / 2) / (ms / 1e3) / 1e12,
                "finite": bool(torch.isfinite(out[0]).all())}
    elif a.config == 4 and a.train and a.shim:
        # distilled_trainer.py:241-316 as the trainer sees it: the module's parameters are the source of truth, the
        # engine holds a snapshot that the shim refreshes after every optimizer step
        sd = make_device_weights(cfg, 0, dev)

        class Student(torch.nn.Module):              # attribute surface of WanModel (model.py:445-460)
            def __init__(self):
                super().__init__()
                self.model_type, self.dim, self.ffn_dim, self.num_heads, self.num_layers = "t2v", 1536, 8960, 12, a.layers
                self.in_dim, self.out_dim, self.text_dim, self.text_len, self.freq_dim, self.eps = 16, 16, 4096, 512, 256, 1e-6
                self._keys = list(sd)
                self.ps = torch.nn.ParameterList([torch.nn.Parameter(sd[k].float()) for k in self._keys])

            def named_parameters(self, *aa, **kw):
                return iter(zip(self._keys, self.ps))

            def state_dict(self, *aa, **kw):
                return {k: p.detach() for k, p in zip(self._keys, self.ps)}

            def forward(self, *aa, **kw):
                raise AssertionError("the original forward must not run")

        m = Student()
        eng = b200dit.install(m)
        opt = torch.optim.AdamW(list(m.ps), lr=1e-6)
        gi = torch.Generator().manual_seed(7)
        noises = [torch.randn(16, 1, 60, 104, generator=gi).to(dev) for _ in range(a.items)]
        ctxs = [torch.randn(512, 4096, generator=gi).to(dev) for _ in range(a.items)]
        vts = [torch.randn(16, 1, 60, 104, generator=gi).to(dev) for _ in range(a.items)]
        t1000 = torch.full((a.batch,), 1000.0, device=dev)

        def step(i0):
            idx = list(range(i0, min(i0 + a.batch, a.items)))
            opt.zero_grad(set_to_none=True)
            out = m([noises[i] for i in idx], t=t1000[:len(idx)], context=[ctxs[i] for i in idx], seq_len=1560)
            loss = sum(torch.nn.functional.mse_loss(o, vts[i]) for o, i in zip(out, idx)) / len(idx)
            loss.backward()
            opt.step()
            return loss.detach()
        step(0); step(0)                                       # warm-up: workspaces, AdamW state, transposed weights
        ls, ms = timed(lambda: torch.stack([step(i0) for i0 in range(0, a.items, a.batch)]))
        line = {"config": 4, "workload": f"{a.items} APT stage-1 training items through wan_shim.install on a module with live "
                f"fp32 parameters ({a.batch} item(s) per step): engine forward + backward, param.grad read-back, AdamW step, "
                "engine weight reload before the next forward", "n_gpus": world, "ms": ms,
                "items_per_s": a.items / (ms / 1e3), "ms_per_step": ms / ((a.items + a.batch - 1) // a.batch),
                "reloads": getattr(m, "_b200_reloads", 0), "mean_loss": float(ls.mean()), "finite": bool(torch.isfinite(ls).all())}
    elif a.config == 4 and a.train:
        # the student's training step: items i % world == rank, forward + backward per step of `batch` items;
        # gradients accumulate in the engine (DDP's all-reduce of the 5.7 GB fp32 gradient is the trainer's, not timed)
        eng = b200dit.DitEngine(**cfg, device=dev)
        eng.load_state_dict(make_device_weights(cfg, 0, dev))
        gi = torch.Generator().manual_seed(7)
        noises = [torch.randn(16, 1, 60, 104, generator=gi).to(dev) for _ in range(a.items)]
        ctxs = [torch.randn(512, 4096, generator=gi).to(dev) for _ in range(a.items)]
        vts = [torch.randn(16, 1, 60, 104, generator=gi).to(dev) for _ in range(a.items)]
        mine = list(range(rank, a.items, world))

        def run(idx):
            ls = []
            eng.zero_grad()
            for s0 in range(0, len(idx), a.batch):
                part = idx[s0:s0 + a.batch]
                ls.append(P.student_step(eng, [noises[i] for i in part], [ctxs[i] for i in part], [vts[i] for i in part]))
            return torch.cat(ls)
        run(mine[:a.batch])                                    # warm-up (workspaces, transposed weights)
        with torch.no_grad():
            _, ms_f = timed(lambda: [eng.forward([noises[i]], torch.tensor([1000.0], device=dev), [ctxs[i]], 1560) for i in mine])
        ls, ms = timed(lambda: run(mine))
        _, ms_ar = timed(lambda: parallel.all_reduce_gradients(eng))      # DDP's exchange: once per optimizer step
        gn = float(eng.read_grad("blocks.0.self_attn.q.weight", (1536, 1536)).norm())
        line = {"config": 4, "workload": f"{a.items} APT stage-1 student training steps (forward at t=1000 + MSE + backward, "
                f"{a.batch} item(s) per step) on [16,1,60,104], FFN of blocks > 10 detached as in model.py:318-325, "
                "items i % world == rank", "n_gpus": world, "ms": ms, "items_per_s": a.items / (ms / 1e3),
                "forward_only_ms": ms_f, "fwd_bwd_over_fwd": ms / ms_f, "grad_all_reduce_ms": ms_ar,
                "grad_all_reduce": "two NCCL all_reduce (AVG) calls over the engine's contiguous fp32 gradient stores "
                                   f"({sum(b.numel() for b in eng.grad_buffers()) * 4 / 1e9:.2f} GB), once per optimizer step",
                "mean_loss": float(ls.mean()),
                "grad_norm_blocks0_q": gn, "finite": bool(torch.isfinite(ls).all())}
    elif a.config == 4:
        eng = b200dit.DitEngine(**cfg, device=dev)
        eng.load_state_dict(make_device_weights(cfg, 0, dev))
        gi = torch.Generator().manual_seed(7)                 # identical items on every rank; each rank runs its share
        noises = [torch.randn(16, 1, 60, 104, generator=gi).to(dev) for _ in range(a.items)]
        ctxs = [torch.randn(512, 4096, generator=gi).to(dev) for _ in range(a.items)]
        ctx0 = torch.randn(512, 4096, generator=gi).to(dev)
        sweep = P.teacher_student_pair_split if a.pair_split else P.teacher_student_sweep
        sweep(eng, noises[:2 * world], ctxs[:2 * world], ctx0)     # warm-up
        (vt, vs, ls), ms = timed(lambda: sweep(eng, noises, ctxs, ctx0))
        mode = ("pair-split: rank 2k teacher-cond + student, rank 2k+1 teacher-uncond, one 0.4 MB send per item"
                if a.pair_split else "3 forwards co-batched per item, items i % world == rank")
        line = {"config": 4, "workload": f"{a.items} APT stage-1 items (teacher cond+uncond at t=999, cfg 7.5, student at "
                f"t=1000, MSE) on [16,1,60,104], {mode}, one all_gather",
                "n_gpus": world, "ms": ms, "items_per_s": a.items / (ms / 1e3),
                "forwards_per_s": 3 * a.items / (ms / 1e3), "mean_loss": float(torch.cat(ls).mean()),
                "gathered": [len(vt), len(vs)]}
    else:
        eng = b200dit.DitEngine(**cfg, device=dev)
        eng.load_state_dict(make_device_weights(cfg, 0, dev))
        vae = b200dit.VaeEngine.from_state_dict(synthetic.vae_decoder_weights(dim=96, seed=0), device=dev)
        if a.cfg_parallel:                                    # both ranks of a pair work on the same video
            g = torch.Generator().manual_seed(1000 + rank // 2)
        x0 = rn(16, T, 60, 104).to(dev)
        ctx, ctx0 = rn(512, 4096).bfloat16().to(dev), rn(512, 4096).bfloat16().to(dev)
        run = P.sample_cfg_parallel if a.cfg_parallel else P.sample
        run(eng, [x0], [ctx], [ctx0], steps=3)
        lat, ms_d = timed(lambda: run(eng, [x0], [ctx], [ctx0], steps=a.steps, shift=5.0, guide_scale=5.0))
        if a.cfg_parallel:
            one = P.sample(eng, [x0], [ctx], [ctx0], steps=a.steps, shift=5.0, guide_scale=5.0)
            cfgpar_rel = float((lat[0] - one[0]).norm() / one[0].norm())
        if a.cfg_parallel and a.pipe_decode:                  # the pair that denoised the video also decodes it together
            dec = lambda: [vae.decode_pipelined(lat[0], group=parallel.pair_group())]
        else:
            dec = lambda: vae.decode(lat)
        dec()                                                 # warm-up at full size (workspaces grow on first sight)
        vid, ms_v = timed(dec)
        allv, ms_g = timed(lambda: parallel.gather_items(lat, world))
        line = {"config": 5, "workload": f"per rank: {a.steps}-step UniPC CFG denoise of [16,{T},60,104] + WanVAE decode to "
                f"{list(vid[0].shape)}; one all_gather of the final latents", "n_gpus": world, "denoise_ms": ms_d,
                "vae_decode_ms": ms_v, "gather_ms": ms_g, "videos_per_s": world / ((ms_d + ms_v + ms_g) / 1e3),
                "denoise_steps_per_s": world * a.steps / (ms_d / 1e3), "gathered": len(allv),
                "finite": bool(torch.isfinite(vid[0]).all())}
        if a.cfg_parallel:
            vids = world // 2
            line.update({"mode": "cfg-parallel: one video per rank pair, cond on rank 2k, uncond on rank 2k+1, one all_gather "
                         "inside the pair per step", "videos_per_s": vids / ((ms_d + ms_v + ms_g) / 1e3),
                         "denoise_steps_per_s": vids * a.steps / (ms_d / 1e3),
                         "latency_ms_per_step": ms_d / a.steps, "rel_l2_vs_single_gpu_loop": cfgpar_rel,
                         "vae_decode": "time-chunked across the pair" if a.pipe_decode else "each rank decodes the whole video"})
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
