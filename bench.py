#!/usr/bin/env python
"""Headline benchmark: DiT denoise-steps/sec, Wan2.1-T2V-1.3B, 480x832 latent [16,1,60,104]
(BASELINE.json configs[1]): one step = cond forward + uncond forward + CFG combine + UniPC solver step
(the reference's default sampler, text2video.py:204-252: 50 steps, shift 5.0, guide 5.0) per sample; every rank (GPU) denoises its own independent samples (weak scaling; default 2 per GPU, run as
four co-batched items -- samples never interact, model.py:515-527 -- so every GEMM sees 6240 token rows; the
single-sample rate is reported beside it as `single_sample`), one NCCL all_gather of the final latents closes
the timed region.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--samples-per-gpu S] [--frames T]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Prints ONE JSON line (rank 0).  Keys: see the task contract; `value` is device-resident throughput,
`e2e` the same metric through the public Python API with pinned-host inputs copied in and the
result copied out every step, `roofline` the tensor-core GEMM family timed with CUDA events inside
real (eager) steps, `cpu_baseline` the CPU oracle port on a bounded sample.
Weights are random-init (reference init, model.py:590-612) because no checkpoint is reachable
offline; data is synthetic of the reference's shapes.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG_13B = dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=30, in_dim=16, out_dim=16, text_dim=4096,
               text_len=512, freq_dim=256)
GUIDE = 5.0          # text2video.py:118
SHIFT = 5.0          # text2video.py:115
NUM_STEPS = 50


def workload_config(T, S, layers, world, graphs=True):
    L = 1560 * T
    return {"workload": f"Wan2.1-T2V-1.3B ({layers} blocks, dim 1536, ffn 8960, 12 heads) CFG denoise step on "
                        f"latent [16,{T},60,104] (L={L} tokens), guide {GUIDE}, UniPC solver step (shift {SHIFT}, "
                        f"{NUM_STEPS}-step schedule), {S} sample(s) per GPU, cond+uncond co-batched",
            "samples_per_gpu": S, "parallelism": f"replicas x{world} (independent samples, one all_gather)",
            "l2_policy": "weights 2.84 GB per forward exceed the 126 MB L2 (no flush needed)",
            "cuda_graphs": graphs, "operands": "fp16 x fp16 -> fp32 accumulate (model.py:540)",
            "context": "text embedding + cross-attention K/V computed once per trajectory (step-invariant), "
                       "not counted in per-step FLOPs"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi samples during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        busy = [v for v in sm if mx and v > 0.3 * mx] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


def make_device_weights(cfg, seed, device, i2v=False):
    """Random-init weights of the reference architecture (no checkpoint is reachable offline)."""
    from b200dit import synthetic
    return synthetic.dit_weights(cfg, seed, device, i2v=i2v)


def _host_threads():
    """All host cores, set explicitly: torchrun exports OMP_NUM_THREADS=1, which made round 1's reference arm
    single-threaded under N > 1."""
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0)) or n
    except Exception:
        pass
    torch.set_num_threads(n)
    return n


class CpuStep:
    """One whole CFG denoise step of the workload on the CPU oracle port (oracle/dit_oracle.py, fp32): cond forward
    + uncond forward through all `layers` blocks + the CFG combine -- the reference's algorithm on the host cores."""

    def __init__(self, frames, layers):
        import torch
        from oracle import dit_oracle as O
        self.O, self.torch = O, torch
        self.cores = _host_threads()
        g = torch.Generator().manual_seed(42)
        self.x = [torch.randn(16, frames, 60, 104, generator=g)]
        self.ctx = [torch.randn(512, 4096, generator=g)]
        self.ctx0 = [torch.randn(512, 4096, generator=g)]
        self.t = torch.tensor([999.0])
        self.L = 1560 * frames
        self.layers = layers
        self.sd = O.make_synthetic_weights(num_layers=layers, seed=0)

    def run(self, layers=None):
        O, nl = self.O, layers or self.layers
        t0 = time.perf_counter()
        with self.torch.no_grad():
            c = O.dit_forward(self.sd, self.x, self.t, self.ctx, self.L, num_layers=nl)[0]
            u = O.dit_forward(self.sd, self.x, self.t, self.ctx0, self.L, num_layers=nl)[0]
            v = O.cfg_combine(c, u, GUIDE)
        if nl == self.layers:
            self.last = (c, u, v)                            # the checker's outputs of the whole step (parity leg)
        return time.perf_counter() - t0

    def parity(self, dev):
        """The engine on exactly the sample the CPU leg just computed (same weights, inputs, t, guide scale):
        rel-L2 of one forward and of the CFG-combined step against the fp32 oracle.  The checker, not the product."""
        import b200dit
        torch = self.torch
        eng = b200dit.DitEngine.from_state_dict(self.sd, num_heads=CFG_13B["num_heads"], device=dev)
        with torch.no_grad():
            c = eng.forward(self.x, self.t, self.ctx, self.L)[0].cpu()
            v = eng.forward_cfg(self.x, self.t, self.ctx, self.ctx0, self.L, GUIDE)[0].cpu()
        rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
        out = {"forward_rel_l2": rel(c, self.last[0]), "cfg_step_rel_l2": rel(v, self.last[2]), "guide_scale": GUIDE,
               "layers": self.layers, "L": self.L, "tolerance": "1e-3 per forward (BASELINE.json north_star); the CFG step "
               "multiplies the cond - uncond difference by the guide scale",
               "checker": "oracle/dit_oracle.py (fp32 CPU) on the cpu_baseline sample: same weights, inputs, t"}
        # yardstick (SURVEY 7 hard part 1): how far the reference's own GPU path (cuBLAS fp16 autocast + flash_attn +
        # eager elementwise under the caller's bf16 autocast) lands from the engine on a CFG step of this workload --
        # measured on a B200 by tools/library_baseline.py, committed under profiles/
        yp = os.path.join(ROOT, "profiles", "r2_library_baseline.jsonl")
        if os.path.isfile(yp):
            for ln in open(yp):
                rec = json.loads(ln)
                if "rel_l2_vs_engine" in rec and rec.get("pattern", "").startswith("co-batched"):
                    out["library_gpu_path_cfg_step_rel_l2_vs_engine"] = rec["rel_l2_vs_engine"]
                    out["library_gpu_path_source"] = "profiles/r2_library_baseline.jsonl (tools/library_baseline.py, same guide scale)"
        eng.close()
        return out


def cpu_baseline(frames, layers=None, full_steps=1):
    """cpu_baseline leg of the GPU arm: the CPU oracle port on all host threads on a BOUNDED sample of the same
    workload.  T = 1: `full_steps` whole 30-block CFG steps (about 6 s each on 16 cores) after a 2-block warm-up --
    nothing is extrapolated.  T > 1 (a whole step is minutes of CPU): 2 blocks timed, block cost scaled to 30."""
    layers = layers or CFG_13B["num_layers"]
    L = 1560 * frames
    if frames == 1:
        st = CpuStep(frames, layers)
        st.run(2)                                            # warm-up: thread pool, allocator
        ts = sorted(st.run() for _ in range(full_steps))
        step_s = ts[len(ts) // 2]
        sample = (f"{full_steps} whole CFG step(s): 2 forwards x {layers} blocks + embeddings/head + combine at L={L}, "
                  f"fp32 CPU oracle (oracle/dit_oracle.py), after a 2-block warm-up; no extrapolation")
    else:
        st = CpuStep(frames, 2)
        st.run(1)
        t1, t2 = st.run(1), st.run(2)
        step_s = t1 + max(t2 - t1, 1e-9) * (layers - 1)
        sample = (f"2 of {layers} blocks x 2 CFG branches + embeddings/head at L={L}, fp32 CPU oracle, block cost "
                  f"scaled to {layers} layers (a whole step at T={frames} is minutes of CPU work)")
    info = {"value": 1.0 / step_s, "unit": "denoise-steps/s", "cores": st.cores, "kind": "port", "sample": sample}
    return info, (st if frames == 1 else None)


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation (oracle port; the Python reference itself
    cannot travel to the GPU box) on all host threads.  Times WHOLE steps (no block extrapolation) and prints the
    number of steps and warm-ups it really ran: `--steps K --warmup W` are honoured up to a wall-clock budget of
    about 150 s (a 30-block CFG step is ~6 s on 16 cores); under torchrun rank 0 alone works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget_s = float(os.environ.get("B200_REF_BUDGET_S", "150"))
    st = CpuStep(args.frames, args.layers)
    st.run(2)                                                # thread pool / allocator warm-up (2 blocks)
    t_first = st.run()                                       # warm-up step 1 (whole step)
    warm = 1
    while warm < args.warmup and (warm + 2) * t_first < 0.2 * budget_s:
        st.run()
        warm += 1
    k = int(max(1, min(args.steps, (budget_s - warm * t_first) // max(t_first, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(k):
        st.run()
    step_s = (time.perf_counter() - t0) / k
    v = 1.0 / step_s
    info = {"value": v, "unit": "denoise-steps/s", "cores": st.cores, "kind": "port",
            "sample": f"{k} whole CFG steps (2 forwards x {args.layers} blocks + embeddings/head + combine, L={st.L}), "
                      f"{warm} whole-step warm-up(s); requested --steps {args.steps} --warmup {args.warmup}, bounded "
                      f"by a {budget_s:.0f} s budget; no extrapolation; the CPU runs one sample at a time (the metric counts "
                      f"sample-steps, so co-batching does not change its meaning)"}
    line = {"impl": "reference", "metric": "DiT denoise-steps/sec (Wan2.1-T2V-1.3B, CFG, 480x832)", "value": v,
            "unit": "denoise-steps/s", "n_gpus": args.gpus, "steps": k, "warmup": warm,
            "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.frames, args.samples_per_gpu, args.layers, args.gpus),
            "cpu_baseline": info,
            "e2e": {"value": v, "unit": "denoise-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "host_threads": st.cores}
    print(json.dumps(line), flush=True)


BLOCK_SHAPES = [("1 x L=6240 (one sequence, latent [16,4,60,104])", 1, 4), ("4 x L=1560 (bench step: 2 samples x cond/uncond)", 4, 1),
                ("1 x L=1560", 1, 1), ("1 x L=32760 (T=21)", 1, 21)]


def vae_leg(dev, burst):
    """SURVEY 8(a) A14-A15 / 8(d) config 5's decode leg: WanVAE decode of one 81-frame latent [16,21,60,104] ->
    [3,81,480,832] on this GPU (synthetic weights of the reference's widths), device-timed; TFLOP/s of the
    reference's convolution FLOPs (no padding counted)."""
    import torch
    import b200dit
    eng = b200dit.VaeEngine.from_state_dict(b200dit.synthetic.vae_decoder_weights(dim=96, seed=0), device=dev)
    z = torch.randn(16, 21, 60, 104, generator=torch.Generator().manual_seed(21)).to(dev)
    for _ in range(2):
        out = eng.decode([z])[0]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for _ in range(n):
        out = eng.decode([z])[0]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = b200dit.flops.vae_decode_flops(21)
    return {"workload": "WanVAE decode [16,21,60,104] -> [3,81,480,832], one GPU, weights and latent resident in HBM",
            "ms": ms, "frames_per_s": 81 / (ms / 1e3), "tflops": fl / (ms / 1e3) / 1e12,
            "frac_of_burst": fl / (ms / 1e3) / 1e12 / burst, "finite": bool(torch.isfinite(out).all())}


def block_table(dev, burst, sustained, shapes=BLOCK_SHAPES):
    """SURVEY 8(d) fused-block micro-benchmark: ONE WanAttentionBlock (LayerNorm+modulation -> QKV -> RMSNorm/RoPE ->
    self-attention -> o -> norm3 -> cross-attention -> FFN, context K/V cached) at the four shapes.  Block time =
    (forward of a 6-block engine - forward of a 2-block engine) / 4 under CUDA graphs, so embeddings, head and
    launch overheads cancel; device-resident inputs, CUDA events, 3 warm-ups."""
    import torch
    import b200dit
    engs = {}
    for n in (2, 6):
        cfg = dict(CFG_13B, num_layers=n)
        engs[n] = b200dit.DitEngine(**cfg, device=dev)
        engs[n].load_state_dict(make_device_weights(cfg, 1, dev))
    g = torch.Generator().manual_seed(9)
    rows = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, B, T in shapes:
        L = 1560 * T
        x = [torch.randn(16, T, 60, 104, generator=g).to(dev) for _ in range(B)]
        ctx = [torch.randn(512, 4096, generator=g).bfloat16().to(dev) for _ in range(B)]
        t = torch.full((B,), 500.0, device=dev)
        reps = 3 if T > 8 else 12
        ms = {}
        for n, eng in engs.items():
            for _ in range(3):
                eng.forward(x, t, ctx, L)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                eng.forward(x, t, ctx, L)
            e1.record()
            torch.cuda.synchronize()
            ms[n] = e0.elapsed_time(e1) / reps
        blk_ms = (ms[6] - ms[2]) / 4.0
        fl = B * b200dit.flops.dit_block_flops(L)
        tf = fl / (blk_ms / 1e3) / 1e12
        rows.append({"shape": name, "items": B, "L": L, "block_ms": blk_ms, "block_gflop": fl / 1e9, "tflops": tf,
                     "frac_of_burst": tf / burst, "frac_of_sustained": tf / sustained})
    for eng in engs.values():
        eng.close()
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--samples-per-gpu", type=int, default=2,
                    help="independent samples denoised together on each GPU (default 2: cond+uncond of two samples = "
                         "four co-batched items, 6240 token rows per GEMM -- the north-star's block shape)")
    ap.add_argument("--frames", type=int, default=1, help="latent frames T (1 = configs[1]; 21 = 81-frame video)")
    ap.add_argument("--layers", type=int, default=CFG_13B["num_layers"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-block-table", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="N > 1: skip the T = 21 and CFG-parallel legs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import b200dit

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = args.steps
    S = args.samples_per_gpu
    T = args.frames
    cfg = dict(CFG_13B, num_layers=args.layers)
    L = 1560 * T

    eng = b200dit.DitEngine(**cfg, device=dev)
    eng.load_state_dict(make_device_weights(cfg, 0, dev))
    eng.set_graphs(not args.no_graphs)
    torch.cuda.synchronize()

    # ---- synthetic inputs: noise seed 42+rank, N(0,1) contexts as the T5 stage emits (bf16)
    g = torch.Generator().manual_seed(42 + rank)
    host_x = [torch.randn(16, T, 60, 104, generator=g).pin_memory() for _ in range(S)]
    host_ctx = [torch.randn(512, 4096, generator=g).bfloat16().pin_memory() for _ in range(S)]
    host_ctx0 = [torch.randn(512, 4096, generator=g).bfloat16().pin_memory() for _ in range(S)]
    lat = [h.to(dev) for h in host_x]
    ctx = [h.to(dev) for h in host_ctx]
    ctx0 = [h.to(dev) for h in host_ctx0]

    # the reference's sampler (text2video.py:204-211): FlowUniPC, 50 steps, shift 5.0; one scheduler per sample
    def new_schedulers():
        out = []
        for _ in range(S):
            sc = b200dit.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
            sc.set_timesteps(NUM_STEPS, device=dev, shift=SHIFT)
            out.append(sc)
        return out

    ts_host = new_schedulers()[0].timesteps.cpu().tolist()          # int64 timesteps, as the reference feeds them
    t_dev = [torch.full((S,), float(t), device=dev) for t in ts_host]
    t_pin = [torch.full((S,), float(t)).pin_memory() for t in ts_host]
    state = {"sched": None}

    def step_device(i, x):
        """one denoise step with everything resident in HBM: fused cond/uncond/CFG forward + UniPC update"""
        k = i % NUM_STEPS
        if k == 0 or state["sched"] is None:
            state["sched"] = new_schedulers()                       # a new trajectory starts
        v = eng.forward_cfg(x, t_dev[k], ctx, ctx0, L, GUIDE)
        return [sc.step(vi.unsqueeze(0), ts_host[k], xi.unsqueeze(0), return_dict=False)[0].squeeze(0)
                for sc, xi, vi in zip(state["sched"], x, v)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    x = lat
    for i in range(W):
        x = step_device(i, x)
    state["sched"] = None
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = b200dit.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    x = lat
    for i in range(K):
        x = step_device(i, x)
    if world > 1:                                         # the single collective of the path: gather final latents
        mine = torch.stack(x)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = b200dit.kernel_launches() - n0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    ms_step = ms / K
    value = world * S / (ms_step / 1e3)

    # ---- e2e: public API with HOST buffers.  Every step copies that step's inputs (latent, t) from pinned
    # host memory and the updated latent back; the prompt contexts are per-TRAJECTORY inputs (the reference
    # encodes them once before its loop, text2video.py:172-182) and are copied in at the first step of each
    # trajectory of NUM_STEPS steps, inside the timed region.
    e2e_state = {"ctx": None, "h2d": 0, "sched": None}
    # results land in pinned host buffers (two sets, alternating: a step reads its inputs from one set and writes
    # its results to the other), so both copies of every step are true asynchronous DMA
    out_pin = [[torch.empty(16, T, 60, 104).pin_memory() for _ in range(S)] for _ in range(2)]

    def step_e2e(i, hx):
        k = i % NUM_STEPS
        if k == 0 or e2e_state["ctx"] is None:
            e2e_state["ctx"] = ([h.to(dev, non_blocking=True) for h in host_ctx],
                                [h.to(dev, non_blocking=True) for h in host_ctx0])
            e2e_state["h2d"] += S * 2 * 512 * 4096 * 2
            e2e_state["sched"] = new_schedulers()
        cs, c0 = e2e_state["ctx"]
        xs = [h.to(dev, non_blocking=True) for h in hx]
        tt = t_pin[k].to(dev, non_blocking=True)
        e2e_state["h2d"] += S * (16 * T * 60 * 104 * 4 + 4)
        v = eng.forward_cfg(xs, tt, cs, c0, L, GUIDE)
        out = out_pin[i & 1]
        for sc, xi, vi, ho in zip(e2e_state["sched"], xs, v, out):
            ho.copy_(sc.step(vi.unsqueeze(0), ts_host[k], xi.unsqueeze(0), return_dict=False)[0].squeeze(0),
                     non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the step's result is on the host before the next step
        return out

    hx = host_x
    for i in range(2):
        step_e2e(i, hx)
    barrier()
    e2e_state["ctx"], e2e_state["h2d"] = None, 0
    t0 = time.perf_counter()
    for i in range(K):
        hx_out = step_e2e(i, hx)
        hx = hx_out
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_val = world * S / (e2e_s / K)
    h2d = e2e_state["h2d"] // K
    d2h = S * 16 * T * 60 * 104 * 4

    # ---- roofline of the dominant kernel family (tcgen05 GEMM), CUDA events around every launch of real steps
    roof = None
    if rank == 0:
        eng.set_graphs(False)
        b200dit.profile_enable(True)
        xs = lat
        state["sched"] = None
        for i in range(2):
            xs = step_device(i, xs)
        torch.cuda.synchronize()
        prof = b200dit.profile_collect()
        b200dit.profile_enable(False)
        eng.set_graphs(not args.no_graphs)
        burst, sustained, hbm, src = peaks()
        gm = prof["gemm"]
        tot_ms = sum(p["ms"] for p in prof.values())
        ach = gm["flops"] / (gm["ms"] / 1e3) / 1e12 if gm["ms"] > 0 else 0.0
        # DRAM traffic per launch of the same kernel family from the committed `ncu --set full` capture at HEAD
        # (profiles/r2_ncu_traffic.json: the six GEMM launches of one block at M = 6240, cold caches), averaged
        # over the launches of a block; the compulsory bytes beside it (operands + outputs once, fp32 residual
        # read + written by the reduce-add epilogues)
        traffic, tsrc, compulsory = None, None, None
        tp = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
        if os.path.isfile(tp) and S == 2 and T == 1:
            tj = json.load(open(tp))
            per = [v["dram_read_mb"] + v["dram_write_mb"] for k, v in tj.items()
                   if isinstance(v, dict) and not k.startswith("patch")]
            traffic = 1e6 * sum(per) / len(per)
            tsrc = ("profiles/r2_ncu_traffic.json (ncu --set full at round-2 HEAD, dram__bytes_read.sum + "
                    "dram__bytes_write.sum, mean over the six GEMM launches of a block)")
            Mr, d_, f_ = 2 * S * L, cfg["dim"], cfg["ffn_dim"]
            comp = [Mr * d_ * 2 + 3 * d_ * d_ * 2 + Mr * 3 * d_ * 2,            # qkv: A + W + fp16 out
                    3 * (Mr * d_ * 2 + d_ * d_ * 2) + 2 * (2 * Mr * d_ * 4) + Mr * d_ * 2,   # o, cross-o (fp32 RMW), cross-q
                    Mr * d_ * 2 + d_ * f_ * 2 + Mr * f_ * 2,                    # ffn.0
                    Mr * f_ * 2 + d_ * f_ * 2 + 2 * Mr * d_ * 4]                # ffn.2
            compulsory = sum(comp) / 6.0
        roof = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 128xN / 256xN tiles, fused epilogues)",
                "achieved": ach, "peak": burst, "peak_sustained": sustained, "unit": "TFLOP/s", "frac": ach / burst,
                "frac_of_sustained": ach / sustained,
                "peak_source": src + " MEASURED_PEAKS.json: `peak` = cuBLAS bf16 burst (BASELINE.md section 2's primary "
                               "denominator), `peak_sustained` = the seconds-long figure under the power cap",
                "traffic": traffic, "traffic_source": tsrc, "compulsory_bytes_per_launch": compulsory,
                "launches_per_step": gm["launches"] // 2,
                "avg_launch_us": 1e3 * gm["ms"] / max(gm["launches"], 1),
                "flops_per_launch": gm["flops"] / max(gm["launches"], 1),
                "share_of_step": gm["ms"] / tot_ms if tot_ms else None,
                "by_category_ms_per_step": {k: v["ms"] / 2 for k, v in prof.items()},
                "attention_tflops": (prof["attention"]["flops"] / (prof["attention"]["ms"] / 1e3) / 1e12)
                if prof["attention"]["ms"] > 0 else None}

    # ---- multi-GPU legs beyond T = 1 (extra keys): the 81-frame shape of configs 3 / 5 (T = 21, L = 32 760), one
    # sample per GPU, and -- for an even number of ranks -- CFG-parallel sampling (cond / uncond of ONE video on a
    # rank pair, one all_gather per step inside the pair): the latency mode of config 5.
    legs = {}
    if world > 1 and T == 1 and not args.no_extra_legs:
        from b200dit import pipelines as P
        T2, L2 = 21, 1560 * 21
        g2 = torch.Generator().manual_seed(4242 + rank)
        x21 = [torch.randn(16, T2, 60, 104, generator=g2).to(dev)]
        sc21 = b200dit.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
        sc21.set_timesteps(NUM_STEPS, device=dev, shift=SHIFT)

        def step21(i, x):
            v = eng.forward_cfg(x, t_dev[i][:1].contiguous(), ctx[:1], ctx0[:1], L2, GUIDE)
            return [sc21.step(v[0].unsqueeze(0), ts_host[i], x[0].unsqueeze(0), return_dict=False)[0].squeeze(0)]

        xx = x21
        for i in range(3):                                   # eager, capture, replay
            xx = step21(i, xx)
        barrier()
        e0.record()
        n21 = 2
        for i in range(3, 3 + n21):
            xx = step21(i, xx)
        e1.record()
        barrier()
        ms21 = e0.elapsed_time(e1) / n21
        tt = torch.tensor([ms21], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms21 = float(tt.item())
        fl21 = 2 * b200dit.flops.dit_forward_flops(L2, layers=cfg["num_layers"], context_cached=True)
        legs["t21"] = {"workload": "same CFG denoise step on latent [16,21,60,104] (L=32760), 1 sample per GPU",
                       "value": world * 1e3 / ms21, "unit": "denoise-steps/s", "ms_per_step": ms21, "steps": n21,
                       "step_tflops_per_gpu": fl21 / (ms21 / 1e3) / 1e12, "finite": bool(torch.isfinite(xx[0]).all())}
        if world % 2 == 0:
            gp = torch.Generator().manual_seed(777 + rank // 2)      # both ranks of a pair hold the same video
            xp = [torch.randn(16, T2, 60, 104, generator=gp).to(dev)]
            cp = [torch.randn(512, 4096, generator=gp).bfloat16().to(dev)]
            c0p = [torch.randn(512, 4096, generator=gp).bfloat16().to(dev)]
            P.sample_cfg_parallel(eng, xp, cp, c0p, steps=3, shift=SHIFT, guide_scale=GUIDE)
            barrier()
            e0.record()
            nps = 3
            outp = P.sample_cfg_parallel(eng, xp, cp, c0p, steps=nps, shift=SHIFT, guide_scale=GUIDE)
            e1.record()
            barrier()
            msp = e0.elapsed_time(e1) / nps
            tt = torch.tensor([msp], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            legs["cfg_parallel_t21"] = {"workload": "one [16,21,60,104] video per rank PAIR: cond forward on rank 2k, uncond on "
                                        "rank 2k+1, one all_gather (8.4 MB) inside the pair per step, same UniPC step on both",
                                        "latency_ms_per_step": float(tt.item()), "videos_in_flight": world // 2,
                                        "value": (world // 2) * 1e3 / float(tt.item()), "unit": "denoise-steps/s",
                                        "steps": nps, "finite": bool(torch.isfinite(outp[0]).all())}

    single = None
    if rank == 0 and S > 1:
        # the same step for ONE sample per GPU (latency-oriented use: M = 3120 token rows per GEMM)
        x1, c1, c01 = lat[:1], ctx[:1], ctx0[:1]
        t1 = [t[:1].contiguous() for t in t_dev]

        def step_one(i, x, sc):
            v = eng.forward_cfg(x, t1[i], c1, c01, L, GUIDE)
            return [sc.step(v[0].unsqueeze(0), ts_host[i], x[0].unsqueeze(0), return_dict=False)[0].squeeze(0)]

        for rep in range(2):
            sc = new_schedulers()[0]
            xx = x1
            n1 = min(K, NUM_STEPS)
            if rep == 1:
                torch.cuda.synchronize()
                e0.record()
            for i in range(n1 if rep == 1 else W):
                xx = step_one(i, xx, sc)
            if rep == 1:
                e1.record()
                torch.cuda.synchronize()
                ms1 = e0.elapsed_time(e1) / n1
                single = {"value": 1e3 / ms1, "unit": "denoise-steps/s", "ms_per_step": ms1, "samples_per_gpu": 1,
                          "step_tflops": 2 * b200dit.flops.dit_forward_flops(L, layers=cfg["num_layers"],
                                                                               context_cached=True) / (ms1 / 1e3) / 1e12}

    if rank == 0:
        # context work (text embedding, cross k/v projections) runs once per trajectory, not per step
        flops_step = 2 * S * b200dit.flops.dit_forward_flops(L, layers=cfg["num_layers"], context_cached=True)
        burst, sustained, hbm, src = peaks()
        line = {"metric": "DiT denoise-steps/sec (Wan2.1-T2V-1.3B, CFG, 480x832)", "value": value,
                "unit": "denoise-steps/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
                "data": "synthetic",
                "config": workload_config(T, S, cfg["num_layers"], world, not args.no_graphs),
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "denoise-steps/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / K},
                "gpu_launches": int(launches),
                "step_tflops": flops_step / (ms_step / 1e3) / 1e12,
                "step_frac_of_peak": flops_step / (ms_step / 1e3) / 1e12 / burst,
                "step_frac_of_sustained": flops_step / (ms_step / 1e3) / 1e12 / sustained,
                "roofline": roof, "single_sample": single}
        line.update(legs)
        if world == 1 and not args.no_block_table:
            # extra legs: a failure in one of them is reported in its key and never costs the headline line
            for key, leg in (("block_table", lambda: block_table(dev, burst, sustained)),
                             ("vae_decode", lambda: vae_leg(dev, burst))):
                try:
                    line[key] = leg()
                except Exception as exc:  # noqa: BLE001
                    line[key] = {"error": f"{type(exc).__name__}: {exc}"}
                    print(f"bench: leg {key} failed: {exc!r}", file=sys.stderr)
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"], cpu_step = cpu_baseline(T)
            if cpu_step is not None and hasattr(cpu_step, "last"):
                try:
                    line["parity"] = cpu_step.parity(dev)       # live, at the full 30-block size
                except Exception as exc:  # noqa: BLE001
                    line["parity"] = {"error": f"{type(exc).__name__}: {exc}"}
                    print(f"bench: parity leg failed: {exc!r}", file=sys.stderr)
        elif world > 1:
            line["cpu_baseline"] = {"value": None, "unit": "denoise-steps/s", "cores": 0, "kind": "port",
                                    "sample": "measured at N=1 only"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
