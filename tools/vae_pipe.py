"""Multi-GPU time-chunked WanVAE decode, launched under torchrun with N >= 2 ranks on one node:
every rank decodes its chunks of ONE latent, per-conv caches cross NVLink through peer memory
(b200vae_decode_pipelined).  Rank 0 also runs the single-GPU decode and reports agreement and both times.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/vae_pipe.py [--frames 21 --h 60 --w 104 --dim 96 --chunk 0 --reps 3]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import b200dit  # noqa: E402
from b200dit import parallel, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=21)
    ap.add_argument("--h", type=int, default=60)
    ap.add_argument("--w", type=int, default=104)
    ap.add_argument("--dim", type=int, default=96)
    ap.add_argument("--chunk", type=int, default=0, help="latent frames per chunk (0 = pick for the ring size)")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    rank, world = parallel.init()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    vae = b200dit.VaeEngine.from_state_dict(synthetic.vae_decoder_weights(dim=a.dim, seed=0), device=dev)
    z = torch.randn(16, a.frames, a.h, a.w, generator=torch.Generator().manual_seed(5)).to(dev)
    cf = a.chunk or parallel.pipeline_chunk_frames(a.frames, world)

    def timed(fn):
        out = fn()                                             # warm-up (workspaces, arena, IPC mapping)
        ms = []
        for _ in range(a.reps):
            torch.cuda.synchronize()
            if world > 1:
                torch.distributed.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(parallel.max_over_ranks(e0.elapsed_time(e1)))
        return out, sorted(ms)[len(ms) // 2]

    piped, ms_p = timed(lambda: vae.decode_pipelined(z, chunk_frames=cf))
    line = None
    if rank == 0:
        single, ms_1 = None, None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        vae.decode([z])
        torch.cuda.synchronize()
        e0.record()
        single = vae.decode([z])[0]
        e1.record()
        torch.cuda.synchronize()
        ms_1 = e0.elapsed_time(e1)
        line = {"what": "WanVAE decode, one latent, time-chunked across the ranks of one node (per-conv caches over NVLink peer memory)",
                "latent": [16, a.frames, a.h, a.w], "video": list(piped.shape), "n_gpus": world, "chunk_frames": cf,
                "chunks": len(parallel.pipeline_schedule(a.frames, world, cf)),
                "ms_pipelined_max_over_ranks": ms_p, "ms_single_gpu": ms_1, "speedup": ms_1 / ms_p,
                "max_abs_vs_single_gpu": float((piped - single).abs().max()), "finite": bool(torch.isfinite(piped).all())}
    if world > 1:
        torch.distributed.barrier()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
