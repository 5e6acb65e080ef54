"""Multi-GPU layer: one process per GPU (torchrun), independent items sharded across ranks, ONE
collective -- an all_gather of the finished latents -- at the end.

The path has no exchange step: `WanModel.forward` never mixes batch items (model.py:515-527 take a
per-item `t`; probed co-batching equivalence), the teacher sweep iterates independent
seeds (generate.py:209-232) and cond/uncond share inputs but no state (text2video.py:238-241).  So
items `i % world == rank` run on a full weight replica and nothing crosses NVLink until the gather.
Results are bit-identical to a single-GPU run as long as the per-call co-batch size is the same
(every output row's reduction order is fixed by the tile schedule, not by the world size).

Everything here is host logic over torch.distributed and works on CPU tensors with the gloo
backend (tests/test_cpu_parallel.py); on the B200 box the backend is NCCL over NVLink 5/NVSwitch.
"""
import os

import torch
import torch.distributed as dist


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 or dist.is_initialized():
        return rank(), max(world, world_size())
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        kw["device_id"] = torch.device("cuda", local)
    dist.init_process_group(backend, **kw)
    return dist.get_rank(), dist.get_world_size()


def rank():
    return dist.get_rank() if dist.is_initialized() else 0


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def shard_indices(n_items, rank_=None, world=None):
    """Round-robin ownership: item i belongs to rank i % world."""
    r = rank() if rank_ is None else rank_
    w = world_size() if world is None else world
    return list(range(r, n_items, w))


def gather_items(local, n_items, group=None):
    """The single collective of the path.  `local` holds this rank's results in shard_indices order
    (tensors of identical shape); returns all n_items results in original order on every rank."""
    w = world_size()
    if w == 1:
        return list(local)
    if n_items < w:
        # some ranks own nothing: every rank knows that from (n_items, world) alone, so all of them take the
        # metadata-exchanging path together instead of one rank failing before the collective (deadlock)
        counts = [len(range(r, n_items, w)) for r in range(w)]
        lists = gather_from_ranks(local, counts, group=group)
        res = [None] * n_items
        for r in range(w):
            for j, i in enumerate(range(r, n_items, w)):
                res[i] = lists[r][j]
        return res
    per_rank = (n_items + w - 1) // w
    proto = local[0]
    buf = torch.zeros((per_rank,) + tuple(proto.shape), dtype=proto.dtype, device=proto.device)
    for j, t in enumerate(local):
        buf[j].copy_(t)
    out = [torch.empty_like(buf) for _ in range(w)]
    dist.all_gather(out, buf, group=group)
    res = [None] * n_items
    for r in range(w):
        for j, i in enumerate(range(r, n_items, w)):
            res[i] = out[r][j]
    return res


def sharded_map(fn, items, group=None):
    """Runs fn(item) for the items this rank owns and gathers every result (one all_gather)."""
    idx = shard_indices(len(items))
    local = [fn(items[i]) for i in idx]
    return gather_items(local, len(items), group=group)


_pair_groups = {}


def pair_group():
    """The process group of this rank's pair (2k, 2k+1).  Every rank creates every pair's group (new_group is a
    collective over the whole world); for world == 2 the pair is the default group."""
    w, r = world_size(), rank()
    if w % 2:
        raise ValueError("rank pairs need an even world size")
    if w == 2:
        return None
    if not _pair_groups:
        for k in range(w // 2):
            _pair_groups[k] = dist.new_group(ranks=[2 * k, 2 * k + 1])
    return _pair_groups[r // 2]


def exchange_pair(t):
    """One all_gather inside the rank pair: returns (tensor of rank 2k, tensor of rank 2k+1) on both ranks.  This is
    the per-step exchange of CFG-parallel sampling (pipelines.sample_cfg_parallel): 0.4 MB per 480p latent frame."""
    t = t.contiguous()
    out = [torch.empty_like(t), torch.empty_like(t)]
    dist.all_gather(out, t, group=pair_group())
    return out[0], out[1]


def send_to(t, dst):
    """Point-to-point half of the pair-split exchange (pipelines.teacher_student_pair_split)."""
    dist.send(t.contiguous(), dst)


def recv_from(like, src):
    buf = torch.empty_like(like)
    dist.recv(buf, src)
    return buf


def gather_from_ranks(local, counts, group=None):
    """all_gather of per-rank stacks whose lengths differ: `local` is this rank's list of same-shape tensors
    (possibly empty), counts[r] the length of rank r's list.  Returns the per-rank lists on every rank."""
    w = world_size()
    if w == 1:
        return [list(local)]
    proto_shape, dtype, device = None, None, None
    if local:
        proto_shape, dtype, device = tuple(local[0].shape), local[0].dtype, local[0].device
    meta = [None] * w
    dist.all_gather_object(meta, (proto_shape, str(dtype) if dtype else None), group=group)
    if all(m[0] is None for m in meta):       # seen identically by every rank after the exchange: a clean failure
        raise ValueError("gather_from_ranks: no rank holds an item")
    shape = next(m[0] for m in meta if m[0] is not None)
    if dtype is None:
        dtype = getattr(torch, next(m[1] for m in meta if m[1] is not None).split(".")[-1])
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    n = max(counts)
    buf = torch.zeros((n,) + shape, dtype=dtype, device=device)
    for j, t in enumerate(local):
        buf[j].copy_(t)
    out = [torch.empty_like(buf) for _ in range(w)]
    dist.all_gather(out, buf, group=group)
    return [[out[r][j] for j in range(counts[r])] for r in range(w)]


def all_reduce_gradients(engine, group=None, average=True):
    """Data-parallel training step (the DDP that accelerate wraps around the student, distilled_trainer.py:79):
    every rank ran `backward` on its own items; the gradients -- two contiguous fp32 buffers inside the engine
    (`DitEngine.grad_buffers`) -- are summed over the ranks in place (NCCL all_reduce over NVLink on the B200 box) and
    divided by the world size, so `read_grad` / `param.grad` then hold the mean gradient on every rank."""
    if world_size() == 1:
        return
    w = dist.get_world_size(group)
    nccl = dist.get_backend(group) == "nccl"
    for buf in engine.grad_buffers():
        if average and nccl:
            dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=group)      # averaged inside the collective
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
            if average:
                buf.mul_(1.0 / w)


def pipeline_schedule(T, world, chunk_frames):
    """Chunk schedule of the multi-GPU time-chunked VAE decode (b200vae_decode_pipelined): chunk 0 is latent frame
    0 alone (one output frame), chunk k >= 1 holds up to `chunk_frames` latent frames (four output frames each);
    chunk c runs on rank c % world.  Returns [(chunk, rank, first output frame, output frames)]."""
    out, t0, f0, c = [], 0, 0, 0
    while t0 < T:
        tc = 1 if t0 == 0 else min(chunk_frames, T - t0)
        nf = 1 if t0 == 0 else 4 * tc
        out.append((c, c % world, f0, nf))
        t0, f0, c = t0 + tc, f0 + nf, c + 1
    return out


def pipeline_chunk_frames(T, world, max_chunk=4):
    """Latent frames per chunk for a ring of `world` ranks: the ranks run one conv layer apart, so the decode takes
    about ceil(chunks / world) rounds of one chunk each.  Larger chunks feed the convolution GEMMs better, so a
    smaller chunk must win by more than 10 % to be chosen."""
    best, best_cost = max_chunk, None
    for cf in range(max_chunk, 0, -1):
        chunks = len(pipeline_schedule(T, world, cf))
        cost = ((chunks + world - 1) // world) * min(cf, max(T - 1, 1))
        if best_cost is None or cost < 0.9 * best_cost:
            best, best_cost = cf, cost
    return best


def sum_disjoint(t, group=None):
    """Assembles a tensor of which every rank wrote a disjoint part and left the rest zero: one all_reduce
    (x + 0 = x exactly).  Used for the frames of the pipelined VAE decode."""
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def max_over_ranks(value, device=None):
    """Timing reduction used by bench.py: every multi-GPU number is the max over ranks."""
    if world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
