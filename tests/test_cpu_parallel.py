"""N>1 host path on CPU: world_size-2 gloo processes shard items round-robin, run a stand-in for the
per-item computation, and gather with the single all_gather (omnihuman-1-hack_b200/parallel.py)."""
import os
import socket
import subprocess
import sys

from conftest import ROOT

WORKER = r'''
import os, sys, torch
sys.path.insert(0, os.environ["B200_ROOT"])
import importlib.util
spec = importlib.util.spec_from_file_location("b200par", os.path.join(os.environ["B200_ROOT"], "omnihuman-1-hack_b200", "parallel.py"))
par = importlib.util.module_from_spec(spec); spec.loader.exec_module(par)
rank, world = par.init("gloo")
assert world == 2
items = [torch.full((16, 1, 4, 6), float(i)) for i in range(5)]
calls = []
def fn(x):
    calls.append(int(x[0, 0, 0, 0]))
    return x * 2 + 1
res = par.sharded_map(fn, items)
assert calls == list(range(rank, 5, 2)), calls                       # ownership i % world == rank
for i, r in enumerate(res):
    assert torch.equal(r, items[i] * 2 + 1), i                        # original order, every rank
assert par.max_over_ranks(1.0 + rank) == 2.0
assert par.shard_indices(7, 1, 4) == [1, 5]
# pipelined VAE decode, host side: every rank writes the frames of its chunks, one all_reduce assembles the video
sched = par.pipeline_schedule(21, 2, 4)
assert [s[1] for s in sched] == [0, 1, 0, 1, 0, 1] and sched[0][2:] == (0, 1) and sched[1][2:] == (1, 16) and sched[-1][2] + sched[-1][3] == 81
full = torch.arange(81.0)[None, :, None].expand(3, 81, 4).contiguous()
mine = torch.zeros_like(full)
for c, r, f0, nf in sched:
    if r == rank:
        mine[:, f0:f0 + nf] = full[:, f0:f0 + nf]
assert torch.equal(par.sum_disjoint(mine), full)
assert par.pipeline_chunk_frames(21, 8) == 3 and par.pipeline_chunk_frames(21, 2) == 4 and par.pipeline_chunk_frames(2, 2) == 4
# fewer items than ranks: rank 1 owns nothing and still takes part in the gather (no deadlock, no early raise)
one = par.sharded_map(fn, items[:1])
assert len(one) == 1 and torch.equal(one[0], items[0] * 2 + 1)
try:
    par.gather_from_ranks([], [0, 0])
    raise SystemExit("expected ValueError")
except ValueError:
    pass
print("ok", rank)
'''


def test_gloo_world2_shard_and_gather(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), B200_ROOT=ROOT, CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for r, p in enumerate(procs):
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out
        assert f"ok {r}" in out


GRAD_WORKER = r'''
import os, sys, torch
sys.path.insert(0, os.environ["B200_ROOT"])
import b200dit
from b200dit import parallel as par
rank, world = par.init("gloo")


class FakeEngine:                      # DitEngine.grad_buffers: two contiguous fp32 gradient stores
    def __init__(self):
        self.b = [torch.full((1000,), float(rank + 1)), torch.arange(7, dtype=torch.float32) * (rank + 1)]

    def grad_buffers(self):
        return self.b


e = FakeEngine()
par.all_reduce_gradients(e)
assert torch.allclose(e.b[0], torch.full((1000,), 1.5)) and torch.allclose(e.b[1], torch.arange(7, dtype=torch.float32) * 1.5)
print("ok", rank)
'''


def test_gloo_world2_gradient_all_reduce(tmp_path):
    """Data-parallel training step (DDP around the student, distilled_trainer.py:79): the engine's gradient stores are
    averaged over the ranks in place."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "grad_worker.py"
    script.write_text(GRAD_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), B200_ROOT=ROOT, CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for r, p in enumerate(procs):
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out
        assert f"ok {r}" in out


PAIR_WORKER = r'''
import os, sys, torch
sys.path.insert(0, os.environ["B200_ROOT"])
import b200dit
from b200dit import parallel as par, pipelines as P, solvers as PS
rank, world = par.init("gloo")
assert world == 2


def cpu_lincomb(inputs, coeffs, like):          # host-logic stand-in for the fused CUDA combine (tests only)
    ins = [t if t is not None else like for t in inputs]
    return [sum(c * t for c, t in zip(row, ins)) for row in coeffs]


PS._lincomb = cpu_lincomb


class FakeEngine:
    """Stand-in with DitEngine.forward's signature: a deterministic per-item function of (x, t, context)."""
    device = torch.device("cpu")

    def forward(self, xs, t, ctx, seq_len):
        return [torch.tanh(x * c.mean() + tt / 1000.0) + 0.1 * c.std() for x, tt, c in zip(xs, t, ctx)]


g = torch.Generator().manual_seed(3)
noises = [torch.randn(16, 1, 4, 6, generator=g) for _ in range(5)]
ctxs = [torch.randn(7, 8, generator=g) for _ in range(5)]
ctx0 = torch.randn(7, 8, generator=g)
eng = FakeEngine()
vt, vs, ls = P.teacher_student_pair_split(eng, noises, ctxs, ctx0, seq_len=6)
assert len(vt) == len(vs) == len(ls) == 5
for i in range(5):                                   # every rank holds every item, equal to the unsharded item
    a, b, l = P.teacher_student_item(eng, noises[i], ctxs[i], ctx0, seq_len=6)
    assert torch.allclose(vt[i], a, atol=1e-6) and torch.allclose(vs[i], b, atol=1e-6), i
    assert torch.allclose(ls[i], l.reshape(1), atol=1e-6), i
# the default mode gives the same triple
vt2, vs2, ls2 = P.teacher_student_sweep(eng, noises, ctxs, ctx0, seq_len=6)
for i in range(5):
    assert torch.allclose(vt2[i], vt[i], atol=1e-6) and torch.allclose(ls2[i], ls[i], atol=1e-6)
print("ok", rank)
'''


def test_gloo_world2_pair_split_matches_item(tmp_path):
    """Config 4, pair-split mode: rank 0 = teacher-cond + student, rank 1 = teacher-uncond, one send per item."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "pair_worker.py"
    script.write_text(PAIR_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), B200_ROOT=ROOT, CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for r, p in enumerate(procs):
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out
        assert f"ok {r}" in out


CFGPAR_WORKER = r'''
import os, sys, torch
sys.path.insert(0, os.environ["B200_ROOT"])
import b200dit
from b200dit import parallel as par, pipelines as P, solvers as PS
rank, world = par.init("gloo")
assert world == 2


def cpu_lincomb(inputs, coeffs, like):          # host-logic stand-in for the fused CUDA update (tests only)
    ins = [t if t is not None else like for t in inputs]
    return [sum(float(c) * t for c, t in zip(row, ins)) for row in coeffs]


PS._lincomb = cpu_lincomb


class FakeEngine:
    device = torch.device("cpu")
    calls = 0

    def forward(self, xs, t, ctx, seq_len):
        FakeEngine.calls += len(xs)
        return [torch.tanh(0.3 * x + c.mean() + tt / 1000.0) - 0.2 * x for x, tt, c in zip(xs, t, ctx)]

    def forward_cfg(self, xs, t, ctx, ctx_n, seq_len, s, clip_fea=None, y=None):
        if len(ctx_n) == 1 and len(xs) > 1:
            ctx_n = ctx_n * len(xs)
        c, u = self.forward(xs, t, ctx, seq_len), self.forward(xs, t, ctx_n, seq_len)
        return [ui + s * (ci - ui) for ci, ui in zip(c, u)]

    def nonfinite_rows(self):
        return 0


g = torch.Generator().manual_seed(5)
noise = [torch.randn(16, 2, 4, 6, generator=g) for _ in range(2)]
ctx = [torch.randn(9, 8, generator=g) for _ in range(2)]
ctx0 = [torch.randn(4, 8, generator=g)]
eng = FakeEngine()
for solver, anneal in (("unipc", False), ("dpm++", True)):
    FakeEngine.calls = 0
    out = P.sample_cfg_parallel(eng, noise, ctx, ctx0, steps=5, shift=3.0, guide_scale=4.0, solver=solver, cfg_anneal=anneal)
    assert FakeEngine.calls == 5 * 2, FakeEngine.calls       # one forward per sample per step on this rank (not two)
    ref = P.sample(eng, noise, ctx, ctx0, steps=5, shift=3.0, guide_scale=4.0, solver=solver, cfg_anneal=anneal)
    for a, b in zip(out, ref):
        assert torch.allclose(a, b, atol=1e-5), float((a - b).abs().max())
    both = par.exchange_pair(torch.stack(out))               # the two ranks hold identical latents
    assert torch.equal(both[0], both[1])
print("ok", rank)
'''


def test_gloo_world2_cfg_parallel_matches_sample(tmp_path):
    """CFG-parallel sampling: rank 0 conditional, rank 1 unconditional, one exchange per step; same trajectory as
    the single-process loop, identical on both ranks, half the forwards per rank."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "cfgpar_worker.py"
    script.write_text(CFGPAR_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), B200_ROOT=ROOT, CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for r, p in enumerate(procs):
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out
        assert f"ok {r}" in out
