"""`loss.backward()` through the B200 engine: the student step of APT stage 1.

The reference trains the student `WanModel` with `v = model(noise, t, context, seq_len)` under autocast,
`loss = F.mse_loss(v, v_teacher)`, `scaler.scale(loss).backward()` (seaweed_apt/distilled_trainer.py:268-301), its
blocks checkpointed (model.py:544-548).  `dit_forward` is that forward as ONE autograd node per latent grid: the
forward is `DitEngine.train_forward`, the backward `DitEngine.backward` (block-wise recompute + adjoint on the
tcgen05 GEMM), and the node hands the gradients to autograd for the latents and for every parameter of the module,
so optimizers, GradScaler and gradient accumulation upstream keep working on `param.grad` unchanged.
"""
import torch

from . import _lib


class _DitNode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, names, t, context, seq_len, ffn_grad_blocks, n_x, *tensors):
        xs = [u.detach() for u in tensors[:n_x]]
        outs = eng.train_forward(xs, t, context, seq_len)
        ctx.eng, ctx.seq, ctx.names, ctx.n_x = eng, eng._train_seq, names, n_x
        ctx.inputs = (xs, t, context, seq_len)
        ctx.ffn_grad_blocks = ffn_grad_blocks
        ctx.shapes = [tuple(p.shape) for p in tensors[n_x:]]
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        eng = ctx.eng
        if eng._train_seq != ctx.seq:                 # another forward used the engine since: run this one again
            outs = eng.train_forward(*ctx.inputs)
        else:
            outs = None
        xs = ctx.inputs[0]
        ds = []
        for i, d in enumerate(douts):
            ds.append(d if d is not None else torch.zeros((eng.cfg["out_dim"],) + tuple(xs[i].shape[1:]),
                                                          dtype=torch.float32, device=eng.device))
        del outs
        want_dx = any(ctx.needs_input_grad[7:7 + ctx.n_x])
        eng.zero_grad()
        dxs = eng.backward(ds, ffn_grad_blocks=ctx.ffn_grad_blocks, want_dx=want_dx)
        grads = [None] * 7
        for i in range(ctx.n_x):
            grads.append(dxs[i] if want_dx and ctx.needs_input_grad[7 + i] else None)
        for j, (name, shape) in enumerate(zip(ctx.names, ctx.shapes)):
            grads.append(eng.read_grad(name, shape) if ctx.needs_input_grad[7 + ctx.n_x + j] else None)
        return tuple(grads)


def supports(eng, clip_fea=None, y=None):
    """The engine differentiates the t2v student (distilled_trainer.py:262-278: latents, t, text contexts)."""
    return clip_fea is None and y is None and not eng.cfg["i2v"]


def dit_forward(eng, named_params, x, t, context, seq_len, ffn_grad_blocks=11):
    """`WanModel.forward(x, t, context, seq_len)` recorded by autograd.  named_params: [(state_dict key, Parameter)]
    of the module whose weights the engine holds; x: list of [C, F, H, W] latents.  Items are co-batched per latent
    grid exactly as `DitEngine.forward` does.  `ffn_grad_blocks=11` keeps the reference's behaviour of treating the
    FFN of blocks with block_idx > 10 as a constant (model.py:318-325); None differentiates every FFN."""
    xs = list(x)
    n = len(xs)
    tt = eng._t_tensor(t, n)
    names = [k for k, _ in named_params]
    params = [p for _, p in named_params]
    groups = {}
    for i, u in enumerate(xs):
        groups.setdefault(tuple(u.shape[1:]), []).append(i)
    outs = [None] * n
    for (F, H, W), idx in groups.items():
        chunk = _lib.MAX_ITEMS if (F * (H // 2) * (W // 2)) % 8 == 0 else 1
        for s in range(0, len(idx), chunk):
            part = idx[s:s + chunk]
            o = _DitNode.apply(eng, names, tt[part].contiguous(), [context[i] for i in part], seq_len, ffn_grad_blocks,
                               len(part), *[xs[i] for i in part], *params)
            for i, q in zip(part, o):
                outs[i] = q
    return outs
