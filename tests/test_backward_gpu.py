"""F1: the backward of the student forward through the C ABI (b200dit_train_forward /
b200dit_backward / b200dit_read_grad) against
 (a) tests/golden/dit_grad_tiny.pt -- gradients of the APT stage-1 loss (distilled_trainer.py:262-289) through the
     UNMODIFIED WanModel (oracle/make_golden.py: make_grad_golden), and
 (b) the CPU fp32 oracle's autograd at the 1.3B width (co-batched items, padded rows, detached FFNs).
Tolerance: rel-L2 <= 2e-3 per gradient tensor (fp16 operands / fp32 accumulation, like the reference's autocast backward);
tensors whose gradient is itself at rounding level (norm << the others) are held to an absolute bar instead."""
import os

import pytest
import torch

from conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu
TOL = 2e-3


def _load(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=True)


def _params(sd, device):
    return [(k, torch.nn.Parameter(v.float().to(device))) for k, v in sd.items() if k != "freqs"]


def test_grad_golden_tiny():
    import b200dit
    from b200dit import autograd as A
    g, r = _load("dit_t2v_tiny.pt"), _load("dit_grad_tiny.pt")
    sd = {k: v.float() for k, v in g["sd"].items()}
    eng = b200dit.DitEngine.from_state_dict(sd, num_heads=g["cfg"]["num_heads"])
    named = _params(sd, eng.device)
    x = [u.float().to(eng.device).requires_grad_(True) for u in g["x"]]
    out = A.dit_forward(eng, named, x, r["t"], g["context"], g["seq_len"], ffn_grad_blocks=11)
    loss = sum(torch.nn.functional.mse_loss(o, v.to(eng.device)) for o, v in zip(out, r["v_teacher"]))
    assert abs(float(loss.detach()) - float(r["loss"])) < 1e-3 * float(r["loss"])
    loss.backward()
    for u, ref in zip(x, r["dx"]):
        assert rel_l2(u.grad.cpu(), ref) < TOL
    params = dict(named)
    worst = {}
    for k, ref in r["grads"].items():
        worst[k] = rel_l2(params[k].grad.cpu().reshape(ref.shape), ref)
    print("grad rel-L2 vs the unmodified reference:", {k: f"{v:.2e}" for k, v in worst.items()})
    assert max(worst.values()) < TOL, worst
    # every parameter's gradient norm (the fixture keeps all of them)
    for k, n in r["grad_norms"].items():
        got = float(params[k].grad.norm())
        assert abs(got - n) <= 5e-3 * n + 1e-7, (k, got, n)


def test_grad_golden_deep_detached_ffns():
    """13 blocks from the unmodified reference: blocks 11 and 12 run their FFN under no_grad (model.py:318-325), so
    their FFN weights get NO gradient and nothing flows back through them -- `ffn_grad_blocks=11`, the shim's default."""
    import b200dit
    from b200dit import autograd as A
    from oracle import dit_oracle as O
    r = _load("dit_grad_deep.pt")
    c = r["cfg"]
    sd = O.make_synthetic_weights(c["dim"], c["ffn_dim"], c["num_heads"], c["num_layers"], in_dim=c["in_dim"],
                                  text_dim=c["text_dim"], seed=c["seed"])
    eng = b200dit.DitEngine.from_state_dict(sd, num_heads=1)
    named = _params(sd, eng.device)
    x = [u.to(eng.device).requires_grad_(True) for u in r["x"]]
    out = A.dit_forward(eng, named, x, r["t"], r["context"], r["seq_len"], ffn_grad_blocks=11)
    assert rel_l2(out[0].detach().cpu(), r["out"][0]) < 1e-3
    loss = sum(torch.nn.functional.mse_loss(o, v.to(eng.device)) for o, v in zip(out, r["v_teacher"]))
    loss.backward()
    assert rel_l2(x[0].grad.cpu(), r["dx"][0]) < TOL
    params = dict(named)
    for k, ref in r["grads"].items():
        assert rel_l2(params[k].grad.cpu().reshape(ref.shape), ref) < TOL, k
    for k in r["no_grad"]:
        assert float(params[k].grad.abs().max()) == 0.0, k          # the engine hands back exact zeros
    for k, n in r["grad_norms"].items():
        assert abs(float(params[k].grad.norm()) - n) <= 5e-3 * n + 1e-7, (k, float(params[k].grad.norm()), n)


def _oracle_grads(sd, xs, t, ctx, seq_len, v_teacher, heads, detach_from=None):
    from oracle import dit_oracle as O
    sd_o = {k: v.clone().float().requires_grad_(True) for k, v in sd.items() if k != "freqs"}
    x_o = [u.clone().requires_grad_(True) for u in xs]
    out = O.dit_forward(sd_o, x_o, t, ctx, seq_len, num_heads=heads, ffn_no_grad_from=detach_from)
    loss = sum(torch.nn.functional.mse_loss(o, v) for o, v in zip(out, v_teacher))
    loss.backward()
    return loss.detach(), [u.grad for u in x_o], {k: v.grad for k, v in sd_o.items()}


@pytest.mark.parametrize("detach_from", [None, 1])
def test_grad_vs_oracle_1p3b_width(detach_from):
    """Two co-batched items at dim 1536 / 12 heads / ffn 8960, 2 layers, latent [16,1,16,24] (L = 96), contexts of
    different lengths; detach_from = 1: block 1's FFN is a constant of the backward (model.py:318-325 does that
    for block_idx > 10)."""
    import b200dit
    from b200dit import autograd as A
    from oracle import dit_oracle as O
    sd = O.make_synthetic_weights(1536, 8960, 12, 2, seed=5)
    gen = torch.Generator().manual_seed(11)
    xs = [torch.randn(16, 1, 16, 24, generator=gen) for _ in range(2)]
    ctx = [torch.randn(40, 4096, generator=gen), torch.randn(17, 4096, generator=gen)]
    vt = [torch.randn(16, 1, 16, 24, generator=gen) for _ in range(2)]
    t = torch.tensor([1000.0, 1000.0])
    loss_o, dx_o, g_o = _oracle_grads(sd, xs, t, ctx, 96, vt, 12, detach_from)
    eng = b200dit.DitEngine.from_state_dict(sd, num_heads=12)
    named = _params(sd, eng.device)
    x = [u.to(eng.device).requires_grad_(True) for u in xs]
    out = A.dit_forward(eng, named, x, t, ctx, 96, ffn_grad_blocks=detach_from)
    loss = sum(torch.nn.functional.mse_loss(o, v.to(eng.device)) for o, v in zip(out, vt))
    loss.backward()
    assert abs(float(loss.detach()) - float(loss_o)) < 2e-3 * float(loss_o)
    for u, ref in zip(x, dx_o):
        assert rel_l2(u.grad.cpu(), ref) < TOL
    errs = {}
    gmax = max(float(v.norm()) for v in g_o.values() if v is not None)
    for k, p in named:
        ref = g_o[k]
        if ref is None or float(ref.norm()) == 0.0:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        errs[k] = rel_l2(p.grad.cpu().reshape(ref.shape), ref)
        # gradients that are themselves at rounding level are held to an absolute bar
        if float(ref.norm()) < 1e-4 * gmax:
            errs[k] = float((p.grad.cpu().reshape(ref.shape) - ref).norm()) / (1e-4 * gmax)
    # softmax is invariant to a shift of all keys, so d loss / d (key bias) is a near-cancelling sum (it only
    # survives through the RMSNorm that follows the bias): |g| ~ 1e-3 of the key weight's gradient, and the fp16
    # rounding of the rows being summed shows up as a few 1e-3 of that remainder
    kb = {k: v for k, v in errs.items() if k.endswith("attn.k.bias")}
    errs = {k: v for k, v in errs.items() if k not in kb}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    print("worst gradient rel-L2 vs the fp32 oracle:", [(k, f"{v:.2e}") for k, v in worst],
          "key biases:", {k: f"{v:.2e}" for k, v in kb.items()})
    assert worst[0][1] < TOL, worst
    assert max(kb.values()) < 1e-2, kb


def test_grad_padded_rows_and_long_sequence():
    """(a) `pad_to_seq_len`: the seq_len - L zero rows of every item are carried through the blocks (model.py:522) --
    queries but never keys; they must not change any gradient.  (b) 2 co-batched items of L = 2048 tokens (16 key
    tiles x 32 query steps per head in the fused attention backward), 2 heads, against the oracle's autograd."""
    import b200dit
    from b200dit import autograd as A
    from oracle import dit_oracle as O
    sd = O.make_synthetic_weights(256, 512, 2, 1, text_dim=64, seed=9)
    gen = torch.Generator().manual_seed(4)
    for shape, seq_len, pad in (((16, 1, 8, 24), 64, True), ((16, 8, 32, 32), 2048, False)):
        xs = [torch.randn(*shape, generator=gen) for _ in range(2)]
        ctx = [torch.randn(33, 64, generator=gen), torch.randn(64, 64, generator=gen)]
        vt = [torch.randn(*shape, generator=gen) for _ in range(2)]
        t = torch.tensor([1000.0, 1000.0])
        loss_o, dx_o, g_o = _oracle_grads(sd, xs, t, ctx, seq_len, vt, 2)
        eng = b200dit.DitEngine.from_state_dict(sd, num_heads=2)
        eng.set_pad_to_seq_len(pad)
        named = _params(sd, eng.device)
        x = [u.to(eng.device).requires_grad_(True) for u in xs]
        out = A.dit_forward(eng, named, x, t, ctx, seq_len, ffn_grad_blocks=None)
        loss = sum(torch.nn.functional.mse_loss(o, v.to(eng.device)) for o, v in zip(out, vt))
        loss.backward()
        for u, ref in zip(x, dx_o):
            assert rel_l2(u.grad.cpu(), ref) < TOL, (shape, rel_l2(u.grad.cpu(), ref))
        errs = {k: rel_l2(p.grad.cpu().reshape(g_o[k].shape), g_o[k]) for k, p in named
                if not k.endswith("attn.k.bias")}
        worst = max(errs.items(), key=lambda kv: kv[1])
        print(shape, "worst gradient rel-L2 vs the oracle:", worst)
        assert worst[1] < TOL, worst


def test_backward_needs_its_forward_and_accumulates():
    """b200dit_backward refuses to run after another forward; gradients accumulate until zero_grad."""
    import b200dit
    g, r = _load("dit_t2v_tiny.pt"), _load("dit_grad_tiny.pt")
    sd = {k: v.float() for k, v in g["sd"].items()}
    eng = b200dit.DitEngine.from_state_dict(sd, num_heads=1)
    x, c = [g["x"][0].float()], [g["context"][0]]
    out = eng.train_forward(x, r["t"][:1], c, g["seq_len"])
    d = [torch.ones_like(out[0]) * 1e-3]
    eng.zero_grad()
    eng.backward(d, want_dx=False)
    a = eng.read_grad("blocks.0.ffn.0.weight", (256, 128)).clone()
    eng.backward(d, want_dx=False)                      # same forward, second backward: accumulates
    b = eng.read_grad("blocks.0.ffn.0.weight", (256, 128))
    assert rel_l2(b.cpu(), 2 * a.cpu()) < 1e-6
    # the two contiguous gradient stores a data-parallel trainer all-reduces hold exactly the per-parameter gradients
    bufs = eng.grad_buffers()
    total = sum(float(u.double().pow(2).sum()) for u in bufs)
    per_param = sum(float(eng.read_grad(k, v.shape).double().pow(2).sum()) for k, v in sd.items() if k != "freqs")
    assert abs(total - per_param) <= 1e-9 * per_param and total > 0
    bufs[0].zero_(); bufs[1].zero_()                    # zero-copy views: clearing them clears the engine's gradients
    assert float(eng.read_grad("blocks.0.ffn.0.weight", (256, 128)).abs().max()) == 0.0
    eng.forward(x, r["t"][:1], c, g["seq_len"])
    with pytest.raises(b200dit.B200Error):
        eng.backward(d, want_dx=False)
    with pytest.raises(b200dit.B200Error):
        eng.read_grad("no.such.weight", (1,))


def test_shim_trains_through_the_engine():
    """`wan_shim.install` under autograd: model(x, t, context, seq_len) -> loss.backward() fills param.grad from
    the engine, and an optimizer step is picked up by the next call (weights signature)."""
    import b200dit
    g, r = _load("dit_t2v_tiny.pt"), _load("dit_grad_tiny.pt")
    sd = {k: v.float() for k, v in g["sd"].items() if k != "freqs"}

    class Student(torch.nn.Module):                     # attribute surface of WanModel (model.py:445-460)
        def __init__(self):
            super().__init__()
            self.model_type, self.dim, self.ffn_dim, self.num_heads, self.num_layers = "t2v", 128, 256, 1, 2
            self.in_dim, self.out_dim, self.text_dim, self.text_len, self.freq_dim, self.eps = 16, 16, 32, 512, 256, 1e-6
            self._keys = list(sd)
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(sd[k].cuda()) for k in self._keys])

        def named_parameters(self, *a, **kw):
            return iter(zip(self._keys, self.ps))

        def state_dict(self, *a, **kw):
            return {k: p.detach() for k, p in zip(self._keys, self.ps)}

        def forward(self, *a, **kw):
            raise AssertionError("the original forward must not run")

    m = Student()
    eng = b200dit.install(m)
    x = [u.float().cuda() for u in g["x"]]
    out = m(x, t=r["t"], context=g["context"], seq_len=g["seq_len"])
    loss = sum(torch.nn.functional.mse_loss(o, v.cuda()) for o, v in zip(out, r["v_teacher"]))
    loss.backward()
    grads = dict(zip(m._keys, [p.grad for p in m.ps]))
    for k, ref in r["grads"].items():
        assert rel_l2(grads[k].cpu().reshape(ref.shape), ref) < TOL, k
    opt = torch.optim.SGD(list(m.ps), lr=0.5)
    opt.step()
    with torch.no_grad():
        out2 = m(x, t=r["t"], context=g["context"], seq_len=g["seq_len"])
    assert getattr(m, "_b200_reloads", 0) == 1
    loss2 = sum(torch.nn.functional.mse_loss(o, v.cuda()) for o, v in zip(out2, r["v_teacher"]))
    assert float(loss2) < float(loss)                   # one SGD step on the engine's gradients lowers the loss
    del eng
