"""Per-parameter gradient error of the engine's backward against the CPU fp32 oracle's autograd (developer tool).
usage: python tools/grad_check.py [tiny|wide] [detach_from]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200dit  # noqa: E402
from b200dit import autograd as A  # noqa: E402
from oracle import dit_oracle as O  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "tiny"
detach = int(sys.argv[2]) if len(sys.argv) > 2 else None
gen = torch.Generator().manual_seed(3)
if mode == "tiny":
    g = torch.load("tests/golden/dit_t2v_tiny.pt", weights_only=True)
    sd = {k: v.float() for k, v in g["sd"].items() if k != "freqs"}
    heads, xs, ctx, seq_len = 1, [u.float() for u in g["x"]], g["context"], g["seq_len"]
else:
    heads = 12
    sd = O.make_synthetic_weights(1536, 8960, 12, 2, seed=5)
    sd = {k: v for k, v in sd.items() if k != "freqs"}
    xs = [torch.randn(16, 1, 16, 24, generator=gen) for _ in range(2)]
    ctx = [torch.randn(40, 4096, generator=gen), torch.randn(17, 4096, generator=gen)]
    seq_len = 96
vt = [torch.randn(u.shape[0], *u.shape[1:], generator=gen) for u in xs]
t = torch.full((len(xs),), 1000.0)
sd_o = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
x_o = [u.clone().requires_grad_(True) for u in xs]
out_o = O.dit_forward(sd_o, x_o, t, ctx, seq_len, num_heads=heads, ffn_no_grad_from=detach)
loss_o = sum(torch.nn.functional.mse_loss(o, v) for o, v in zip(out_o, vt))
loss_o.backward()
eng = b200dit.DitEngine.from_state_dict(sd, num_heads=heads)
named = [(k, torch.nn.Parameter(v.cuda())) for k, v in sd.items()]
x = [u.cuda().requires_grad_(True) for u in xs]
out = A.dit_forward(eng, named, x, t, ctx, seq_len, ffn_grad_blocks=detach)
loss = sum(torch.nn.functional.mse_loss(o, v.cuda()) for o, v in zip(out, vt))
loss.backward()
torch.cuda.synchronize()
print("loss", float(loss), "oracle", float(loss_o))
for u, r in zip(x, x_o):
    print("dx rel", float((u.grad.cpu() - r.grad).norm() / r.grad.norm()))
for k, p in named:
    r = sd_o[k].grad
    if r is None:
        print(f"{k:44s} oracle None, engine max {0.0 if p.grad is None else float(p.grad.abs().max()):.3e}")
        continue
    gp = p.grad.cpu().reshape(r.shape)
    print(f"{k:44s} |g| {float(r.norm()):.3e}  rel {float((gp - r).norm() / (r.norm() + 1e-30)):.3e}  finite {bool(torch.isfinite(gp).all())}")
