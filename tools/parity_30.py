"""Whole-model parity at BASELINE configs[1] size (30 blocks, [16,1,60,104]) against the CPU fp32 oracle
(developer tool acting as a checker; ~1 min of CPU time).  Prints rel-L2 of one forward and of one CFG step."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import b200dit  # noqa: E402
from oracle import dit_oracle as O  # noqa: E402

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 30
sd = O.make_synthetic_weights(num_layers=layers, seed=0)
eng = b200dit.DitEngine.from_state_dict(sd, num_heads=12)
g = torch.Generator().manual_seed(42)
x = [torch.randn(16, 1, 60, 104, generator=g)]
ctx, ctx0 = [torch.randn(512, 4096, generator=g)], [torch.randn(300, 4096, generator=g)]
t = torch.tensor([999.0])
out = eng.forward(x, t, ctx, 1560)[0].cpu()
cfg = eng.forward_cfg(x, t, ctx, ctx0, 1560, 5.0)[0].cpu()
t0 = time.time()
with torch.no_grad():
    rc = O.dit_forward(sd, x, t, ctx, 1560)[0]
    ru = O.dit_forward(sd, x, t, ctx0, 1560)[0]
rel = lambda a, b: float((a - b).norm() / b.norm())
print(f"{layers} layers: forward rel-L2 {rel(out, rc):.3e}, CFG(5.0) rel-L2 {rel(cfg, O.cfg_combine(rc, ru, 5.0)):.3e}, "
      f"max-abs {float((out - rc).abs().max()):.3e} (output rms {float(rc.pow(2).mean().sqrt()):.3f}); oracle {time.time() - t0:.0f}s",
      flush=True)
