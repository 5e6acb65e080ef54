"""ncu target (developer tool): ONE eager pass of a hot path between cudaProfilerStart/Stop, after a warm-up pass.
  dit   two-sample CFG forward of a 2-block 1.3B-width model (M = 6240: every per-block kernel of the bench step)
  vae [T]  WanVAE decode of a [16,T,60,104] latent (default T = 2: frame 0 alone + one full-resolution chunk)
  bwd   one student training step (forward + backward), one item, 2 blocks
usage: ncu --profile-from-start off ... python tools/ncu_step_target.py dit"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200dit  # noqa: E402
from b200dit import pipelines as P, synthetic  # noqa: E402
from bench import CFG_13B, make_device_weights  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "dit"
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(7)
rn = lambda *s: torch.randn(*s, generator=g).to(dev)
if what == "vae":
    vae = b200dit.VaeEngine.from_state_dict(synthetic.vae_decoder_weights(dim=96, seed=0), device=dev)
    z = [rn(16, int(sys.argv[2]) if len(sys.argv) > 2 else 2, 60, 104)]
    run = lambda: vae.decode(z)
else:
    cfg = dict(CFG_13B, num_layers=2)
    eng = b200dit.DitEngine(**cfg, device=dev)
    eng.load_state_dict(make_device_weights(cfg, 0, dev))
    eng.set_graphs(False)
    if what == "dit":
        x, c, c0 = [rn(16, 1, 60, 104) for _ in range(2)], [rn(512, 4096) for _ in range(2)], [rn(512, 4096)] * 2
        t = torch.full((2,), 500.0, device=dev)
        run = lambda: eng.forward_cfg(x, t, c, c0, 1560, 5.0)
    else:
        x, c, v = [rn(16, 1, 60, 104)], [rn(512, 4096)], [rn(16, 1, 60, 104)]
        eng.zero_grad()
        run = lambda: P.student_step(eng, x, c, v, ffn_grad_blocks=None)
with torch.no_grad():
    run(); run()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    run()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done", what)
