"""One student training step (forward + backward) on a 2-layer 1.3B-width model: the launch list target for
`ncu --metrics gpu__time_duration.sum` (developer tool)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200dit  # noqa: E402
from b200dit import pipelines as P  # noqa: E402
from bench import CFG_13B, make_device_weights  # noqa: E402

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
cfg = dict(CFG_13B, num_layers=layers)
eng = b200dit.DitEngine(**cfg, device=dev)
eng.load_state_dict(make_device_weights(cfg, 0, dev))
g = torch.Generator().manual_seed(7)
x = [torch.randn(16, 1, 60, 104, generator=g).to(dev) for _ in range(batch)]
c = [torch.randn(512, 4096, generator=g).to(dev) for _ in range(batch)]
v = [torch.randn(16, 1, 60, 104, generator=g).to(dev) for _ in range(batch)]
eng.zero_grad()
P.student_step(eng, x, c, v, ffn_grad_blocks=None)       # warm-up: workspaces, transposed weights
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
P.student_step(eng, x, c, v, ffn_grad_blocks=None)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
