"""Developer probe (gpurun): tiny i2v engine, one call per variant, each in its own process."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(mode):
    import torch
    import b200dit
    from oracle import dit_oracle as O
    sd = O.make_synthetic_weights(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, in_dim=32, i2v=True, seed=4)
    eng = b200dit.DitEngine.from_state_dict(sd, num_heads=2)
    g = torch.Generator().manual_seed(1)
    x0, y = torch.randn(16, 2, 8, 12, generator=g), torch.randn(16, 2, 8, 12, generator=g)
    ctx, ctx0 = torch.randn(30, 64, generator=g), torch.randn(9, 64, generator=g)
    clip = torch.randn(1, 257, 1280, generator=g)
    t = torch.tensor([500.0])
    for rep in range(4):
        if mode == "cfg":
            out = eng.forward_cfg([x0], t, [ctx], [ctx0], 48, 7.5, clip_fea=clip, y=[y])
        elif mode == "fwd1":
            out = eng.forward([x0], t, [ctx], 48, clip_fea=clip, y=[y])
        else:
            out = eng.forward([x0, x0], torch.tensor([500.0, 500.0]), [ctx, ctx0], 48, clip_fea=torch.cat([clip, clip]), y=[y, y])
        torch.cuda.synchronize()
        print(f"  {mode} rep {rep}: ok, finite {bool(torch.isfinite(out[0]).all())}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        child(sys.argv[2])
    else:
        for env_extra, mode in [({}, "fwd1"), ({}, "fwd2"), ({}, "cfg"), ({"B200_PDL": "0"}, "cfg"), ({"B200_ATTN_V1": "1"}, "cfg"),
                                ({"CUDA_LAUNCH_BLOCKING": "1"}, "cfg")]:
            print("variant", env_extra, mode, flush=True)
            r = subprocess.run(["timeout", "120", sys.executable, os.path.abspath(__file__), "child", mode],
                               env=dict(os.environ, **env_extra), capture_output=True, text=True)
            print(r.stdout[-600:], r.stderr[-900:], flush=True)
