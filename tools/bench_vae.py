"""WanVAE decode timing on the B200 engine (SURVEY 8d config 5's decode leg):
python tools/bench_vae.py [T ...]   -> one JSON line per T with frames/s, conv TFLOP/s, category split."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200dit


def main():
    Ts = [int(a) for a in sys.argv[1:]] or [1, 5, 21]
    torch.cuda.set_device(0)
    sd = b200dit.synthetic.vae_decoder_weights(dim=96, seed=0)
    eng = b200dit.VaeEngine.from_state_dict(sd)
    for T in Ts:
        z = torch.randn(16, T, 60, 104, generator=torch.Generator().manual_seed(T)).cuda()
        for _ in range(2):
            out = eng.decode([z])[0]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3 if T > 5 else 5
        e0.record()
        for _ in range(n):
            out = eng.decode([z])[0]
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        b200dit.profile_enable(True)
        eng.decode([z])
        torch.cuda.synchronize()
        prof = b200dit.profile_collect()
        b200dit.profile_enable(False)
        fl = b200dit.flops.vae_decode_flops(T)
        frames = 1 + 4 * (T - 1)
        conv = prof["conv"]
        print(json.dumps({"T": T, "frames": frames, "ms": ms, "frames_per_s": frames / (ms / 1e3),
                          "algorithmic_tflops": fl / 1e12, "achieved_tflops": fl / (ms / 1e3) / 1e12,
                          "conv_kernel_tflops": conv["flops"] / (conv["ms"] / 1e3) / 1e12 if conv["ms"] else None,
                          "ms_by_category": {k: round(v["ms"], 2) for k, v in prof.items()},
                          "finite": bool(torch.isfinite(out).all()), "shape": list(out.shape)}), flush=True)


if __name__ == "__main__":
    main()
