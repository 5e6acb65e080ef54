"""WanVAE decode / encode timing on the B200 engine (SURVEY 8d config 5's decode leg; F3 encode):
python tools/bench_vae.py [T ...]          -> one JSON line per latent length T: frames/s, conv TFLOP/s, category split
python tools/bench_vae.py encode [F ...]   -> one JSON line per clip length F (frames, 1 + 4k) at 480x832"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200dit


def main_encode():
    Fs = [int(a) for a in sys.argv[2:]] or [1, 17, 81]
    torch.cuda.set_device(0)
    eng = b200dit.VaeEngine.from_state_dict(b200dit.synthetic.vae_decoder_weights(dim=96, seed=0, encoder=True))
    for Fr in Fs:
        v = (torch.rand(3, Fr, 480, 832, generator=torch.Generator().manual_seed(Fr)) * 2 - 1).cuda()
        for _ in range(2):
            out = eng.encode([v])[0]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3
        e0.record()
        for _ in range(n):
            out = eng.encode([v])[0]
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = b200dit.flops.vae_encode_flops(Fr)
        print(json.dumps({"encode_frames": Fr, "ms": ms, "frames_per_s": Fr / (ms / 1e3),
                          "algorithmic_tflops": fl / 1e12, "achieved_tflops": fl / (ms / 1e3) / 1e12,
                          "finite": bool(torch.isfinite(out).all()), "shape": list(out.shape)}), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "encode":
        return main_encode()
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
