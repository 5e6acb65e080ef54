"""WanVAE decode parity through the C ABI (b200vae_decode) against the golden vector produced by the
unmodified reference (tests/golden/vae_tiny.pt) and the CPU oracle at the real decoder width.
Tolerance (pixels in [-1,1]): max-abs <= 1e-2, rel-L2 <= 5e-3 -- the reference's own
cuDNN-TF32 path sits ~1e-3 from fp32; operands here are fp16, accumulation fp32."""
import os

import pytest
import torch

from conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu


def _check(pix, ref):
    assert pix.shape == ref.shape
    assert float((pix - ref).abs().max()) < 1e-2
    assert rel_l2(pix, ref) < 5e-3


def test_golden_vae_tiny():
    import b200dit
    g = torch.load(os.path.join(GOLDEN, "vae_tiny.pt"), map_location="cpu", weights_only=True)
    eng = b200dit.VaeEngine.from_state_dict({k: v.float() for k, v in g["sd"].items()})
    pix = eng.decode([g["z"]])[0].cpu()
    _check(pix, g["out"])
    assert float(pix.max()) <= 1.0 and float(pix.min()) >= -1.0


@pytest.mark.parametrize("T,h,w", [(1, 8, 8), (6, 6, 10)])
def test_vae_dim96_vs_oracle(T, h, w):
    """real decoder width (384/192/96 channels); T=6 crosses a chunk boundary (chunks 1,4,1)"""
    import b200dit
    from oracle import vae_oracle as VO
    sd = VO.make_synthetic_vae_weights(dim=96, seed=3)
    eng = b200dit.VaeEngine.from_state_dict(sd)
    z = torch.randn(16, T, h, w, generator=torch.Generator().manual_seed(T))
    pix = eng.decode([z])[0].cpu()
    ref = VO.vae_decode(sd, z)
    _check(pix, ref)


def test_vae_decode_full_resolution_60x104():
    """The real decode size of every config: latent [16,2,60,104] -> [3,5,480,832] (first chunk of one frame +
    a second chunk through the temporal upsample; 480x832 exercises the implicit-GEMM pixel-tile picker and TMA
    border clipping at 60/120/240/480 x 104/208/416/832) against the CPU oracle (vae.py:544-568; ~20-40 s of CPU)."""
    import b200dit
    from oracle import vae_oracle as VO
    sd = VO.make_synthetic_vae_weights(dim=96, seed=3)
    eng = b200dit.VaeEngine.from_state_dict(sd)
    z = torch.randn(16, 2, 60, 104, generator=torch.Generator().manual_seed(11))
    pix = eng.decode([z])[0].cpu()
    with torch.no_grad():
        ref = VO.vae_decode(sd, z)
    print(f"\nVAE 60x104: max-abs {float((pix - ref).abs().max()):.3e} rel-L2 {rel_l2(pix, ref):.3e}")
    _check(pix, ref)


def test_vae_halo_conv_partial_tiles_and_determinism():
    """The halo convolution (csrc/conv_tc.cu) tiles a frame into columns of 16 x 8 pixel sub-tiles: a 7 x 9 latent
    gives 56 x 72 pixels at the last stage (a half-empty last sub-tile row, bands of 5 + a short band) and odd
    sizes at every earlier stage; T = 3 crosses the chunk boundary with the fused norm epilogues writing the
    next conv's causal operand.  Against the CPU oracle (vae.py:544-568), and twice: the result must not depend
    on the timing of the asynchronous staging stores."""
    import b200dit
    from oracle import vae_oracle as VO
    sd = VO.make_synthetic_vae_weights(dim=96, seed=5)
    eng = b200dit.VaeEngine.from_state_dict(sd)
    z = torch.randn(16, 3, 7, 9, generator=torch.Generator().manual_seed(3))
    pix = eng.decode([z])[0].cpu()
    with torch.no_grad():
        ref = VO.vae_decode(sd, z)
    _check(pix, ref)
    for _ in range(3):
        assert torch.equal(eng.decode([z])[0].cpu(), pix)


def test_install_vae_shim():
    import b200dit
    from oracle import vae_oracle as VO
    sd = VO.make_synthetic_vae_weights(dim=8, seed=6)

    class FakeVAE:                                   # surface of WanVAE (vae.py:619-663): .model, .decode(zs)
        class _M:
            def state_dict(self):
                return sd
        model = _M()

        def decode(self, zs):
            raise RuntimeError("reference decode must not run")

    v = FakeVAE()
    b200dit.install_vae(v)
    z = torch.randn(16, 2, 4, 6, generator=torch.Generator().manual_seed(1))
    _check(v.decode([z])[0].cpu(), VO.vae_decode(sd, z))


def test_golden_vae_encode_tiny():
    """b200vae_encode vs the unmodified reference WanVAE_.encode (tests/golden/vae_enc_tiny.pt): first chunk of one
    frame + two chunks of four (stride-2 spatial convs, stride-2 temporal convs with the one-frame history).
    Tolerance: fp16 operands / fp32 accumulation through 20+ convs on O(0.4) latents: max-abs <= 1e-2."""
    import b200dit
    g = torch.load(os.path.join(GOLDEN, "vae_enc_tiny.pt"), map_location="cpu", weights_only=True)
    sd = {k: v.float() for k, v in g["sd"].items()}
    vae = b200dit.VaeEngine.from_state_dict(sd)
    assert vae.has_encoder
    mu = vae.encode([g["video"].float()])[0].cpu()
    assert mu.shape == g["out"].shape
    assert float((mu - g["out"]).abs().max()) < 1e-2
    mu1 = vae.encode([g["video"].float()[:, :1]])[0].cpu()             # a single image
    assert float((mu1 - g["out"][:, :1]).abs().max()) < 1e-2


def test_vae_encode_dim96_vs_oracle():
    """Full-width encoder (dim 96) on a small clip vs the CPU oracle; then decode(encode(x)) runs end to end."""
    import b200dit
    from oracle import vae_oracle as VO
    sd = VO.make_synthetic_vae_weights(dim=96, seed=3, encoder=True)
    vae = b200dit.VaeEngine.from_state_dict(sd)
    video = torch.rand(3, 5, 64, 96, generator=torch.Generator().manual_seed(4)) * 2 - 1
    mu = vae.encode([video])[0]
    ref = VO.vae_encode(sd, video, dim=96)
    assert float((mu.cpu() - ref).abs().max()) < 1e-2
    rec = vae.decode([mu])[0]
    assert rec.shape == video.shape and bool(torch.isfinite(rec).all())


def test_pipelined_decode_two_gpus_matches_single(tmp_path):
    """F3b: two ranks decode one latent together (b200vae_decode_pipelined: chunks alternate between the ranks, every
    causal conv's two-frame cache crosses to the other GPU through peer memory).  With the single-GPU chunking (4
    latent frames per chunk) the video is bit-identical to the single-GPU decode; with one frame per chunk the
    convolution GEMMs see other tile counts (other K-split tails), so fp16 operand roundings may differ by an ulp:
    max-abs <= 5e-3 on pixels in [-1, 1] (the oracle bar is 1e-2).  Needs two GPUs (skipped on a one-GPU box; run
    with `gpurun --gpus 2`)."""
    import json
    import socket
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    for extra, bar in ((["--frames", "6", "--h", "12", "--w", "16", "--chunk", "1"], 5e-3),
                       (["--frames", "7", "--h", "30", "--w", "52"], 0.0)):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.join(root, "tools", "vae_pipe.py"), "--reps", "1"] + extra
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        assert line["finite"] and line["max_abs_vs_single_gpu"] <= bar, line
