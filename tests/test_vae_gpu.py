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
