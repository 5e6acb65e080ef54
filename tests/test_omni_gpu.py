"""OmniHuman audio front-end through the C ABI (b200omni_audio_tokens) against the golden vector produced by the
unmodified OmniConditionsModule.process_audio (tests/golden/omni_audio_tiny.pt) and the CPU oracle at the real widths
(wav2vec 1024 -> Wan 1.3B dim 1536, 21 frames).  Tolerance: fp16 GEMM operands, fp32 accumulate: rel-L2 <= 1e-3."""
import os

import pytest
import torch

from conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu


def test_golden_omni_audio_tiny():
    import b200dit
    g = torch.load(os.path.join(GOLDEN, "omni_audio_tiny.pt"), map_location="cpu", weights_only=True)
    ap = b200dit.AudioProcessor({k: v.float() for k, v in g["sd"].items()})
    for c in g["cases"]:
        out = ap(c["feats"])
        assert out.shape == c["out"].shape
        assert rel_l2(out.cpu(), c["out"]) < 1e-3


def test_omni_audio_real_widths_vs_oracle():
    import b200dit
    from oracle import omni_oracle as OO
    gen = torch.Generator().manual_seed(2)
    sd = {"0.weight": (torch.randn(1536, 1024, generator=gen) / 32).half().float(), "0.bias": torch.randn(1536, generator=gen) * 0.02,
          "2.weight": (torch.randn(1536, 1536, generator=gen) / 39).half().float(), "2.bias": torch.randn(1536, generator=gen) * 0.02}
    ap = b200dit.AudioProcessor({"audio_processor." + k: v for k, v in sd.items()})     # parent-module key prefix accepted
    feats = torch.randn(2, 21, 1024, generator=gen)
    out = ap.process_audio(feats)
    ref = OO.process_audio(sd, feats)
    assert out.shape == (2, 20, 3072)
    assert rel_l2(out.cpu(), ref) < 1e-3
