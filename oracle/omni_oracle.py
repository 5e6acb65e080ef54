"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the OmniHuman audio front-end.

Pinned against the UNMODIFIED `OmniConditionsModule` (Omnihuman/omnihuman_wan_t2v.py:13-94) executed in the
authoring container (oracle/ref_loader.load_reference_omni, tests/test_cpu_oracle.py) and against the fixture it
produced (tests/golden/omni_audio_tiny.pt, oracle/make_golden.py omni).  Only tests/ and smoke() may import it.
"""
import torch
import torch.nn.functional as F


def process_audio(sd, feats):
    """omnihuman_wan_t2v.py:55-60 (== :180-200): audio_processor = Linear . SiLU . Linear (:30-34) per frame, then
    frames t and t+1 concatenated on the channel axis when T > 1.  sd: `0.weight`, `0.bias`, `2.weight`, `2.bias`.
    feats [B, T, audio_dim] -> [B, T-1, 2 D] or [B, 1, D]."""
    h = F.silu(F.linear(feats.float(), sd["0.weight"], sd.get("0.bias")))
    tok = F.linear(h, sd["2.weight"], sd.get("2.bias"))
    if tok.shape[1] > 1:
        tok = torch.cat([tok[:, :-1], tok[:, 1:]], dim=-1)
    return tok
