"""Host side of the OmniHuman omni-conditions front-end, mirroring
`OmniConditionsModule.process_audio` / `OmniHumanWanT2V.process_audio`
(Omnihuman/omnihuman_wan_t2v.py:55-60, 180-200): wav2vec features [B, T, audio_dim] -> audio tokens.

    ap = b200dit.AudioProcessor.from_module(omni.audio_processor)      # nn.Sequential(Linear, SiLU, Linear), :30-34
    tokens = ap(audio_features)                                        # [B, T-1, 2 D]  (T > 1),  [B, 1, D]  (T = 1)

The arithmetic runs in libb200dit.so (`b200omni_audio_tokens`: two tcgen05 GEMMs, SiLU + cast, adjacent-frame
concat).  `process_pose` is not mirrored: as shipped it feeds [B, T, K, H, W] to a Conv3d that expects the K
key-point channels on axis 1 and raises for every input the docstring describes, so there is
no reference behaviour to pin parity to; the pose stack enters the DiT through the `y` channel hook instead.
"""
import ctypes as C

import torch

from ._lib import B200Error, check, lib
from .engine import _stream_ptr


class AudioProcessor:
    def __init__(self, state_dict, device=None):
        """`state_dict`: the audio_processor Sequential's own keys -- `0.weight` [D, audio_dim], `0.bias`, `2.weight`
        [D, D], `2.bias` (an `audio_processor.` prefix, as in the parent module's state_dict, is accepted)."""
        if not torch.cuda.is_available():
            raise B200Error("no CUDA device: the B200 engine has no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        sd = {k[len("audio_processor."):] if k.startswith("audio_processor.") else k: v for k, v in state_dict.items()}
        missing = [k for k in ("0.weight", "2.weight") if k not in sd]
        if missing:
            raise B200Error(f"audio_processor weights missing: {missing}")
        to = lambda t, dt: t.detach().to(self.device, dt).contiguous()
        self.w0, self.w2 = to(sd["0.weight"], torch.float16), to(sd["2.weight"], torch.float16)
        self.b0 = to(sd["0.bias"], torch.float32) if "0.bias" in sd else None
        self.b2 = to(sd["2.bias"], torch.float32) if "2.bias" in sd else None
        self.model_dim, self.audio_dim = self.w0.shape
        if tuple(self.w2.shape) != (self.model_dim, self.model_dim):
            raise B200Error("audio_processor.2.weight must be [model_dim, model_dim]")
        self._scratch = None

    @classmethod
    def from_module(cls, seq, device=None):
        return cls(seq.state_dict(), device=device)

    @staticmethod
    def scratch_bytes(rows, audio_dim, model_dim):
        return rows * (2 * audio_dim + 10 * model_dim) + 1024

    @torch.no_grad()
    def __call__(self, audio_features):
        x = audio_features.to(self.device, torch.float32).contiguous()
        if x.dim() != 3 or x.shape[2] != self.audio_dim:
            raise B200Error(f"audio features must be [B, T, {self.audio_dim}]")
        B, T, _ = x.shape
        D = self.model_dim
        out = torch.empty((B, T - 1, 2 * D) if T > 1 else (B, 1, D), dtype=torch.float32, device=self.device)
        need = self.scratch_bytes(B * T, self.audio_dim, D)
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().b200omni_audio_tokens(
                C.c_void_p(x.data_ptr()), B, T, self.audio_dim, D, C.c_void_p(self.w0.data_ptr()),
                C.c_void_p(self.b0.data_ptr() if self.b0 is not None else 0), C.c_void_p(self.w2.data_ptr()),
                C.c_void_p(self.b2.data_ptr() if self.b2 is not None else 0), C.c_void_p(out.data_ptr()),
                C.c_void_p(self._scratch.data_ptr()), self._scratch.numel(), _stream_ptr()))
        return out

    process_audio = __call__
