"""Installs the B200 engine over live reference objects without touching their callers.

The reference's own plug-in convention is instance-level monkey-patching
(`types.MethodType(usp_dit_forward, self.model)`, seaweed_apt/wan/text2video.py:95-98); `install`
does the same for `WanModel.forward` (model.py:502) and `install_vae` for `WanVAE.decode`
(vae.py:657) and, when the encoder weights are present, `WanVAE.encode` (vae.py:641).  Callers -- WanT2V.generate (text2video.py:238-241,259), generate.py:227-228,
distilled_trainer.py:273-278, eval_ema.py:124,140 -- keep calling `model(x, t=..., context=...,
seq_len=...)` / `vae.decode(zs)` unchanged.

The engine is inference-only: when autograd is recording and any input or parameter requires a
gradient, the call falls through to the ORIGINAL reference forward (the unmodified PyTorch path,
not a CPU fallback) -- SURVEY.md section 8b "Autograd".
"""
import types

import torch

from .engine import DitEngine, VaeEngine


def _needs_grad(model, x):
    if not torch.is_grad_enabled():
        return False
    if any(getattr(u, "requires_grad", False) for u in x):
        return True
    return any(p.requires_grad for p in model.parameters())


def install(model, device=None, engine=None):
    """Routes `model.forward` through a DitEngine built from the module's own weights."""
    eng = engine or DitEngine.from_module(model, device=device)
    original = model.forward

    def forward(self, x, t, context, seq_len, clip_fea=None, y=None):
        xs = list(x) if not isinstance(x, (list, tuple)) else x
        if _needs_grad(self, xs):
            return original(x, t, context, seq_len, clip_fea=clip_fea, y=y)
        return eng.forward(xs, t, context, seq_len, clip_fea=clip_fea, y=y)

    model._b200_original_forward = original
    model._b200_engine = eng
    model.forward = types.MethodType(forward, model)
    return eng


def install_vae(vae, device=None, engine=None):
    """Routes `WanVAE.decode` (vae.py:657-663) through a VaeEngine built from `vae.model`."""
    eng = engine or VaeEngine.from_state_dict(vae.model.state_dict(), device=device)
    original = vae.decode

    def decode(self, zs):
        return eng.decode(zs)

    vae._b200_original_decode = original
    vae._b200_engine = eng
    vae.decode = types.MethodType(decode, vae)
    if getattr(eng, "has_encoder", False) and hasattr(vae, "encode"):      # WanVAE.encode (vae.py:641-655)
        vae._b200_original_encode = vae.encode
        vae.encode = types.MethodType(lambda self, videos: eng.encode(videos), vae)
    return eng


def uninstall(obj):
    if hasattr(obj, "_b200_original_forward"):
        obj.forward = obj._b200_original_forward
        del obj._b200_original_forward, obj._b200_engine
    if hasattr(obj, "_b200_original_encode"):
        obj.encode = obj._b200_original_encode
        del obj._b200_original_encode
    if hasattr(obj, "_b200_original_decode"):
        obj.decode = obj._b200_original_decode
        del obj._b200_original_decode, obj._b200_engine
