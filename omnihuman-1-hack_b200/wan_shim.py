"""Installs the B200 engine over live reference objects without touching their callers.

The reference's own plug-in convention is instance-level monkey-patching
(`types.MethodType(usp_dit_forward, self.model)`, seaweed_apt/wan/text2video.py:95-98); `install`
does the same for `WanModel.forward` (model.py:502) and `install_vae` for `WanVAE.decode`
(vae.py:657) and, when the encoder weights are present, `WanVAE.encode` (vae.py:641).  Callers -- WanT2V.generate (text2video.py:238-241,259), generate.py:227-228,
distilled_trainer.py:273-278, eval_ema.py:124,140 -- keep calling `model(x, t=..., context=...,
seq_len=...)` / `vae.decode(zs)` unchanged.

Autograd: when autograd is recording and an input or parameter requires a
gradient, a t2v call (the student step of seaweed_apt/distilled_trainer.py:268-301) becomes one autograd node per
latent grid whose backward runs on the engine (`autograd.dit_forward`); calls the engine's backward does not cover
(`y` / `clip_fea`, contexts that require gradients) fall through to the ORIGINAL reference forward -- the
unmodified PyTorch path, not a CPU fallback.
"""
import math
import random
import sys
import types

import torch

from . import autograd as b200_autograd
from . import pipelines
from .engine import DitEngine, VaeEngine


def _needs_grad(model, x):
    if not torch.is_grad_enabled():
        return False
    if any(getattr(u, "requires_grad", False) for u in x):
        return True
    return any(p.requires_grad for p in model.parameters())


def _weights_signature(model):
    """Changes whenever a parameter is updated in place (optimizer.step, EMA copy_, load_state_dict) or replaced:
    the sum of the tensors' version counters plus their storage addresses."""
    if not hasattr(model, "parameters"):
        return None
    ver, ptr = 0, 0
    for p in model.parameters():
        ver += p._version
        ptr ^= p.data_ptr()
    return ver, ptr


def install(model, device=None, engine=None, ffn_grad_blocks=11):
    """Routes `model.forward` through a DitEngine built from the module's own weights.

    Under autograd the backward runs on the engine too; `ffn_grad_blocks=11` reproduces the reference's blocks with
    block_idx > 10, whose FFN runs under no_grad (model.py:318-325), None differentiates every FFN.

    The engine holds a SNAPSHOT of the weights (fp16 GEMM operands repacked on the device).  The reference calls
    the generator under `torch.no_grad()` while it is being trained (seaweed_apt/apt_trainer.py:118-119,254), so
    every call compares the parameters' version counters with the snapshot's and reloads the engine when an
    optimizer step / EMA update / load_state_dict has touched them -- never a stale answer."""
    eng = engine or DitEngine.from_module(model, device=device)
    original = model.forward
    state = {"sig": _weights_signature(model)}

    def forward(self, x, t, context, seq_len, clip_fea=None, y=None):
        xs = list(x) if not isinstance(x, (list, tuple)) else x
        grad = _needs_grad(self, xs)
        if grad and (not b200_autograd.supports(eng, clip_fea, y)
                     or any(getattr(c, "requires_grad", False) for c in context)):
            return original(x, t, context, seq_len, clip_fea=clip_fea, y=y)
        sig = _weights_signature(self)
        if sig != state["sig"]:
            eng.load_state_dict(self.state_dict())
            state["sig"] = sig
            self._b200_reloads = getattr(self, "_b200_reloads", 0) + 1
        if grad:
            return b200_autograd.dit_forward(eng, list(self.named_parameters()), xs, t, context, seq_len,
                                             ffn_grad_blocks=ffn_grad_blocks)
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


def install_t2v(pipe, device=None, dit_engine=None, vae_engine=None):
    """Routes `WanT2V.generate` (seaweed_apt/wan/text2video.py:111-269) through the engines, signature unchanged:
    prompt encoding stays the pipeline's own T5 (out of scope), noise comes from the same seeded
    device generator (:167-169,186-195), and the loop (:231-252: cond + uncond forward, CFG, scheduler step) and
    the decode (:258-259) run as `pipelines.sample` + `VaeEngine.decode`.  `offload_model` is accepted and has
    nothing to do: the engines keep their own device copies of the weights."""
    eng = dit_engine or DitEngine.from_module(pipe.model, device=device)
    vae = vae_engine or VaeEngine.from_state_dict(pipe.vae.model.state_dict(), device=device)
    pipe._b200_original_generate = pipe.generate
    pipe._b200_engines = (eng, vae)
    pipe.generate = types.MethodType(_t2v_generate, pipe)
    return eng, vae


def _t2v_generate(self, input_prompt, size=(720, 512), frame_num=81, shift=5.0, sample_solver='unipc',
                  sampling_steps=50, guide_scale=5.0, n_prompt="", seed=-1, offload_model=True):
    """Body of the patched `WanT2V.generate`; parameter names and defaults are the reference's (text2video.py:111-121,
    checked by tests/test_cpu_abi.py)."""
    eng, vae = self._b200_engines
    if sample_solver not in ("unipc", "dpm++"):
        raise NotImplementedError("Unsupported solver.")                                   # :221-222
    st, ps, sp = self.vae_stride, self.patch_size, getattr(self, "sp_size", 1)
    shape = (self.vae.model.z_dim, (frame_num - 1) // st[0] + 1, size[1] // st[1], size[0] // st[2])
    seq_len = math.ceil(shape[2] * shape[3] / (ps[1] * ps[2]) * shape[1] / sp) * sp        # :161-164
    negative = n_prompt if n_prompt != "" else self.sample_neg_prompt
    gen = torch.Generator(device=self.device)
    gen.manual_seed(seed if seed >= 0 else random.randint(0, sys.maxsize))
    enc_dev = torch.device("cpu") if self.t5_cpu else self.device                          # :172-183
    if not self.t5_cpu:
        self.text_encoder.model.to(self.device)
    context = [c.to(self.device) for c in self.text_encoder([input_prompt], enc_dev)]
    context_null = [c.to(self.device) for c in self.text_encoder([negative], enc_dev)]
    if not self.t5_cpu and offload_model:
        self.text_encoder.model.cpu()
    noise = [torch.randn(*shape, dtype=torch.float32, device=self.device, generator=gen)]
    x0 = pipelines.sample(eng, noise, context, context_null, steps=sampling_steps, shift=shift,
                          guide_scale=guide_scale, solver=sample_solver, seq_len=seq_len)
    videos = vae.decode(x0) if self.rank == 0 else None
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.barrier()
    return videos[0] if self.rank == 0 else None


def uninstall(obj):
    if hasattr(obj, "_b200_original_generate"):
        obj.generate = obj._b200_original_generate
        del obj._b200_original_generate, obj._b200_engines
    if hasattr(obj, "_b200_original_forward"):
        obj.forward = obj._b200_original_forward
        del obj._b200_original_forward, obj._b200_engine
    if hasattr(obj, "_b200_original_encode"):
        obj.encode = obj._b200_original_encode
        del obj._b200_original_encode
    if hasattr(obj, "_b200_original_decode"):
        obj.decode = obj._b200_original_decode
        del obj._b200_original_decode, obj._b200_engine
