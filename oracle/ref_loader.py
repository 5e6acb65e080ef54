"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference modules by file path.

Used in the authoring container (where /root/reference is mounted) to
  (a) validate the in-repo restatements in oracle/dit_oracle.py / oracle/vae_oracle.py, and
  (b) generate the committed golden fixtures under tests/golden/ (oracle/make_golden.py).
The GPU box has no /root/reference: nothing under tests -m gpu, smoke() or bench.py calls this.

Shims:
  * `diffusers.configuration_utils.{ConfigMixin, register_to_config}` and
    `diffusers.models.modeling_utils.ModelMixin` stubs (model.py:7-8 import them; diffusers is
    not installed here),
  * a `logger` module with a `.logger` attribute (model.py:10),
  * `model.flash_attention` rebound to an exact masked softmax attention on CPU, because
    attention.py:54 asserts CUDA. Key j of item b is valid iff j < k_lens[b] (attention.py:72-80).
No reference source is copied; the files are executed from where they lie.
"""
import importlib.util
import logging
import os
import sys
import types

import torch

REF_ROOTS = [os.environ.get("B200DIT_REF", ""), "/root/reference"]


def find_reference():
    for r in REF_ROOTS:
        if r and os.path.isfile(os.path.join(r, "seaweed_apt/wan/modules/model.py")):
            return r
    return None


def _masked_attention(q, k, v, q_lens=None, k_lens=None, dropout_p=0., softmax_scale=None,
                      q_scale=None, causal=False, window_size=(-1, -1), deterministic=False,
                      dtype=torch.bfloat16, version=None):
    """CPU stand-in for attention.py:24-130 (exact softmax attention, fp32)."""
    assert not causal and q_lens is None
    b, lq, n, d = q.shape
    lk = k.shape[1]
    out_dtype = q.dtype
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    if q_scale is not None:
        qf = qf * q_scale
    scale = softmax_scale if softmax_scale is not None else d ** -0.5
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    if k_lens is not None:
        kl = torch.as_tensor(k_lens).view(b, 1, 1, 1).clamp(max=lk)
        mask = torch.arange(lk).view(1, 1, 1, lk) >= kl
        s = s.masked_fill(mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, vf).permute(0, 2, 1, 3).contiguous()
    return o.to(out_dtype)


def _install_stubs():
    if "diffusers" not in sys.modules:
        d = types.ModuleType("diffusers")
        cu = types.ModuleType("diffusers.configuration_utils")
        mo = types.ModuleType("diffusers.models")
        mu = types.ModuleType("diffusers.models.modeling_utils")

        class _Config(dict):
            def __getattr__(self, name):
                try:
                    return self[name]
                except KeyError:
                    raise AttributeError(name) from None

        class ConfigMixin:
            """Minimal stand-in for diffusers' ConfigMixin: `self.config` holds the constructor arguments."""

            def register_to_config(self, **kw):
                cfg = _Config(getattr(self, "_b200_cfg", {}))
                cfg.update(kw)
                object.__setattr__(self, "_b200_cfg", cfg)

            @property
            def config(self):
                return self._b200_cfg

        def register_to_config(init):
            import functools
            import inspect

            @functools.wraps(init)
            def inner(self, *args, **kwargs):
                names = [n for i, n in enumerate(inspect.signature(init).parameters) if i > 0]
                defaults = {n: p.default for n, p in inspect.signature(init).parameters.items() if n != "self"}
                vals = dict(defaults)
                vals.update(dict(zip(names, args)))
                vals.update(kwargs)
                self.register_to_config(**vals)
                init(self, *args, **kwargs)
            return inner

        class ModelMixin(torch.nn.Module):
            pass

        cu.ConfigMixin, cu.register_to_config = ConfigMixin, register_to_config
        mu.ModelMixin = ModelMixin
        d.configuration_utils, d.models = cu, mo
        mo.modeling_utils = mu
        # scheduler-side imports of fm_solvers*.py (:11-16)
        import enum
        sch = types.ModuleType("diffusers.schedulers")
        su = types.ModuleType("diffusers.schedulers.scheduling_utils")
        ut = types.ModuleType("diffusers.utils")
        tu = types.ModuleType("diffusers.utils.torch_utils")

        class KarrasDiffusionSchedulers(enum.Enum):
            DDIMScheduler = 1

        class SchedulerMixin:
            pass

        class SchedulerOutput:
            def __init__(self, prev_sample):
                self.prev_sample = prev_sample

        def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
            return torch.randn(shape, generator=generator, device=device, dtype=dtype)

        su.KarrasDiffusionSchedulers, su.SchedulerMixin, su.SchedulerOutput = KarrasDiffusionSchedulers, SchedulerMixin, SchedulerOutput
        ut.deprecate = lambda *a, **k: None
        ut.is_scipy_available = lambda: False
        tu.randn_tensor = randn_tensor
        ut.torch_utils = tu
        sch.scheduling_utils = su
        d.schedulers, d.utils = sch, ut
        sys.modules.update({"diffusers": d, "diffusers.configuration_utils": cu,
                            "diffusers.models": mo, "diffusers.models.modeling_utils": mu,
                            "diffusers.schedulers": sch, "diffusers.schedulers.scheduling_utils": su,
                            "diffusers.utils": ut, "diffusers.utils.torch_utils": tu})
    if "logger" not in sys.modules:
        lg = types.ModuleType("logger")
        lg.logger = logging.getLogger("ref")
        sys.modules["logger"] = lg


def load_reference_modules(root=None):
    """Returns (model_module, vae_module) of the reference, executed in place."""
    root = root or find_reference()
    if root is None:
        raise FileNotFoundError("reference tree not available (container-only tool)")
    _install_stubs()
    pkg_name = "_refwan_modules"
    if pkg_name + ".model" in sys.modules:
        return sys.modules[pkg_name + ".model"], sys.modules[pkg_name + ".vae"]
    mdir = os.path.join(root, "seaweed_apt/wan/modules")
    pkg = types.ModuleType(pkg_name)
    pkg.__path__ = [mdir]
    sys.modules[pkg_name] = pkg
    mods = {}
    for name in ("attention", "model", "vae"):
        spec = importlib.util.spec_from_file_location(f"{pkg_name}.{name}", os.path.join(mdir, f"{name}.py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"{pkg_name}.{name}"] = m
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec.loader.exec_module(m)
        mods[name] = m
    mods["model"].flash_attention = _masked_attention
    # model.py:503 calls torch.cuda.empty_cache(); harmless without CUDA.
    return mods["model"], mods["vae"]


def load_reference_solvers(root=None):
    """Returns (fm_solvers_unipc module, fm_solvers module) of the reference, executed in place
    (seaweed_apt/wan/utils/fm_solvers_unipc.py, fm_solvers.py)."""
    root = root or find_reference()
    if root is None:
        raise FileNotFoundError("reference tree not available (container-only tool)")
    _install_stubs()
    udir = os.path.join(root, "seaweed_apt/wan/utils")
    out = []
    for name in ("fm_solvers_unipc", "fm_solvers"):
        full = "_refwan_utils_" + name
        if full not in sys.modules:
            spec = importlib.util.spec_from_file_location(full, os.path.join(udir, name + ".py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules[full] = m
            spec.loader.exec_module(m)
        out.append(sys.modules[full])
    return tuple(out)


def load_reference_apt(root=None):
    """Returns the reference's seaweed_apt/model.py module (WanCrossAttentionDiscriminatorBlock,
    WanAPTDiscriminator), executed in place; its `wan.modules.*` imports resolve to the modules loaded by
    load_reference_modules (T5 is never touched by the discriminator and is stubbed)."""
    root = root or find_reference()
    if root is None:
        raise FileNotFoundError("reference tree not available (container-only tool)")
    full = "_ref_apt_model"
    if full in sys.modules:
        return sys.modules[full]
    model, vae = load_reference_modules(root)
    wan = types.ModuleType("wan"); wan.__path__ = []
    wm = types.ModuleType("wan.modules"); wm.__path__ = []
    t5 = types.ModuleType("wan.modules.t5"); t5.T5EncoderModel = object
    added = {"wan": wan, "wan.modules": wm, "wan.modules.model": model, "wan.modules.t5": t5, "wan.modules.vae": vae}
    saved = {k: sys.modules.get(k) for k in added}
    sys.modules.update(added)
    try:
        spec = importlib.util.spec_from_file_location(full, os.path.join(root, "seaweed_apt/model.py"))
        m = importlib.util.module_from_spec(spec)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec.loader.exec_module(m)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    sys.modules[full] = m
    return m


def load_reference_omni(root=None):
    """Returns the reference's Omnihuman/omnihuman_wan_t2v.py module (OmniConditionsModule, OmniHumanWanT2V),
    executed in place.  Its heavyweight imports (the `wan` package, omegaconf) are only needed by
    OmniHumanWanT2V.__init__ (checkpoints, T5) and are stubbed; OmniConditionsModule -- the pre-nets -- is plain
    torch.nn and runs as shipped."""
    root = root or find_reference()
    if root is None:
        raise FileNotFoundError("reference tree not available (container-only tool)")
    full = "_ref_omnihuman_wan_t2v"
    if full in sys.modules:
        return sys.modules[full]
    _install_stubs()
    names = ["wan", "wan.modules", "wan.modules.t5", "wan.modules.vae", "wan.configs", "wan.utils", "wan.utils.fm_solvers",
             "omegaconf"]
    stubs = {n: types.ModuleType(n) for n in names}
    for n in ("wan", "wan.modules", "wan.utils"):
        stubs[n].__path__ = []
    stubs["wan"].WanT2V = object
    stubs["wan.modules.t5"].T5EncoderModel = object
    stubs["wan.modules.vae"].WanVAE = object
    stubs["wan.configs"].t2v_14B = {}
    stubs["wan.utils.fm_solvers"].FlowDPMSolverMultistepScheduler = object
    stubs["omegaconf"].DictConfig = dict
    stubs["omegaconf"].OmegaConf = object
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location(full, os.path.join(root, "Omnihuman/omnihuman_wan_t2v.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    sys.modules[full] = m
    return m
