"""Drop-in check of the reference-facing Python surface: every method the reference's callers use has the
reference's parameter names, order and defaults.  The expected signatures are written out here (they travel to the
GPU box); where the reference tree is mounted they are also re-derived from its source with `ast`."""
import ast
import inspect
import os

import pytest

from oracle import ref_loader

# (reference file, class, method) -> [(name, default or inspect._empty), ...] without `self`
E = inspect.Parameter.empty
EXPECTED = {
    ("seaweed_apt/wan/text2video.py", "WanT2V", "generate"): [
        ("input_prompt", E), ("size", (720, 512)), ("frame_num", 81), ("shift", 5.0), ("sample_solver", "unipc"),
        ("sampling_steps", 50), ("guide_scale", 5.0), ("n_prompt", ""), ("seed", -1), ("offload_model", True)],
    ("seaweed_apt/wan/modules/model.py", "WanModel", "forward"): [
        ("x", E), ("t", E), ("context", E), ("seq_len", E), ("clip_fea", None), ("y", None)],
    ("seaweed_apt/model.py", "WanAPTDiscriminator", "forward"): [
        ("x", E), ("t", E), ("context", E), ("seq_len", E), ("return_features", False)],
    ("seaweed_apt/wan/modules/vae.py", "WanVAE", "decode"): [("zs", E)],
    ("seaweed_apt/wan/modules/vae.py", "WanVAE", "encode"): [("videos", E)],
    ("seaweed_apt/wan/utils/fm_solvers_unipc.py", "FlowUniPCMultistepScheduler", "step"): [
        ("model_output", E), ("timestep", E), ("sample", E), ("return_dict", True), ("generator", None)],
    ("seaweed_apt/wan/utils/fm_solvers.py", "FlowDPMSolverMultistepScheduler", "step"): [
        ("model_output", E), ("timestep", E), ("sample", E), ("generator", None), ("variance_noise", None),
        ("return_dict", True)],
    ("seaweed_apt/wan/utils/fm_solvers_unipc.py", "FlowUniPCMultistepScheduler", "set_timesteps"): [
        ("num_inference_steps", None), ("device", None), ("sigmas", None), ("mu", None), ("shift", None)],
    ("seaweed_apt/wan/utils/fm_solvers.py", "FlowDPMSolverMultistepScheduler", "set_timesteps"): [
        ("num_inference_steps", None), ("device", None), ("sigmas", None), ("mu", None), ("shift", None)],
}


def _ours():
    import b200dit
    from b200dit import wan_shim
    return {
        ("seaweed_apt/wan/text2video.py", "WanT2V", "generate"): wan_shim._t2v_generate,
        ("seaweed_apt/wan/modules/model.py", "WanModel", "forward"): b200dit.DitEngine.forward,
        ("seaweed_apt/model.py", "WanAPTDiscriminator", "forward"): b200dit.AptDiscriminator.forward,
        ("seaweed_apt/wan/modules/vae.py", "WanVAE", "decode"): b200dit.VaeEngine.decode,
        ("seaweed_apt/wan/modules/vae.py", "WanVAE", "encode"): b200dit.VaeEngine.encode,
        ("seaweed_apt/wan/utils/fm_solvers_unipc.py", "FlowUniPCMultistepScheduler", "step"):
            b200dit.FlowUniPCMultistepScheduler.step,
        ("seaweed_apt/wan/utils/fm_solvers.py", "FlowDPMSolverMultistepScheduler", "step"):
            b200dit.FlowDPMSolverMultistepScheduler.step,
        ("seaweed_apt/wan/utils/fm_solvers_unipc.py", "FlowUniPCMultistepScheduler", "set_timesteps"):
            b200dit.FlowUniPCMultistepScheduler.set_timesteps,
        ("seaweed_apt/wan/utils/fm_solvers.py", "FlowDPMSolverMultistepScheduler", "set_timesteps"):
            b200dit.FlowDPMSolverMultistepScheduler.set_timesteps,
    }


def _sig(fn):
    params = list(inspect.signature(fn).parameters.values())
    assert params[0].name == "self"
    return [(p.name, p.default) for p in params[1:]]


@pytest.mark.parametrize("key", sorted(EXPECTED))
def test_signature_matches_reference(key):
    assert _sig(_ours()[key]) == EXPECTED[key], key


def _ref_sig(root, path, cls, method):
    tree = ast.parse(open(os.path.join(root, path)).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name == method:
                    names = [a.arg for a in f.args.args][1:]
                    defaults = [ast.literal_eval(d) for d in f.args.defaults]
                    pad = [E] * (len(names) - len(defaults))
                    return list(zip(names, pad + defaults))
    raise LookupError((path, cls, method))


@pytest.mark.skipif(ref_loader.find_reference() is None, reason="reference tree not mounted (container-only check)")
@pytest.mark.parametrize("key", sorted(EXPECTED))
def test_expected_signatures_are_the_references(key):
    assert _ref_sig(ref_loader.find_reference(), *key) == EXPECTED[key], key
