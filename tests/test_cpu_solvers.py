"""Scheduler parity on CPU: the oracle restatement (oracle/solver_oracle.py) against the golden trajectories
produced by the UNMODIFIED reference schedulers (tests/golden/solver_traj.pt, oracle/make_golden.py) and --
where /root/reference is mounted -- against the live reference classes; plus the HOST logic of the product
schedulers (coefficient composition in omnihuman-1-hack_b200/solvers.py) with the CUDA launch replaced by a
torch stand-in, which is exactly the arithmetic the kernel performs.  Tolerances: oracle bit-exact; product
coefficients are composed in one linear combination, so rounding differs: max-abs <= 2e-5 on O(1) latents."""
import os

import pytest
import torch

from conftest import GOLDEN
from oracle import ref_loader, solver_oracle as SO


def _golden():
    return torch.load(os.path.join(GOLDEN, "solver_traj.pt"), map_location="cpu", weights_only=True)


def test_solver_oracle_vs_golden():
    g = _golden()
    for c in g["cases"]:
        ts, traj = SO.run_trajectory(c["kind"], c["steps"], c["shift"], g["x0"])
        assert torch.equal(ts, c["timesteps"])
        assert torch.equal(torch.stack(traj), c["traj"]), (c["kind"], c["steps"])


@pytest.mark.skipif(ref_loader.find_reference() is None, reason="reference tree not mounted")
def test_solver_oracle_vs_live_reference():
    U, D = ref_loader.load_reference_solvers()
    x0 = torch.randn(1, 16, 1, 4, 6, generator=torch.Generator().manual_seed(5))
    for kind, steps, shift in [("unipc", 7, 5.0), ("dpm++", 7, 5.0), ("unipc", 30, 2.0), ("dpm++", 16, 1.0)]:
        if kind == "unipc":
            s = U.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
            s.set_timesteps(steps, device="cpu", shift=shift)
            ts = s.timesteps
        else:
            s = D.FlowDPMSolverMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
            ts, _ = D.retrieve_timesteps(s, device="cpu", sigmas=D.get_sampling_sigmas(steps, shift))
        x, ref = x0.clone(), []
        for t in ts:
            x = s.step(SO.toy_velocity(x, t), t, x, return_dict=False)[0]
            ref.append(x.clone())
        ts2, traj = SO.run_trajectory(kind, steps, shift, x0)
        assert torch.equal(ts, ts2)
        assert torch.equal(torch.stack(ref), torch.stack(traj))


def _cpu_lincomb(inputs, coeffs, like):
    ins = [t if t is not None else like for t in inputs]
    outs = []
    for row in coeffs:
        acc = torch.zeros_like(like)
        for c, t in zip(row, ins):
            if float(c) != 0.0:
                acc = acc + float(torch.tensor(float(c), dtype=torch.float32)) * t
        outs.append(acc)
    return outs


def test_product_scheduler_host_logic(monkeypatch):
    import b200dit
    from b200dit import solvers as PS
    monkeypatch.setattr(PS, "_lincomb", _cpu_lincomb)
    g = _golden()
    for c in g["cases"]:
        if c["kind"] == "unipc":
            s = PS.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
            s.set_timesteps(c["steps"], device="cpu", shift=c["shift"])
            ts = s.timesteps
        else:
            s = PS.FlowDPMSolverMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
            ts, _ = PS.retrieve_timesteps(s, device="cpu", sigmas=PS.get_sampling_sigmas(c["steps"], c["shift"]))
        assert torch.equal(ts, c["timesteps"]) and torch.equal(s.sigmas, c["sigmas"])
        x = g["x0"].clone()
        for k, t in enumerate(ts):
            x = s.step(SO.toy_velocity(x, t), t, x, return_dict=False)[0]
            assert float((x - c["traj"][k]).abs().max()) < 2e-5, (c["kind"], c["steps"], k)
    assert b200dit.FlowUniPCMultistepScheduler is PS.FlowUniPCMultistepScheduler


def test_product_scheduler_rejects_unsupported():
    from b200dit import solvers as PS
    with pytest.raises(NotImplementedError):
        PS.FlowUniPCMultistepScheduler(solver_order=3)
    with pytest.raises(NotImplementedError):
        PS.FlowDPMSolverMultistepScheduler(algorithm_type="sde-dpmsolver++")
    s = PS.FlowUniPCMultistepScheduler()
    with pytest.raises(ValueError):
        s.step(torch.zeros(1), 0, torch.zeros(1))
