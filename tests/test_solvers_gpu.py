"""Scheduler parity on the GPU: product schedulers (host coefficients + b200_solver_lincomb through the C ABI)
against the golden trajectories of the unmodified reference schedulers (tests/golden/solver_traj.pt).
Tolerance: max-abs <= 2e-5 on O(1) fp32 latents (one fused linear combination per step instead of the
reference's chain of rounded elementwise ops)."""
import os

import pytest
import torch

from conftest import GOLDEN
from oracle import solver_oracle as SO

pytestmark = pytest.mark.gpu


def test_schedulers_vs_golden():
    import b200dit
    g = torch.load(os.path.join(GOLDEN, "solver_traj.pt"), map_location="cpu", weights_only=True)
    for c in g["cases"]:
        if c["kind"] == "unipc":
            s = b200dit.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
            s.set_timesteps(c["steps"], device="cuda", shift=c["shift"])
            ts = s.timesteps
        else:
            s = b200dit.FlowDPMSolverMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
            ts, _ = b200dit.retrieve_timesteps(s, device="cuda", sigmas=b200dit.get_sampling_sigmas(c["steps"], c["shift"]))
        assert torch.equal(ts.cpu(), c["timesteps"])
        x = g["x0"].cuda()
        n0 = b200dit.kernel_launches()
        for k, t in enumerate(ts):
            v = SO.toy_velocity(x.cpu(), int(t)).cuda()
            x = s.step(v, t, x, return_dict=False)[0]
            assert float((x.cpu() - c["traj"][k]).abs().max()) < 2e-5, (c["kind"], c["steps"], k)
        assert b200dit.kernel_launches() - n0 == len(ts)          # one fused launch per scheduler step


def test_lincomb_odd_sizes_and_aliasing():
    import ctypes as C
    import b200dit
    from b200dit._lib import check, lib, ptr_array
    for n in (4, 1000, 1003, 16 * 21 * 60 * 104):
        a, b = torch.randn(n, device="cuda"), torch.randn(n, device="cuda")
        ref0, ref1 = 2.0 * a - 0.5 * b, a + b
        coeff = (C.c_float * 4)(2.0, -0.5, 1.0, 1.0)
        out1 = torch.empty_like(a)
        check(lib().b200_solver_lincomb(2, ptr_array([a.data_ptr(), b.data_ptr()]), 2,
                                        ptr_array([a.data_ptr(), out1.data_ptr()]), coeff, n,
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)))   # out0 aliases in0
        assert torch.allclose(a, ref0, atol=1e-6) and torch.allclose(out1, ref1, atol=1e-6)
