"""B200 host side of the reference's flow-matching multistep schedulers.

Drop-in for `FlowUniPCMultistepScheduler` (seaweed_apt/wan/utils/fm_solvers_unipc.py:20, the default of
`WanT2V.generate`, text2video.py:204-211) and `FlowDPMSolverMultistepScheduler`
(seaweed_apt/wan/utils/fm_solvers.py:69; text2video.py:212-221, omnihuman_wan_t2v.py:171-176) in the
configuration those callers use: flow prediction, x0 parameterisation, `solver_order` 1 or 2, `bh2` /
`dpmsolver++` + `midpoint`, `lower_order_final`, final sigma 0, no thresholding, no dynamic shifting.
Same constructor keywords, `set_timesteps`, `.timesteps`, `.sigmas`, `.step(...)` and
`retrieve_timesteps` / `get_sampling_sigmas`; diffusers is not needed.

Every tensor update of a step (x0 conversion, UniC corrector, UniP / DPM++ predictor) is a linear
combination of {model_output, sample, last_sample, the two stored x0 predictions}.  The scalar
coefficients are computed here on fp32 CPU scalars with the reference's formulas and the whole step
is ONE launch of `b200_solver_lincomb` (csrc/elementwise.cu) -- the reference issues about twenty
elementwise kernels plus a device->host `.item()` per step.
"""
import ctypes as C

import numpy as np
import torch

from ._lib import check, lib, ptr_array


def get_sampling_sigmas(sampling_steps, shift):
    """fm_solvers.py:22-26."""
    sigma = np.linspace(1, 0, sampling_steps + 1)[:sampling_steps]
    return shift * sigma / (1 + (shift - 1) * sigma)


def retrieve_timesteps(scheduler, num_inference_steps=None, device=None, timesteps=None, sigmas=None, **kwargs):
    """fm_solvers.py:29-66 (custom `timesteps` schedules are not supported by these schedulers either)."""
    if timesteps is not None and sigmas is not None:
        raise ValueError("Only one of `timesteps` or `sigmas` can be passed. Please choose one to set custom values")
    if timesteps is not None:
        raise ValueError(f"The current scheduler class {scheduler.__class__}'s `set_timesteps` does not support custom"
                         f" timestep schedules. Please check whether you are using the correct scheduler.")
    if sigmas is not None:
        scheduler.set_timesteps(sigmas=sigmas, device=device, **kwargs)
        return scheduler.timesteps, len(scheduler.timesteps)
    scheduler.set_timesteps(num_inference_steps, device=device, **kwargs)
    return scheduler.timesteps, num_inference_steps


class SchedulerOutput:
    def __init__(self, prev_sample):
        self.prev_sample = prev_sample


def _lincomb(inputs, coeffs, like):
    """outputs[j] = sum_i coeffs[j][i] * inputs[i]; inputs are fp32 CUDA tensors of one shape (None = unused)."""
    ref = like
    ins = [t if t is not None else ref for t in inputs]
    outs = [torch.empty_like(ref) for _ in coeffs]
    flat = (C.c_float * (len(coeffs) * len(ins)))(*[float(c) for row in coeffs for c in row])
    with torch.cuda.device(ref.device):
        check(lib().b200_solver_lincomb(len(ins), ptr_array([t.data_ptr() for t in ins]), len(outs),
                                        ptr_array([t.data_ptr() for t in outs]), flat, ref.numel(),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return outs


class _FlowScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, solver_order=2, prediction_type="flow_prediction", shift=1.0,
                 use_dynamic_shifting=False, thresholding=False, lower_order_final=True, final_sigmas_type="zero",
                 **unused):
        if prediction_type != "flow_prediction":
            raise ValueError(f"prediction_type given as {prediction_type} must be `flow_prediction`")
        if use_dynamic_shifting or thresholding or final_sigmas_type != "zero" or not lower_order_final:
            raise NotImplementedError("only the configuration used by the reference callers is implemented "
                                      "(no dynamic shifting / thresholding, final sigma 0, lower_order_final)")
        if solver_order not in (1, 2):
            raise NotImplementedError("solver_order must be 1 or 2")
        self.config = type("Config", (), dict(num_train_timesteps=num_train_timesteps, solver_order=solver_order,
                                              prediction_type=prediction_type, shift=shift,
                                              use_dynamic_shifting=False, thresholding=False,
                                              lower_order_final=True, final_sigmas_type="zero"))()
        alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
        sig = torch.from_numpy(1.0 - alphas).to(torch.float32)
        sig = shift * sig / (1 + (shift - 1) * sig)
        self.sigmas = sig
        self.timesteps = sig * num_train_timesteps
        self.sigma_min, self.sigma_max = sig[-1].item(), sig[0].item()
        self.num_inference_steps = None
        self._reset()

    def _reset(self):
        self.model_outputs = [None] * self.config.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self._step_index = None
        self._begin_index = None

    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_begin_index(self, begin_index=0):
        self._begin_index = begin_index

    def scale_model_input(self, sample, *args, **kwargs):
        return sample

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None, shift=None):
        """fm_solvers_unipc.py:160-229 / fm_solvers.py:226-290."""
        if sigmas is None:
            sigmas = np.linspace(self.sigma_max, self.sigma_min, num_inference_steps + 1).copy()[:-1]
        if shift is None:
            shift = self.config.shift
        sigmas = shift * np.asarray(sigmas) / (1 + (shift - 1) * np.asarray(sigmas))
        timesteps = sigmas * self.config.num_train_timesteps
        self.sigmas = torch.from_numpy(np.concatenate([sigmas, [0]]).astype(np.float32))   # host, like the reference
        self._timesteps_host = torch.from_numpy(timesteps).to(torch.int64)
        self.timesteps = self._timesteps_host.to(device=device) if device is not None else self._timesteps_host
        self.num_inference_steps = len(timesteps)
        self._reset()

    def _init_step_index(self, timestep):
        if self._begin_index is not None:
            self._step_index = self._begin_index
            return
        t = int(timestep.item()) if torch.is_tensor(timestep) else int(timestep)   # once per trajectory
        idx = (self._timesteps_host == t).nonzero()
        self._step_index = idx[1 if len(idx) > 1 else 0].item()

    @staticmethod
    def _lam(sigma):
        return torch.log(1 - sigma) - torch.log(sigma)

    @staticmethod
    def _prep(t):
        return t.detach().to(torch.float32).contiguous()


class FlowUniPCMultistepScheduler(_FlowScheduler):
    """UniPC (B(h) = expm1, bh2) predictor-corrector; fm_solvers_unipc.py:350-739."""

    def __init__(self, *args, predict_x0=True, solver_type="bh2", disable_corrector=(), solver_p=None, **kw):
        if not predict_x0 or solver_p is not None:
            raise NotImplementedError("predict_x0=True without solver_p is the implemented configuration")
        if solver_type in ("midpoint", "heun", "logrho"):
            solver_type = "bh2"
        if solver_type != "bh2":
            raise NotImplementedError("solver_type bh2 only")
        super().__init__(*args, **kw)
        self.disable_corrector = list(disable_corrector)
        self.this_order = None

    def _bh(self, h, rks, order):
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        B_h = torch.expm1(hh)
        fact, R, b = 1, [], []
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        return h_phi_1, B_h, torch.stack(R), torch.tensor(b)

    def step(self, model_output, timestep, sample, return_dict=True, generator=None):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating "
                             "the scheduler")
        if self._step_index is None:
            self._init_step_index(timestep)
        i = self._step_index
        out_dtype = sample.dtype
        v, x = self._prep(model_output), self._prep(sample)
        mA, mB = self.model_outputs[-1], (self.model_outputs[-2] if self.config.solver_order > 1 else None)
        sig = self.sigmas
        # basis: [v, sample, last_sample, mA, mB]; every quantity below is a coefficient vector over it
        V, X, L, A, B = (np.eye(5)[k] for k in range(5))
        m_t = X - float(sig[i]) * V                                               # :318-320
        use_corrector = i > 0 and (i - 1) not in self.disable_corrector and self.last_sample is not None
        x_c = X
        if use_corrector:                                                          # :486-626, order = previous this_order
            order = self.this_order
            sig_t, sig_s0 = sig[i], sig[i - 1]
            lam_s0 = self._lam(sig_s0)
            h = self._lam(sig_t) - lam_s0
            rks, D1 = [], []
            if order == 2:
                rk = (self._lam(sig[i - 2]) - lam_s0) / h
                rks.append(rk)
                D1.append((B - A) / float(rk))
            rks.append(1.0)
            h_phi_1, B_h, R, b = self._bh(h, torch.tensor(rks), order)
            rhos = torch.tensor([0.5]) if order == 1 else torch.linalg.solve(R, b)
            e = -float((1 - sig_t) * B_h)
            x_c = float(sig_t / sig_s0) * L - float((1 - sig_t) * h_phi_1) * A
            for r, d in zip(rhos[:-1], D1):
                x_c = x_c + e * float(r) * d
            x_c = x_c + e * float(rhos[-1]) * (m_t - A)
        this_order = min(self.config.solver_order, len(self._timesteps_host) - i)    # lower_order_final
        self.this_order = min(this_order, self.lower_order_nums + 1)
        # predictor (:350-483) from x_c with history [.., mA, m_t]
        sig_t, sig_s0 = sig[i + 1], sig[i]
        lam_s0 = self._lam(sig_s0)
        h = self._lam(sig_t) - lam_s0
        rks = [1.0]
        if self.this_order == 2:
            rk = (self._lam(sig[i - 1]) - lam_s0) / h
            rks = [rk, 1.0]
        h_phi_1, B_h, _, _ = self._bh(h, torch.tensor(rks), self.this_order)
        x_n = float(sig_t / sig_s0) * x_c - float((1 - sig_t) * h_phi_1) * m_t
        if self.this_order == 2:
            x_n = x_n - float((1 - sig_t) * B_h) * 0.5 * (A - m_t) / float(rk)
        coeffs = [m_t, x_c, x_n] if use_corrector else [m_t, x_n]
        outs = _lincomb([v, x, self.last_sample, mA, mB], coeffs, x)
        new_m, prev = outs[0], outs[-1]
        self.last_sample = outs[1] if use_corrector else x
        self.model_outputs = self.model_outputs[1:] + [new_m]
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        prev = prev.to(out_dtype)
        return SchedulerOutput(prev) if return_dict else (prev,)


class FlowDPMSolverMultistepScheduler(_FlowScheduler):
    """DPM-Solver++ (multistep, midpoint); fm_solvers.py:415-593, 706-797."""

    def __init__(self, *args, algorithm_type="dpmsolver++", solver_type="midpoint", euler_at_final=False, **kw):
        if algorithm_type == "deis":
            algorithm_type = "dpmsolver++"
        if solver_type in ("logrho", "bh1", "bh2"):
            solver_type = "midpoint"
        if algorithm_type != "dpmsolver++" or solver_type != "midpoint":
            raise NotImplementedError("dpmsolver++ with the midpoint rule is the implemented configuration")
        super().__init__(*args, **kw)

    def step(self, model_output, timestep, sample, generator=None, variance_noise=None, return_dict=True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating "
                             "the scheduler")
        if self._step_index is None:
            self._init_step_index(timestep)
        i, n = self._step_index, len(self._timesteps_host)
        out_dtype = model_output.dtype
        v, x = self._prep(model_output), self._prep(sample)
        mA = self.model_outputs[-1]
        sig = self.sigmas
        V, X, A = (np.eye(3)[k] for k in range(3))                                  # basis [v, sample, mA]
        m_t = X - float(sig[i]) * V                                                  # :380-384
        sig_t, sig_s0 = sig[i + 1], sig[i]
        lam_s0 = self._lam(sig_s0)
        h = self._lam(sig_t) - lam_s0
        g = float((1 - sig_t) * (torch.exp(-h) - 1.0))
        x_n = float(sig_t / sig_s0) * X - g * m_t                                   # first order (:465-468)
        second = not (self.config.solver_order == 1 or self.lower_order_nums < 1 or i == n - 1)
        if second:                                                                   # :548-553
            r0 = (lam_s0 - self._lam(sig[i - 1])) / h
            x_n = x_n - 0.5 * g * (1.0 / float(r0)) * (m_t - A)
        new_m, prev = _lincomb([v, x, mA], [m_t, x_n], x)
        self.model_outputs = self.model_outputs[1:] + [new_m]
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        prev = prev.to(out_dtype)
        return SchedulerOutput(prev) if return_dict else (prev,)
