"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's flow-matching multistep solvers.

Not part of the product path (only tests/ and bench.py's CPU legs may import it).  Restates, for the
configuration the reference callers use (text2video.py:204-219, omnihuman_wan_t2v.py:171-176:
`prediction_type="flow_prediction"`, `solver_order=2`, `predict_x0=True`, `solver_type="bh2"` /
`algorithm_type="dpmsolver++"` + `"midpoint"`, `lower_order_final=True`, `final_sigmas_type="zero"`,
no thresholding, no dynamic shifting):

  * UniPC      seaweed_apt/wan/utils/fm_solvers_unipc.py:160-229 (set_timesteps), :276-347 (x0 conversion),
               :350-483 (UniP predictor), :486-626 (UniC corrector), :655-739 (step)
  * DPM++      seaweed_apt/wan/utils/fm_solvers.py:22-26 (get_sampling_sigmas), :226-290, :341-395,
               :415-483 (first order), :486-593 (second order, midpoint), :706-797 (step)

Parity status: pinned against the UNMODIFIED reference classes executed in the authoring container
(oracle/ref_loader.load_reference_solvers, tests/test_cpu_solvers.py::test_solver_oracle_vs_live_reference)
and against the trajectories they produced (tests/golden/solver_traj.pt, made by oracle/make_golden.py).

All scalar arithmetic is done on fp32 torch scalars exactly as the reference does (its sigma table is a
float32 tensor; lambda = log(alpha) - log(sigma) in fp32, expm1 / exp in fp32).
"""
import numpy as np
import torch


def _sigma_table(num_train_timesteps, shift):
    """fm_solvers_unipc.py:106-117 / fm_solvers.py:175-186: training sigma table (constructor)."""
    alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
    sigmas = torch.from_numpy(1.0 - alphas).to(torch.float32)
    return shift * sigmas / (1 + (shift - 1) * sigmas)


def get_sampling_sigmas(sampling_steps, shift):
    """fm_solvers.py:22-26."""
    sigma = np.linspace(1, 0, sampling_steps + 1)[:sampling_steps]
    return shift * sigma / (1 + (shift - 1) * sigma)


class _Base:
    def __init__(self, num_train_timesteps=1000, solver_order=2, shift=1.0):
        self.n_train, self.order, self.cfg_shift = num_train_timesteps, solver_order, shift
        table = _sigma_table(num_train_timesteps, shift)
        self.sigma_min, self.sigma_max = table[-1].item(), table[0].item()
        self.sigmas = table
        self.timesteps = table * num_train_timesteps
        self.step_index = None

    def set_timesteps(self, num_inference_steps=None, sigmas=None, shift=None):
        """fm_solvers_unipc.py:160-229 == fm_solvers.py:226-290: linspace(sigma_max, sigma_min) (or caller sigmas),
        shift, append the final 0, timesteps = (sigma * 1000) truncated to int64."""
        if sigmas is None:
            sigmas = np.linspace(self.sigma_max, self.sigma_min, num_inference_steps + 1).copy()[:-1]
        if shift is None:
            shift = self.cfg_shift
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        timesteps = sigmas * self.n_train
        self.sigmas = torch.from_numpy(np.concatenate([sigmas, [0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(timesteps).to(torch.int64)
        self.model_outputs = [None] * self.order
        self.lower_order_nums = 0
        self.last_sample = None
        self.step_index = None
        return self.timesteps

    def _init_step_index(self, timestep):
        """index_for_timestep (fm_solvers_unipc.py:628-640): second match if the timestep occurs twice."""
        idx = (self.timesteps == int(timestep)).nonzero()
        self.step_index = idx[1 if len(idx) > 1 else 0].item()

    @staticmethod
    def _lam(sigma):
        return torch.log(1 - sigma) - torch.log(sigma)


class UniPCOracle(_Base):
    def _bh(self, h, rks, order):
        """Shared by predictor and corrector (fm_solvers_unipc.py:424-446 / :566-588)."""
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        B_h = torch.expm1(hh)
        fact = 1
        R, b = [], []
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        return h_phi_1, B_h, torch.stack(R), torch.tensor(b)

    def _predict(self, x, order):
        i = self.step_index
        m0 = self.model_outputs[-1]
        sig_t, sig_s0 = self.sigmas[i + 1], self.sigmas[i]
        al_t = 1 - sig_t
        lam_s0 = self._lam(sig_s0)
        h = self._lam(sig_t) - lam_s0
        rks, D1s = [], []
        for k in range(1, order):
            mi = self.model_outputs[-(k + 1)]
            rk = (self._lam(self.sigmas[i - k]) - lam_s0) / h
            rks.append(rk)
            D1s.append((mi - m0) / rk)
        rks.append(1.0)
        h_phi_1, B_h, R, b = self._bh(h, torch.tensor(rks), order)
        x_t = sig_t / sig_s0 * x - al_t * h_phi_1 * m0
        if D1s:
            rhos = torch.tensor([0.5]) if order == 2 else torch.linalg.solve(R[:-1, :-1], b[:-1])
            x_t = x_t - al_t * B_h * sum(r * d for r, d in zip(rhos, D1s))
        return x_t

    def _correct(self, model_t, last_sample, order):
        i = self.step_index
        m0 = self.model_outputs[-1]
        sig_t, sig_s0 = self.sigmas[i], self.sigmas[i - 1]
        al_t = 1 - sig_t
        lam_s0 = self._lam(sig_s0)
        h = self._lam(sig_t) - lam_s0
        rks, D1s = [], []
        for k in range(1, order):
            mi = self.model_outputs[-(k + 1)]
            rk = (self._lam(self.sigmas[i - (k + 1)]) - lam_s0) / h
            rks.append(rk)
            D1s.append((mi - m0) / rk)
        rks.append(1.0)
        h_phi_1, B_h, R, b = self._bh(h, torch.tensor(rks), order)
        rhos = torch.tensor([0.5]) if order == 1 else torch.linalg.solve(R, b)
        x_t = sig_t / sig_s0 * last_sample - al_t * h_phi_1 * m0
        corr = sum(r * d for r, d in zip(rhos[:-1], D1s)) if D1s else 0
        return x_t - al_t * B_h * (corr + rhos[-1] * (model_t - m0))

    def step(self, model_output, timestep, sample):
        """fm_solvers_unipc.py:655-739."""
        if self.step_index is None:
            self._init_step_index(timestep)
        i = self.step_index
        use_corrector = i > 0 and self.last_sample is not None
        m_t = sample - self.sigmas[i] * model_output                       # :318-320 (flow prediction -> x0)
        if use_corrector:
            sample = self._correct(m_t, self.last_sample, self.this_order)
        self.model_outputs = self.model_outputs[1:] + [m_t]
        this_order = min(self.order, len(self.timesteps) - i)               # lower_order_final
        self.this_order = min(this_order, self.lower_order_nums + 1)
        self.last_sample = sample
        prev = self._predict(sample, self.this_order)
        if self.lower_order_nums < self.order:
            self.lower_order_nums += 1
        self.step_index += 1
        return prev


class DPMppOracle(_Base):
    def step(self, model_output, timestep, sample):
        """fm_solvers.py:706-797 with algorithm dpmsolver++, midpoint, order 2."""
        if self.step_index is None:
            self._init_step_index(timestep)
        i, n = self.step_index, len(self.timesteps)
        lower_order_final = i == n - 1                                      # final_sigmas_type == "zero"
        m_t = sample - self.sigmas[i] * model_output
        self.model_outputs = self.model_outputs[1:] + [m_t]
        sig_t, sig_s0 = self.sigmas[i + 1], self.sigmas[i]
        al_t = 1 - sig_t
        lam_t, lam_s0 = self._lam(sig_t), self._lam(sig_s0)
        h = lam_t - lam_s0
        if self.order == 1 or self.lower_order_nums < 1 or lower_order_final:
            prev = (sig_t / sig_s0) * sample - (al_t * (torch.exp(-h) - 1.0)) * m_t
        else:
            m1 = self.model_outputs[-2]
            h_0 = lam_s0 - self._lam(self.sigmas[i - 1])
            r0 = h_0 / h
            D1 = (1.0 / r0) * (m_t - m1)
            prev = ((sig_t / sig_s0) * sample - (al_t * (torch.exp(-h) - 1.0)) * m_t
                    - 0.5 * (al_t * (torch.exp(-h) - 1.0)) * D1)
        if self.lower_order_nums < self.order:
            self.lower_order_nums += 1
        self.step_index += 1
        return prev


def toy_velocity(x, t):
    """Deterministic stand-in for the DiT inside solver tests: depends on the state and the timestep."""
    return torch.tanh(0.5 * x + float(t) / 1000.0) - 0.25 * x


def run_trajectory(kind, steps, shift, x0, record=True):
    """kind 'unipc': text2video.py:204-211; 'dpm++': text2video.py:212-221 (sigmas from get_sampling_sigmas)."""
    if kind == "unipc":
        s = UniPCOracle(shift=1.0)
        ts = s.set_timesteps(steps, shift=shift)
    else:
        s = DPMppOracle(shift=1.0)
        ts = s.set_timesteps(sigmas=get_sampling_sigmas(steps, shift))
    x, traj = x0.clone(), []
    for t in ts:
        x = s.step(toy_velocity(x, t), t, x)
        if record:
            traj.append(x.clone())
    return ts, traj if record else x
