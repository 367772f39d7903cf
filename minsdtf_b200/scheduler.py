"""Scheduler with the reference's interface (stable_diffusion/scheduler.py:22-318): `Scheduler(active_tcd)`,
`set_timesteps(n)`, `timesteps`, `signal_rates`, `noise_rates`, `alphas_cumprod`, `step(...)`.

Host side only computes *schedules and per-step scalars* (in float64, like the reference); the update itself runs
on the GPU inside the fused CFG + scheduler kernel, fed by `coefficients()`:

    DDIM  (scheduler.py:285, 308-312):  x0 = (x - nr_t e)/sr_t ;  x' = sr_p x0 + nr_p e      (last step: x' = x0)
      =>  ca = sr_p / sr_t ,  cb = nr_p - sr_p nr_t / sr_t                                  (last: 1/sr_t, -nr_t/sr_t)
    TCD   (scheduler.py:286-307):  s = floor((1-eta) t_prev) ; den = sqrt(a_s) x0 + sqrt(1-a_s) e ;
                                   x' = sqrt(a_p/a_s) den + sqrt(1 - a_p/a_s) z            (last step: x' = den)
      =>  ca = k sqrt(a_s)/sr_t , cb = k (sqrt(1-a_s) - sqrt(a_s) nr_t/sr_t) , cn = sqrt(1 - a_p/a_s) , k = sqrt(a_p/a_s)
"""
from __future__ import annotations

import numpy as np

from ._lib import StepCoef


class Scheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 original_inference_steps: int = 50, active_tcd: bool = True):
        self.active_tcd = active_tcd
        self.num_train_timesteps = num_train_timesteps
        self.original_inference_steps = original_inference_steps
        root = np.linspace(np.sqrt(beta_start), np.sqrt(beta_end), num_train_timesteps)
        self.alphas_cumprod = np.cumprod(1.0 - np.square(root), axis=0)
        self.signal_rates = np.sqrt(self.alphas_cumprod)
        self.noise_rates = np.sqrt(1.0 - self.alphas_cumprod)
        self.final_alpha_cumprod = 1.0
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int32)
        self._step_index = None
        self.engine = None   # set by the pipeline that owns this scheduler; `step` creates one on `device` otherwise
        self.device = 0

    # ------------------------------------------------------------------------------------------ schedules
    def set_timesteps(self, num_inference_steps: int):
        n = int(num_inference_steps)
        if self.active_tcd:
            if n > self.original_inference_steps:
                raise ValueError(f"`num_inference_steps`: {n} cannot be larger than `original_inference_steps`: "
                                 f"{self.original_inference_steps}")
            k = self.num_train_timesteps // self.original_inference_steps
            grid = (np.arange(1, self.original_inference_steps + 1) * k - 1)[::-1]
            pick = np.floor(np.linspace(0, len(grid), num=n, endpoint=False)).astype(np.int32)
            ts = grid[pick]
        else:
            ts = np.linspace(0, 1000, n, dtype=np.int32, endpoint=False)[::-1]
        self.num_inference_steps = n
        self.timesteps = np.array(ts, dtype=np.int32)
        self._step_index = None

    @property
    def step_index(self):
        return self._step_index

    def _index_of(self, timestep) -> int:
        hits = np.nonzero(self.timesteps == timestep)[0]
        if len(hits) == 0:
            raise ValueError(f"timestep {timestep} is not on the current schedule")
        return int(hits[0])

    def step_scalars(self, timestep: int, eta: float = 0.3):
        """(ca, cb, cn) of the update executed at `timestep` (position looked up on the current schedule)."""
        i = self._index_of(timestep)
        last = i == self.num_inference_steps - 1
        if i + 1 < len(self.timesteps):
            prev_t = int(self.timesteps[i + 1])
        else:
            prev_t = 0 if self.active_tcd else int(timestep)
        sr, nr = self.signal_rates[timestep], self.noise_rates[timestep]
        if self.active_tcd:
            s = int(np.floor((1.0 - eta) * prev_t))
            a_s = self.alphas_cumprod[s]
            ca, cb, cn = np.sqrt(a_s) / sr, np.sqrt(1.0 - a_s) - np.sqrt(a_s) * nr / sr, 0.0
            if eta > 0.0 and not last:
                ratio = self.alphas_cumprod[prev_t] / a_s
                k = np.sqrt(ratio)
                ca, cb, cn = k * ca, k * cb, np.sqrt(1.0 - ratio)
        elif not last:
            srp, nrp = self.signal_rates[prev_t], self.noise_rates[prev_t]
            ca, cb, cn = srp / sr, nrp - srp * nr / sr, 0.0
        else:
            ca, cb, cn = 1.0 / sr, -nr / sr, 0.0
        return float(ca), float(cb), float(cn)

    def coefficients(self, exec_timesteps, guidance: float, rescale: float, eta: float = 0.3):
        """StepCoef list for the timesteps in execution order (descending t)."""
        out = []
        for t in exec_timesteps:
            ca, cb, cn = self.step_scalars(int(t), eta)
            out.append(StepCoef(float(guidance), float(rescale), ca, cb, cn, float(self.signal_rates[t]),
                                float(self.noise_rates[t])))
        return out

    # ------------------------------------------------------------------------------------------ reference API
    def step(self, latent, timestep: int, latent_prev, eta: float = 0.3):
        """Drop-in for Scheduler.step(eps, t, latent_prev) (scheduler.py:246-315).  Runs the fused kernel."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if self.engine is None:
            from .engine import Engine
            self.engine = Engine(self.device)
        if self._step_index is None:
            self._step_index = self._index_of(timestep)
        ca, cb, cn = self.step_scalars(int(timestep), eta)
        noise = None
        if cn != 0.0:
            noise = np.random.randn(*np.shape(latent)).astype(np.float32)  # global NumPy RNG, as scheduler.py:301
        coef = StepCoef(0.0, 0.0, ca, cb, cn, 0.0, 0.0)
        out = self.engine.cfg_sched_step(None, latent, latent_prev, coef, noise=noise)
        self._step_index += 1
        return out

    def __len__(self):
        return self.num_train_timesteps


def timestep_embedding(timestep, dim=320, max_period=10000):
    """(dim,) float32 sinusoidal embedding, cos first (stable_diffusion.py:543-553).  Evaluated in float64 and
    rounded once to float32: that is what the reference's expression does under NumPy >= 2 promotion rules
    (np.log returns a float64 scalar) before Keras casts the model input to float32."""
    half = dim // 2
    freqs = np.exp(-np.log(float(max_period)) * np.arange(0, half, dtype=np.float64) / half)
    args = float(timestep) * freqs
    return np.concatenate([np.cos(args), np.sin(args)], axis=0).astype(np.float32)
