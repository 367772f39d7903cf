"""TEST INFRASTRUCTURE — CPU oracle, never imported by the product path.

NumPy (fp64) restatement of the reference scheduler, stable_diffusion/scheduler.py:
  * schedule constants              scheduler.py:52-55
  * DDIM timesteps                  scheduler.py:238-241
  * TCD timesteps (default branch)  scheduler.py:136-151, 233-237
  * step()                          scheduler.py:246-315

Parity PINNED: tests/test_scheduler_oracle.py checks this file against tests/golden/scheduler.npz, which
tools/make_golden.py produced by importing and running the real scheduler.py.
"""
import numpy as np


class OracleScheduler:
    def __init__(self, active_tcd=False, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 original_inference_steps=50):
        self.active_tcd = active_tcd
        self.T = num_train_timesteps
        self.original_inference_steps = original_inference_steps
        betas = np.linspace(np.sqrt(beta_start), np.sqrt(beta_end), num_train_timesteps) ** 2
        self.alphas_cumprod = np.cumprod(1.0 - betas, axis=0)  # scheduler.py:52-53
        self.signal_rates = np.sqrt(self.alphas_cumprod)       # :54
        self.noise_rates = np.sqrt(1.0 - self.alphas_cumprod)  # :55
        self.timesteps = np.arange(0, num_train_timesteps)[::-1].astype(np.int32)
        self.num_inference_steps = None
        self._idx = None

    def set_timesteps(self, n):
        if self.active_tcd:
            k = self.T // self.original_inference_steps               # :147
            origin = np.arange(1, self.original_inference_steps + 1) * k - 1  # :149 (strength 1)
            if n > self.original_inference_steps:
                raise ValueError("num_inference_steps larger than original_inference_steps")
            origin = origin[::-1]
            idx = np.floor(np.linspace(0, len(origin), num=n, endpoint=False)).astype(np.int32)  # :234-236
            ts = origin[idx]
        else:
            ts = np.linspace(0, 1000, n, dtype=np.int32, endpoint=False)[::-1]  # :240-241
        self.num_inference_steps = n
        self.timesteps = np.array(ts, dtype=np.int32)
        self._idx = None

    def step(self, eps, timestep, latent_prev, eta=0.3, noise=None):
        """`noise`: optional explicit N(0,1) draw for the TCD branch (the reference uses the global NumPy RNG,
        scheduler.py:301; pass None to reproduce that)."""
        if self._idx is None:
            self._idx = int(np.nonzero(self.timesteps == timestep)[0][0])  # :69-85
        i = self._idx
        last = i == self.num_inference_steps - 1
        if i + 1 < len(self.timesteps):
            prev_t = int(self.timesteps[i + 1])
        else:
            prev_t = 0 if self.active_tcd else int(timestep)          # :273-277
        sr, nr = self.signal_rates[timestep], self.noise_rates[timestep]
        x0 = (latent_prev - nr * eps) / sr                             # :285
        if self.active_tcd:
            s = int(np.floor((1.0 - eta) * prev_t))                    # :287
            a_s = self.alphas_cumprod[s]
            den = np.sqrt(a_s) * x0 + np.sqrt(1.0 - a_s) * eps         # :292
            if eta > 0.0 and not last:
                a_to = self.alphas_cumprod[prev_t]
                if noise is None:
                    noise = np.random.randn(*eps.shape).astype(np.float32)  # :301
                out = np.sqrt(a_to / a_s) * den + np.sqrt(1.0 - a_to / a_s) * noise  # :302-303
            else:
                out = den
        else:
            if not last:
                out = self.signal_rates[prev_t] * x0 + self.noise_rates[prev_t] * eps  # :309-310
            else:
                out = x0                                                # :312
        self._idx += 1
        return out


def timestep_embedding(timestep, batch, dim=320, max_period=10000):
    """stable_diffusion.py:543-553 (cos first, fp32)."""
    half = dim // 2
    freqs = np.exp(-np.log(max_period) * np.arange(0, half, dtype=np.float32) / half)
    args = np.asarray([timestep], dtype=np.float32) * freqs
    emb = np.concatenate([np.cos(args), np.sin(args)], axis=0).reshape(1, -1)
    return np.repeat(emb, batch, axis=0)


def rescale_noise_cfg(noise_cfg, noise_text, guidance_rescale, epsilon=1e-5):
    """stable_diffusion.py:304-315."""
    ax = tuple(range(1, noise_text.ndim))
    std_text = np.std(noise_text, axis=ax, keepdims=True)
    std_cfg = np.std(noise_cfg, axis=ax, keepdims=True) + epsilon
    resc = noise_cfg * (std_text / std_cfg)
    return guidance_rescale * resc + (1.0 - guidance_rescale) * noise_cfg


def cfg_combine(eps_u, eps_c, scale, guidance_rescale):
    """stable_diffusion.py:458-461."""
    e = eps_u + scale * (eps_c - eps_u)
    if guidance_rescale > 0.0:
        e = rescale_noise_cfg(e, eps_c, guidance_rescale)
    return e
