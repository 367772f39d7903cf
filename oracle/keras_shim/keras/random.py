"""keras.random.normal: TF's stateless Philox stream is not reproducible here, so callers of the harness always pass
`diffusion_noise`; a NumPy generator stands in to keep the call valid."""
import numpy as np


def normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None):
    return (mean + stddev * np.random.default_rng(seed).standard_normal(shape)).astype(np.float32)
