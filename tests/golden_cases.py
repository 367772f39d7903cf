"""Inputs of the reference-run fixtures (tests/golden/reference_run.npz), shared by the script that produces them
(tools/make_golden_ref.py: the reference's own code on the Keras shim) and the tests that consume them."""
import numpy as np

from minsdtf_b200 import synth

H = 128            # image size of the fixtures -> 16 x 16 latent
h = H // 8
PROMPT = "a photo of an (astronaut:1.3) riding a [horse] on ((mars)), \\(literal\\) 4k"
NEGATIVE = "blurry, (low quality:1.4)"
LONG_PROMPT = ", ".join(["a (very:1.2) detailed matte painting of a castle on a hill at sunset", "volumetric light",
                         "[fog]", "trending on artstation", "ultra (wide:0.8) angle", "octane render"])  # 2 windows of 75 tokens
TRUNCATED_PROMPT = ", ".join([LONG_PROMPT] * 3)  # > 300 tokens: cut at 4 windows


def noise(batch=1, seed=123456):
    return synth.latents(batch, h, h, seed=seed)


def contexts(batch=1):
    return synth.context(batch), synth.uncond_context(batch)


def source_image():
    return synth.smooth_image(H, H)


def mask():
    return synth.center_mask(H, H)


def edges():
    return synth.edge_map(H, H)


def ti_embedding(n=3, seed=123470):
    return (0.02 * np.random.default_rng(seed).standard_normal((n, 768))).astype(np.float32)
