"""Seeded synthetic SD1.5 checkpoints and inputs (there are no real checkpoints offline).

Weights are produced in the *reference's checkpoint format*: PyTorch-layout fp32 tensors under the reference's key
names (minsdtf_b200.keys), so the same state dict / .safetensors file feeds this engine, the CPU oracle and — were
Keras available — the reference's own load_weights_from_file (ckpt_loader.py:2136-2193).

Distribution: matrices / filters N(0, g^2 / fan_in) with g = 1, g = 0.5 for the tensors that close
a residual branch (ResBlock conv2, attention to_out, ff.net.2, proj_out, ControlNet zero-convs — deliberately
non-zero so the ControlNet branch is exercised); norm gains 1 + 0.1 N(0,1); biases and norm offsets 0.05 N(0,1).
Every tensor has its own seed derived from (seed, key), so components can be generated independently.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch

from . import keys as K

_HALF_GAIN = ("out_layers.3.weight", "to_out.0.weight", "ff.net.2.weight", "proj_out.weight", "zero_convs",
              "middle_block_out", ".conv2.weight", "proj_attn.weight", "out_proj.weight", "mlp.fc2.weight")


def _tensor(key: str, shape, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFFFFFF)
    if len(shape) == 1:
        if key.endswith(".weight"):  # norm gain
            return 1.0 + 0.1 * torch.randn(shape, generator=g)
        return 0.05 * torch.randn(shape, generator=g)
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    gain = 0.5 if any(t in key for t in _HALF_GAIN) and key.endswith(".weight") else 1.0
    return torch.randn(shape, generator=g) * (gain / fan_in ** 0.5)


def make_state_dict(component: str, seed: int = 123456) -> dict:
    """component in {'unet','controlnet','hintnet','decoder','encoder','text_encoder'} -> {key: fp32 tensor (PyTorch layout)}."""
    return {k: _tensor(k, shp, seed) for k, shp in K.COMPONENT_KEYS[component]().items()}


def prompt_tokens(batch, length=77, seed=123461):
    """synthetic CLIP token ids (no tokenizer vocabulary offline): <start>, random word ids, <end> padding, as the
    reference tokenizer pads (clip_tokenizer.py: start 49406, end / pad 49407)."""
    rng = np.random.default_rng(seed)
    out = np.full((batch, length), 49407, np.int32)
    out[:, 0] = 49406
    for b in range(batch):
        n = int(rng.integers(3, length - 2))
        out[b, 1:1 + n] = rng.integers(0, 49406, n)
    return out


def make_vae_state_dict(seed: int = 123456) -> dict:
    sd = make_state_dict("decoder", seed)
    sd.update(make_state_dict("encoder", seed))
    return sd


def make_controlnet_state_dict(seed: int = 123456) -> dict:
    """hintnet + controlnet keys in one dict, as control_sd15_canny.pth has them."""
    sd = make_state_dict("controlnet", seed)
    sd.update(make_state_dict("hintnet", seed))
    return sd


def save_safetensors(sd: dict, path: str):
    from safetensors.torch import save_file
    save_file({k: v.contiguous() for k, v in sd.items()}, path)


# ----------------------------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------------------------
def latents(batch, h, w, seed=123456):
    return np.random.default_rng(seed).standard_normal((batch, h, w, 4)).astype(np.float32)


def context(batch, tokens=77, seed=123457):
    return np.random.default_rng(seed).standard_normal((batch, tokens, 768)).astype(np.float32)


def uncond_context(batch, tokens=77, seed=123458):
    c = np.random.default_rng(seed).standard_normal((1, tokens, 768)).astype(np.float32)
    return np.repeat(c, batch, axis=0)


def edge_map(h=512, w=512, seed=123459):
    """binary uint8 (h,w,3) control image in {0,255}."""
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w), np.uint8)
    for _ in range(24):
        y0, x0 = rng.integers(0, h), rng.integers(0, w)
        ln = rng.integers(h // 8, h // 2)
        if rng.random() < 0.5:
            img[y0:y0 + 2, x0:x0 + ln] = 255
        else:
            img[y0:y0 + ln, x0:x0 + 2] = 255
    return np.repeat(img[..., None], 3, axis=-1)


def smooth_image(h=512, w=512, seed=123460):
    """smooth uint8 RGB source image for img2img / inpaint."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
    img = np.zeros((h, w, 3), np.float32)
    for c in range(3):
        for _ in range(4):
            fy, fx, ph = rng.uniform(0.5, 3.0), rng.uniform(0.5, 3.0), rng.uniform(0, 6.28)
            img[..., c] += np.sin(6.28 * (fy * yy + fx * xx) + ph)
    img = (img - img.min()) / (img.max() - img.min())
    return (img * 255).astype(np.uint8)


def center_mask(h=512, w=512):
    m = np.zeros((h, w), np.uint8)
    m[h // 4: 3 * h // 4, w // 4: 3 * w // 4] = 255
    return m
