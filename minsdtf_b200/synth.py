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


def make_bpe_vocab(path, n_merges=49152 - 256 - 2, seed=123462):
    """Synthetic stand-in for `bpe_simple_vocab_16e6.txt.gz` (a download in the reference, clip_tokenizer.py:79-82):
    a header line and `n_merges` seeded, well-formed merge rules over the CLIP byte alphabet, so that the vocabulary
    has the published size (49408 ids, <|startoftext|> = 49406, <|endoftext|> = 49407).  The first rules join
    lower-case letters, so ordinary prompts really exercise multi-level merges."""
    import gzip
    from .bpe import byte_symbols
    rng = np.random.default_rng(seed)
    sym = list(byte_symbols())
    letters = [c for c in "abcdefghijklmnopqrstuvwxyz"]
    inner = list(letters)                      # tokens that can start a merge (no end-of-word mark)
    final = [c + "</w>" for c in letters]      # tokens carrying the end-of-word mark
    seen = set(sym) | {s + "</w>" for s in sym}
    rules = []
    draws = iter(())
    while len(rules) < n_merges:
        nxt = next(draws, None)
        if nxt is None:  # drawn in bulk: one generator call per rule would dominate the run time
            draws = iter(rng.random((65536, 4)).tolist())
            continue
        r0, r1, r2, r3 = nxt
        if len(rules) > 30000 and r0 < 0.2:  # some rules over the rest of the alphabet (digits, punctuation)
            a = sym[int(r1 * 256)]
        else:
            a = inner[int(r1 * len(inner))]
        pool = final if r2 < 0.35 else inner
        b = pool[int(r3 * len(pool))]
        tok = a + b
        if tok in seen or len(tok) > 14:
            continue
        seen.add(tok)
        rules.append(f"{a} {b}")
        (final if tok.endswith("</w>") else inner).append(tok)
    with gzip.open(path, "wb") as f:
        f.write(("#version: synthetic\n" + "\n".join(rules) + "\n").encode("utf-8"))
    return path


def make_lora_state_dict(rank=4, alpha=2.0, every=7, seed=123463) -> dict:
    """Synthetic kohya-format LoRA (`<module>.alpha`, `.lora_down.weight`, `.lora_up.weight`) over every `every`-th
    adaptable UNet module — linears, 1x1 and 3x3 convolutions, time projections, shortcuts, resamplers — and the
    attention / MLP linears of some text-encoder layers.  Module names are the flattened diffusers names LoRA files
    carry (ckpt_loader.py:2231-2272)."""
    sd = {}
    shapes, alias = K.unet_keys(), K.unet_alias_map()
    skip = ("time_embedding.", "conv_in.", "conv_out.")
    mods = [(a[:-len(".weight")], shapes[k]) for k, a in alias.items()
            if a.endswith(".weight") and len(shapes[k]) >= 2 and not a.startswith(skip)]
    picked = mods[::every] + [m for m in mods if "samplers" in m[0] or "conv_shortcut" in m[0]][:4]
    for name, shp in picked:
        flat = "lora_unet_" + name.replace(".", "_")
        out_c, in_c = shp[0], shp[1]
        if len(shp) == 4 and shp[2] == 3:
            down, up = (rank, in_c, 3, 3), (out_c, rank, 1, 1)
        elif len(shp) == 4:
            down, up = (rank, in_c, 1, 1), (out_c, rank, 1, 1)
        else:
            down, up = (rank, in_c), (out_c, rank)
        sd[flat + ".lora_down.weight"] = _tensor(flat + ".down", down, seed)
        sd[flat + ".lora_up.weight"] = _tensor(flat + ".up", up, seed) * 0.5
        sd[flat + ".alpha"] = torch.tensor(alpha)
    for layer in (0, 5, 11):
        for leaf, (o, i) in {"mlp_fc1": (3072, 768), "mlp_fc2": (768, 3072), "self_attn_q_proj": (768, 768),
                             "self_attn_out_proj": (768, 768)}.items():
            flat = f"lora_te_text_model_encoder_layers_{layer}_{leaf}"
            sd[flat + ".lora_down.weight"] = _tensor(flat + ".down", (rank, i), seed)
            sd[flat + ".lora_up.weight"] = _tensor(flat + ".up", (o, rank), seed) * 0.5
            sd[flat + ".alpha"] = torch.tensor(alpha)
    return sd
