"""LoRA merge at load time — behaviour of the reference's `load_weights_from_lora`
(`stable_diffusion/ckpt_loader.py:2196-2276`) and of the `w = w + lora_w` branch of its loader (:2169-2180).

A kohya-format LoRA file holds, per adapted module `M`, `M.alpha`, `M.lora_down.weight` and `M.lora_up.weight`.  The
weight delta is `(alpha / rank) * up @ down` (rank = columns of `up`): a matrix product for linears and 1x1 convolutions,
and for 3x3 convolutions the composition of the 3x3 `down` filter with the 1x1 `up` filter.  Module names are flattened
with underscores (`lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_q`); they are turned back into the
diffusers-style parameter names the reference's `UNET_KEY_MAPPING` uses as aliases, or into `text_model.encoder.layers.N.*`
names for the text encoder.  Deltas are added to the checkpoint tensors in PyTorch layout before they are handed to the
engine (`sdtf_load_tensor`), so the packed bf16 weights are those of the merged model.
"""
from __future__ import annotations

import re

import numpy as np

_TE_PREFIX = "lora_te_text_model_encoder_layers_"
_UNET_PREFIX = "lora_unet_"
_TE_LEAVES = {"mlp_fc1": "mlp.fc1", "mlp_fc2": "mlp.fc2", "self_attn_q_proj": "self_attn.q_proj", "self_attn_k_proj": "self_attn.k_proj",
              "self_attn_v_proj": "self_attn.v_proj", "self_attn_out_proj": "self_attn.out_proj"}
# leaf parameter of a UNet module, flattened -> dotted; longest first so `_attn1_to_out_0` wins over `_attn1_to_o...`
_UNET_LEAVES = [
    ("_attn1_to_out_0", ".attn1.to_out.0"), ("_attn2_to_out_0", ".attn2.to_out.0"),
    ("_attn1_to_q", ".attn1.to_q"), ("_attn1_to_k", ".attn1.to_k"), ("_attn1_to_v", ".attn1.to_v"),
    ("_attn2_to_q", ".attn2.to_q"), ("_attn2_to_k", ".attn2.to_k"), ("_attn2_to_v", ".attn2.to_v"),
    ("_ff_net_0_proj", ".ff.net.0.proj"), ("_ff_net_2", ".ff.net.2"), ("_proj_in", ".proj_in"), ("_proj_out", ".proj_out"),
    ("_time_emb_proj", ".time_emb_proj"), ("_conv_shortcut", ".conv_shortcut"),
    ("_downsamplers_0_conv", ".downsamplers.0.conv"), ("_upsamplers_0_conv", ".upsamplers.0.conv"),
    ("_conv1", ".conv1"), ("_conv2", ".conv2"),
]
_BLOCK = re.compile(r"^(down_blocks|up_blocks)_(\d+)_(attentions|resnets)_(\d+)(?:_transformer_blocks_(\d+))?$|"
                    r"^(down_blocks|up_blocks)_(\d+)$|^mid_block_(attentions|resnets)_(\d+)(?:_transformer_blocks_(\d+))?$")


def unet_param_name(module: str):
    """`lora_unet_<flattened>` -> diffusers-style `<dotted>.weight`, or None when the module is not one the reference maps."""
    if not module.startswith(_UNET_PREFIX):
        return None
    body = module[len(_UNET_PREFIX):]
    for flat, dotted in _UNET_LEAVES:
        if body.endswith(flat):
            m = _BLOCK.match(body[:-len(flat)])
            if not m:
                return None
            if m.group(1):
                path = f"{m.group(1)}.{m.group(2)}.{m.group(3)}.{m.group(4)}"
                if m.group(5) is not None:
                    path += f".transformer_blocks.{m.group(5)}"
            elif m.group(6):
                path = f"{m.group(6)}.{m.group(7)}"
            else:
                path = f"mid_block.{m.group(8)}.{m.group(9)}"
                if m.group(10) is not None:
                    path += f".transformer_blocks.{m.group(10)}"
            return path + dotted + ".weight"
    return None


def text_param_name(module: str):
    if not module.startswith(_TE_PREFIX):
        return None
    layer, _, leaf = module[len(_TE_PREFIX):].partition("_")
    if not layer.isdigit() or leaf not in _TE_LEAVES:
        return None
    return f"text_model.encoder.layers.{layer}.{_TE_LEAVES[leaf]}.weight"


def delta_weight(down, up, alpha):
    """(alpha / rank) * (up . down) in PyTorch layout, float32 NumPy."""
    import torch
    import torch.nn.functional as F
    down, up = torch.as_tensor(np.asarray(down, np.float32)), torch.as_tensor(np.asarray(up, np.float32))
    scale = float(np.asarray(alpha, np.float32)) / float(up.shape[1])
    if down.ndim == 2:
        w = up @ down
    elif tuple(down.shape[2:4]) == (1, 1):
        w = (up[:, :, 0, 0] @ down[:, :, 0, 0])[:, :, None, None]
    else:  # 3x3 down filter followed by a 1x1 up filter, composed into one 3x3 filter
        w = F.conv2d(down.permute(1, 0, 2, 3), up).permute(1, 0, 2, 3)
    return w.numpy() * scale


def _to_numpy(t):
    if hasattr(t, "detach"):
        return t.detach().float().cpu().numpy()
    return np.asarray(t, np.float32)


def load_lora(path_or_dict):
    """-> (text_encoder_deltas, unet_deltas): {parameter name: float32 delta in PyTorch layout}."""
    if isinstance(path_or_dict, dict):
        sd = path_or_dict
    elif str(path_or_dict).endswith(".safetensors"):
        from safetensors import safe_open
        sd = {}
        with safe_open(str(path_or_dict), framework="pt", device="cpu") as f:
            for k in f.keys():
                sd[k] = f.get_tensor(k)
    else:
        import torch
        sd = torch.load(str(path_or_dict), map_location="cpu")
    text, unet = {}, {}
    for key in sd:
        if not str(key).endswith(".alpha"):
            continue
        module = str(key)[:-len(".alpha")]
        down, up = sd.get(module + ".lora_down.weight"), sd.get(module + ".lora_up.weight")
        if down is None or up is None:
            continue
        name, dst = text_param_name(module), text
        if name is None:
            name, dst = unet_param_name(module), unet
        if name is None:
            continue
        dst[name] = delta_weight(_to_numpy(down), _to_numpy(up), _to_numpy(sd[key]))
    return text, unet


def merge(state_dict: dict, deltas: dict, alias: dict | None = None):
    """state_dict + deltas.  `alias` maps checkpoint keys to the names the deltas are filed under (the UNet's LDM key ->
    diffusers alias table); without it the deltas are looked up by the checkpoint keys themselves (text encoder).
    Returns (merged dict, number of deltas applied, names that matched nothing)."""
    import torch
    out, used = dict(state_dict), set()
    for key in state_dict:
        names = [key] if alias is None else [alias.get(key), key]  # a checkpoint may already be in diffusers naming
        name = next((n for n in names if n is not None and n in deltas), None)
        if name is None:
            continue
        d = deltas[name]
        w = state_dict[key]
        w = w.detach().float().cpu().numpy() if hasattr(w, "detach") else np.asarray(w, np.float32)
        if w.shape != d.shape:
            raise ValueError(f"LoRA delta for {name} has shape {d.shape}, the checkpoint tensor {key} has {w.shape}")
        out[key] = torch.from_numpy(w + d)
        used.add(name)
    return out, len(used), sorted(set(deltas) - used)
