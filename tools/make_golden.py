#!/usr/bin/env python
"""Generate tests/golden/* from the REAL reference (run in the authoring container only; /root/reference does not
exist on the GPU box).  Two parts of the reference import without Keras:

  * stable_diffusion/scheduler.py (NumPy only)  -> tests/golden/scheduler.npz : known-answer vectors for
    set_timesteps / step (DDIM and TCD), the img2img timestep slicing of stable_diffusion.py:406-416 and the
    initial-latent noising of :559-568.
  * stable_diffusion/ckpt_loader.py tables      -> tests/golden/ckpt_tables.json : per-component key count,
    parameter-independent sha256 digests of the ordered "key|perm" lines and of the UNet alias map.

The Keras graphs themselves cannot be run here (no keras / tensorflow wheel, no network) — see DESIGN.md.
"""
import hashlib
import importlib.util
import json
import os
import sys

import numpy as np

REF = os.environ.get("SDTF_REFERENCE", "/root/reference/stable_diffusion")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def digest(lines):
    h = hashlib.sha256()
    for ln in lines:
        h.update(ln.encode())
        h.update(b"\n")
    return h.hexdigest()


def scheduler_golden():
    S = _load("scheduler").Scheduler
    g = {}
    s = S(active_tcd=False)
    g["alphas_cumprod"] = s.alphas_cumprod
    g["signal_rates"] = s.signal_rates
    g["noise_rates"] = s.noise_rates
    rng = np.random.default_rng(20240517)
    shape = (2, 8, 8, 4)
    for n in (1, 4, 25, 50):
        s = S(active_tcd=False)
        s.set_timesteps(n)
        g[f"ddim_timesteps_{n}"] = s.timesteps
    # DDIM trajectory, 25 steps, driven exactly like stable_diffusion.py:399-400,442,468
    for n in (25, 4):
        s = S(active_tcd=False)
        s.set_timesteps(n)
        ts = s.timesteps[::-1]
        x = rng.standard_normal(shape).astype(np.float32)
        g[f"ddim{n}_x0"] = x
        eps_all, out_all = [], []
        for index, t in list(enumerate(ts))[::-1]:
            eps = rng.standard_normal(shape).astype(np.float32)
            x = s.step(eps, t, x)
            eps_all.append(eps)
            out_all.append(np.asarray(x, dtype=np.float64))
        g[f"ddim{n}_eps"] = np.stack(eps_all)
        g[f"ddim{n}_out"] = np.stack(out_all)
    # img2img slicing: 50 steps, strength 0.8 (stable_diffusion.py:406-416)
    s = S(active_tcd=False)
    s.set_timesteps(50)
    ts = s.timesteps[::-1]
    n = int(50 * 0.8 + 0.5)
    g["i2i_init_time"] = np.asarray(ts[n])
    g["i2i_timesteps"] = ts[:n]
    x = rng.standard_normal(shape).astype(np.float32)
    g["i2i_x0"] = x
    eps_all, out_all = [], []
    for index, t in list(enumerate(ts[:n]))[::-1]:
        eps = rng.standard_normal(shape).astype(np.float32)
        x = s.step(eps, t, x)
        eps_all.append(eps)
        out_all.append(np.asarray(x, dtype=np.float64))
    g["i2i_eps"] = np.stack(eps_all)
    g["i2i_out"] = np.stack(out_all)
    # TCD: timesteps and a 4-step trajectory with the global NumPy RNG seeded (scheduler.py:301)
    for n in (1, 2, 4, 8):
        s = S(active_tcd=True)
        s.set_timesteps(n)
        g[f"tcd_timesteps_{n}"] = s.timesteps
    s = S(active_tcd=True)
    s.set_timesteps(4)
    ts = s.timesteps[::-1]
    x = rng.standard_normal(shape).astype(np.float32)
    g["tcd4_x0"] = x
    np.random.seed(123456)
    eps_all, out_all = [], []
    for index, t in list(enumerate(ts))[::-1]:
        eps = rng.standard_normal(shape).astype(np.float32)
        x = s.step(eps, t, x)
        eps_all.append(eps)
        out_all.append(np.asarray(x, dtype=np.float64))
    g["tcd4_eps"] = np.stack(eps_all)
    g["tcd4_out"] = np.stack(out_all)
    np.savez_compressed(os.path.join(OUT, "scheduler.npz"), **g)
    print("scheduler.npz:", {k: v.shape for k, v in g.items()})


def table_golden():
    m = _load("ckpt_loader")
    out = {}
    for comp, table in m.CKPT_MAPPING.items():
        lines = [f"{k}|{p}" for k, p in table]
        out[comp] = {"count": len(table), "sha256_ordered": digest(lines), "sha256_sorted": digest(sorted(lines)),
                     "sha256_sorted_keys": digest(sorted(k for k, _ in table))}
    al = m.UNET_KEY_MAPPING
    out["unet_alias"] = {"count": len(al), "sha256_sorted": digest(sorted(f"{k}->{v}" for k, v in al.items()))}
    with open(os.path.join(OUT, "ckpt_tables.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("ckpt_tables.json:", {k: v["count"] for k, v in out.items()})
    # verbose cross-check of the rule-based generator while the real tables are at hand
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from minsdtf_b200 import keys as K
    comp_map = {"civitai_model": "unet", "controlnet": "controlnet", "hintnet": "hintnet", "decoder": "decoder",
                "encoder": "encoder"}
    for comp, mine in comp_map.items():
        ref_keys = [k for k, _ in m.CKPT_MAPPING[comp]]
        my = K.COMPONENT_KEYS[mine]()
        missing = [k for k in ref_keys if k not in my]
        extra = [k for k in my if k not in set(ref_keys)]
        perm_bad = []
        for k, p in m.CKPT_MAPPING[comp]:
            if k in my:
                exp = {4: (2, 3, 1, 0), 2: (1, 0), 1: None}[len(my[k])]
                if p != exp:
                    perm_bad.append((k, p, my[k]))
        print(f"{comp}: ref {len(ref_keys)} mine {len(my)} params {K.n_params(my)} missing {missing[:3]} extra {extra[:3]} "
              f"perm_bad {perm_bad[:3]}")
    amap = K.unet_alias_map()
    bad = [(k, al[k], amap.get(k)) for k in al if amap.get(k) != al[k]]
    print("alias mismatches:", bad[:5], len(bad))


def text_tower_golden():
    """tests/golden/text_tower.npz: outputs of transformers.CLIPTextModel (an implementation of the text tower that is
    independent of this repository and of its oracle) on the seeded synthetic text-encoder weights: tokens (2,77) and the
    context for clip_skip -1 / -2, every 4th channel, float16.  The Keras TextEncoder of the reference cannot be run
    here; its checkpoint (`text_encoder/model.safetensors`, text_encoder.py:112) is the HF CLIPTextModel's."""
    import torch
    import transformers
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from minsdtf_b200 import synth
    sd = synth.make_state_dict("text_encoder")
    cfg = transformers.CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                                      num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu",
                                      layer_norm_eps=1e-5, eos_token_id=49407, bos_token_id=49406, pad_token_id=49407)
    model = transformers.CLIPTextModel(cfg).eval()
    model.load_state_dict(sd, strict=False)
    tokens = synth.prompt_tokens(2)
    with torch.no_grad():
        out = model(input_ids=torch.as_tensor(tokens, dtype=torch.long), output_hidden_states=True)
        c1 = out.last_hidden_state.numpy()
        c2 = model.text_model.final_layer_norm(out.hidden_states[-2]).numpy()
    np.savez_compressed(os.path.join(OUT, "text_tower.npz"), tokens=tokens, ctx_skip1=c1[..., ::4].astype(np.float16),
                        ctx_skip2=c2[..., ::4].astype(np.float16), max_abs=np.float32(np.abs(c1).max()))
    print("text tower golden:", c1.shape, float(np.abs(c1).max()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    scheduler_golden()
    table_golden()
    text_tower_golden()
