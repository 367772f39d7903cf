"""CPU suite (-m "not gpu"): the oracle against the golden vectors recorded from the real reference, the host-side
logic of the product (schedules, kernel coefficients, weight-name rules, preprocessing), and the C-ABI surface."""
import ctypes
import hashlib
import json
import os
import re
import subprocess

import numpy as np
import pytest

from minsdtf_b200 import keys as K
from oracle.scheduler_oracle import OracleScheduler, cfg_combine, timestep_embedding as oracle_temb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
G = np.load(os.path.join(GOLD, "scheduler.npz"))


# ------------------------------------------------------------------------------------------------ oracle vs reference
def test_oracle_schedule_constants_match_reference():
    s = OracleScheduler(False)
    assert np.array_equal(s.alphas_cumprod, G["alphas_cumprod"])
    assert np.array_equal(s.signal_rates, G["signal_rates"]) and np.array_equal(s.noise_rates, G["noise_rates"])


@pytest.mark.parametrize("n", [1, 4, 25, 50])
def test_oracle_ddim_timesteps(n):
    s = OracleScheduler(False)
    s.set_timesteps(n)
    assert np.array_equal(s.timesteps, G[f"ddim_timesteps_{n}"])


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_oracle_tcd_timesteps(n):
    s = OracleScheduler(True)
    s.set_timesteps(n)
    assert np.array_equal(s.timesteps, G[f"tcd_timesteps_{n}"])


@pytest.mark.parametrize("name,n,tcd", [("ddim25", 25, False), ("ddim4", 4, False), ("tcd4", 4, True)])
def test_oracle_step_trajectories_bit_exact(name, n, tcd):
    s = OracleScheduler(tcd)
    s.set_timesteps(n)
    x = G[f"{name}_x0"]
    if tcd:
        np.random.seed(123456)
    for i, t in enumerate(s.timesteps):
        x = s.step(G[f"{name}_eps"][i], int(t), x)
        assert np.array_equal(np.asarray(x, np.float64), G[f"{name}_out"][i]), (name, i)


def test_oracle_img2img_slicing_and_steps():
    s = OracleScheduler(False)
    s.set_timesteps(50)
    asc = s.timesteps[::-1]
    n = int(50 * 0.8 + 0.5)
    assert asc[n] == G["i2i_init_time"] == 800
    assert np.array_equal(asc[:n], G["i2i_timesteps"]) and asc[:n][-1] == 780  # first UNet call is one step below
    x = G["i2i_x0"]
    for i, t in enumerate(asc[:n][::-1]):
        x = s.step(G["i2i_eps"][i], int(t), x)
        assert np.array_equal(np.asarray(x, np.float64), G["i2i_out"][i])


# ------------------------------------------------------------------------------------------------ product host logic
def test_product_scheduler_matches_reference_schedules():
    from minsdtf_b200.scheduler import Scheduler
    s = Scheduler(active_tcd=False)
    assert np.array_equal(s.alphas_cumprod, G["alphas_cumprod"])
    for n in (1, 4, 25, 50):
        s.set_timesteps(n)
        assert np.array_equal(s.timesteps, G[f"ddim_timesteps_{n}"])
    s = Scheduler(active_tcd=True)
    for n in (1, 2, 4, 8):
        s.set_timesteps(n)
        assert np.array_equal(s.timesteps, G[f"tcd_timesteps_{n}"])
    with pytest.raises(ValueError):
        s.set_timesteps(51)


@pytest.mark.parametrize("name,n,tcd", [("ddim25", 25, False), ("ddim4", 4, False), ("tcd4", 4, True), ("i2i", 50, False)])
def test_kernel_coefficients_reproduce_reference_steps(name, n, tcd):
    """x' = ca x + cb eps + cn z evaluated in fp32 (what the kernel does) vs the reference's fp64 trajectory"""
    from minsdtf_b200.scheduler import Scheduler
    s = Scheduler(active_tcd=tcd)
    s.set_timesteps(n)
    ts = [int(t) for t in s.timesteps]
    if name == "i2i":
        ts = [int(t) for t in G["i2i_timesteps"][::-1]]
    x = G[f"{name}_x0"].astype(np.float32)
    if tcd:
        np.random.seed(123456)
    coefs = s.coefficients(ts, 7.5, 0.7)
    assert len(coefs) == len(ts)
    for i, (t, c) in enumerate(zip(ts, coefs)):
        z = np.random.randn(*x.shape).astype(np.float32) if c.cn != 0.0 else 0.0
        x = (np.float32(c.ca) * x + np.float32(c.cb) * G[f"{name}_eps"][i] + np.float32(c.cn) * z).astype(np.float32)
        assert np.abs(x - G[f"{name}_out"][i]).max() < 2e-5, (name, i)
        assert abs(c.sig_t - s.signal_rates[t]) < 1e-7 and abs(c.noi_t - s.noise_rates[t]) < 1e-7


def test_timestep_embedding_cos_first():
    from minsdtf_b200.scheduler import timestep_embedding
    e = timestep_embedding(500)
    assert e.shape == (320,) and e.dtype == np.float32
    assert np.allclose(e[0], np.cos(500.0)) and np.allclose(e[160], np.sin(500.0))
    assert np.abs(oracle_temb(500, 2)[1].astype(np.float32) - e).max() <= 1e-6


def _digest(lines):
    h = hashlib.sha256()
    for ln in lines:
        h.update(ln.encode())
        h.update(b"\n")
    return h.hexdigest()


@pytest.mark.parametrize("comp,ref_name,count,params", [
    ("unet", "civitai_model", 686, 859520964), ("controlnet", "controlnet", 324, 360192640),
    ("hintnet", "hintnet", 16, 1086480), ("decoder", "decoder", 140, 49490199), ("encoder", "encoder", 108, 34163664)])
def test_weight_key_rules_match_reference_tables(comp, ref_name, count, params):
    gold = json.load(open(os.path.join(GOLD, "ckpt_tables.json")))
    ks = K.COMPONENT_KEYS[comp]()
    assert len(ks) == count == gold[ref_name]["count"]
    assert K.n_params(ks) == params
    perm = {4: "(2, 3, 1, 0)", 2: "(1, 0)", 1: "None"}
    lines = sorted(f"{k}|{perm[len(shp)]}" for k, shp in ks.items())
    assert _digest(lines) == gold[ref_name]["sha256_sorted"]  # same names AND same layout permutations


def test_unet_diffusers_aliases_match_reference():
    gold = json.load(open(os.path.join(GOLD, "ckpt_tables.json")))
    amap = K.unet_alias_map()
    assert len(amap) == gold["unet_alias"]["count"]
    assert _digest(sorted(f"{k}->{v}" for k, v in amap.items())) == gold["unet_alias"]["sha256_sorted"]


def test_synthetic_weights_are_deterministic_and_shaped():
    from minsdtf_b200 import synth
    a, b = synth.make_state_dict("hintnet"), synth.make_state_dict("hintnet")
    for k, shp in K.hintnet_keys().items():
        assert tuple(a[k].shape) == shp and bool((a[k] == b[k]).all())
    c = synth.make_state_dict("hintnet", seed=1)
    assert not bool((a["control_model.input_hint_block.0.weight"] == c["control_model.input_hint_block.0.weight"]).all())


def test_preprocessing_resize_blur_and_mask():
    from minsdtf_b200.stable_diffusion import StableDiffusionBase
    sd = StableDiffusionBase(img_height=64, img_width=64)
    img = np.arange(32 * 32 * 3, dtype=np.float64).reshape(32, 32, 3)
    assert sd.resize(img, 32, 32) is img
    up = sd.resize(img, 63, 63)  # align-corners: corners preserved, midpoints averaged
    assert np.allclose(up[0, 0], img[0, 0]) and np.allclose(up[-1, -1], img[-1, -1])
    assert np.allclose(up[1, 0], 0.5 * (img[0, 0] + img[1, 0]))
    arr, ten = sd.preprocessed_image(np.full((64, 64, 3), 255, np.uint8))
    assert arr.shape == (1, 64, 64, 3) and np.allclose(arr, 1.0) and np.allclose(ten, 1.0)
    blurred = sd.gaussian_blur(np.ones((8, 8, 1), np.float32), radius=5, h_axis=0, v_axis=1)
    assert np.allclose(blurred, 1.0)  # normalised binomial taps
    m = np.zeros((64, 64), np.uint8)
    m[16:48, 16:48] = 255
    pix, lat = sd.preprocessed_mask(m, 5)
    assert pix.shape == (1, 64, 64, 1) and lat.shape == (1, 8, 8, 1)
    assert 0.0 <= lat.min() and lat.max() <= 1.0 and lat[0, 4, 4, 0] == pytest.approx(1.0)


def test_generate_image_argument_errors_match_reference():
    from minsdtf_b200.stable_diffusion import StableDiffusion
    sd = StableDiffusion(synthetic=True)
    from minsdtf_b200.stable_diffusion import StableDiffusionBase
    with pytest.raises(ValueError):  # stable_diffusion.py:377-382
        StableDiffusionBase.generate_image(sd, np.zeros((77, 768), np.float32), diffusion_noise=np.zeros((64, 64, 4)), seed=1)
    with pytest.raises(FileNotFoundError):
        sd.encode_text("a prompt")  # string prompts need the CLIP BPE vocabulary, a download in the reference: says so
    with pytest.raises(ValueError):  # stable_diffusion.py:205-206
        type(sd).tokenizer.fset(sd, object()) or sd.encode_text("a prompt", "/no/such/embedding.pt")
    with pytest.raises(ValueError):  # both GPUs of a CFG pair must start from the same latent
        StableDiffusionBase.generate_image(sd, np.zeros((77, 768), np.float32), cfg_split=True)


# ------------------------------------------------------------------------------------------------ oracle graphs
def test_oracle_small_graphs_run_and_shapes():
    from minsdtf_b200 import synth
    from oracle import sd15_oracle as O
    sd = synth.make_controlnet_state_dict()
    img = (synth.edge_map(64, 64).astype(np.float32) / 255.0)[None]
    hint = O.hintnet_forward(sd, img)
    assert hint.shape == (1, 8, 8, 320) and np.isfinite(hint).all()
    res = O.controlnet_forward(sd, synth.latents(1, 8, 8), oracle_temb(10, 1), synth.context(1), hint)
    assert [r.shape[-1] for r in res] == [320] * 4 + [640] * 3 + [1280] * 6
    assert [r.shape[1] for r in res] == [8, 8, 8, 4, 4, 4, 2, 2, 2, 1, 1, 1, 1]


def test_cfg_combine_matches_hand_computation():
    rng = np.random.default_rng(0)
    u, c = rng.standard_normal((2, 4, 4, 4)), rng.standard_normal((2, 4, 4, 4))
    e = u + 7.5 * (c - u)
    assert np.allclose(cfg_combine(u, c, 7.5, 0.0), e)
    r = cfg_combine(u, c, 7.5, 0.7)
    f = 0.7 * c.std(axis=(1, 2, 3), keepdims=True) / (e.std(axis=(1, 2, 3), keepdims=True) + 1e-5) + 0.3
    assert np.allclose(r, e * f)


# ------------------------------------------------------------------------------------------------ C ABI surface
def test_library_builds_loads_and_exports_every_declared_symbol():
    from minsdtf_b200 import _lib, build
    path = build.build_lib()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "sdtf.h")).read()
    declared = set(re.findall(r"\b(sdtf_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sdtf.h but not exported"
    assert set(_lib.EXPORTS) <= declared
    lib.sdtf_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.sdtf_version()
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):  # tcgen05.mma / TMA / tcgen05.ld really are in the binary
        assert mnemonic in sass, mnemonic


def test_engine_creation_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from minsdtf_b200.engine import Engine, EngineError
    with pytest.raises(EngineError) as ei:
        Engine(0)
    assert "no CPU fallback" in str(ei.value) or "no fallback" in str(ei.value)


def test_read_checkpoint_formats(tmp_path):
    """.safetensors, a torch pickle, and a pickle wrapped in {"state_dict": ...} all come back as {key: tensor}
    (ckpt_loader.py:2139-2145; `control_sd15_canny.pth` and `.ckpt` files are pickles)"""
    import torch
    from minsdtf_b200 import synth
    from minsdtf_b200.engine import Engine
    sd = {k: v for k, v in list(synth.make_state_dict("hintnet").items())[:6]}
    synth.save_safetensors(sd, str(tmp_path / "a.safetensors"))
    torch.save(sd, str(tmp_path / "b.pth"))
    torch.save({"state_dict": sd, "epoch": 3}, str(tmp_path / "c.ckpt"))
    for name in ("a.safetensors", "b.pth", "c.ckpt"):
        got = Engine.read_checkpoint(str(tmp_path / name))
        assert set(got) == set(sd)
        assert all(torch.equal(got[k], sd[k]) for k in sd)


def test_lora_names_and_merge():
    """kohya module names -> the diffusers-style aliases of the reference's UNET_KEY_MAPPING (ckpt_loader.py:2231-2272), and
    the merge adds (alpha / rank) * up . down to the tensor the alias belongs to"""
    import torch
    from minsdtf_b200 import keys as K, lora, synth
    assert lora.unet_param_name("lora_unet_down_blocks_0_attentions_1_transformer_blocks_0_attn2_to_out_0") == \
        "down_blocks.0.attentions.1.transformer_blocks.0.attn2.to_out.0.weight"
    assert lora.unet_param_name("lora_unet_mid_block_resnets_1_time_emb_proj") == "mid_block.resnets.1.time_emb_proj.weight"
    assert lora.unet_param_name("lora_unet_up_blocks_2_upsamplers_0_conv") == "up_blocks.2.upsamplers.0.conv.weight"
    assert lora.unet_param_name("lora_unet_conv_in") is None
    assert lora.text_param_name("lora_te_text_model_encoder_layers_11_self_attn_out_proj") == \
        "text_model.encoder.layers.11.self_attn.out_proj.weight"
    te, un = lora.load_lora(synth.make_lora_state_dict())
    alias = K.unet_alias_map()
    assert set(un) <= set(alias.values()) and len(te) == 12
    name = next(iter(un))
    key = next(k for k, a in alias.items() if a == name)
    base = {key: torch.zeros(K.unet_keys()[key])}
    merged, n, unused = lora.merge(base, un, alias)
    assert n == 1 and len(unused) == len(un) - 1 and np.allclose(merged[key].numpy(), un[name])
    with pytest.raises(ValueError):
        lora.merge({key: torch.zeros(3, 3)}, un, alias)
