"""The engine against outputs of the REFERENCE'S OWN CODE (tests/golden/reference_run.npz, produced here in the
authoring container by tools/make_golden_ref.py: /root/reference/stable_diffusion/*.py unmodified on the Keras stand-in
of oracle/keras_shim).  The reference tree does not travel to the GPU box, its recorded outputs do.  Bars as in
BASELINE.json: eps / model outputs max|d|/max|ref| <= 2e-2; free-running loops are looser (they accumulate)."""
import os

import numpy as np
import pytest

import golden_cases as G
from minsdtf_b200 import synth
from minsdtf_b200.scheduler import timestep_embedding

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_run.npz")
BAR = 2e-2


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())


def psnr_u8(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def vocab(tmp_path_factory):
    return synth.make_bpe_vocab(str(tmp_path_factory.mktemp("bpe") / "bpe_vocab.txt.gz"))


@pytest.fixture(scope="module")
def pipe(engine_unet, engine_vae, engine_cnet, vocab):
    from minsdtf_b200.stable_diffusion import StableDiffusion
    if "text_encoder" not in engine_unet.loaded:
        engine_unet.load_state_dict(synth.make_state_dict("text_encoder"), "text_encoder")
    return StableDiffusion(img_height=G.H, img_width=G.H, engine=engine_unet, synthetic=True, bpe_vocab=vocab)


def test_models_against_reference_outputs(pipe, gold):
    e = pipe.engine
    lat, ctx = synth.latents(2, G.h, G.h, seed=1), synth.context(2, 77, seed=2)
    te = np.repeat(timestep_embedding(500)[None], 2, 0)
    assert rel(e.unet(lat, te, ctx), gold["unet_eps"]) <= BAR
    img = (G.edges().astype(np.float32) / 255.0)[None]
    hint = e.hintnet(img)
    assert rel(hint[..., ::8], gold["hint"]) <= BAR
    res = e.controlnet(lat[:1], te[:1], ctx[:1], hint)
    worst = max(rel(r[..., ::8], gold[f"control_{i}"]) for i, r in enumerate(res))
    assert worst <= 3e-2, worst  # the hint fed here is the engine's own (bf16), not the reference's
    assert rel(e.unet(lat[:1], te[:1], ctx[:1], res), gold["unet_eps_control"]) <= 3e-2
    l2 = synth.latents(1, G.h, G.h, seed=3) * 0.18215 * 3.0
    dec = e.vae_decode(l2)
    u_ref = np.clip((gold["decoded"] + 1.0) * 0.5 * 255.0, 0, 255).astype(np.uint8)
    assert psnr_u8(e.to_uint8(dec), u_ref) >= 30.0
    src = G.source_image().astype(np.float32)[None] / 127.5 - 1.0
    assert rel(e.vae_encode(src), gold["encoded"]) <= BAR


def test_text_path_against_reference_outputs(pipe, gold):
    """tokenizer + attention syntax + windows + weights + textual inversion -> context (text tower bar: 1e-2)"""
    for name, args in (("ctx_prompt", (G.PROMPT,)), ("ctx_negative", (G.NEGATIVE,)), ("ctx_long", (G.LONG_PROMPT,)),
                       ("ctx_ti", (G.PROMPT, G.ti_embedding()))):
        got = pipe.encode_text(*args)
        assert got.shape[:2] == gold[name].shape[:2], name
        r = rel(got[..., ::4], gold[name])
        print(f"{name}: rel {r:.4g}")
        assert r <= 1e-2, (name, r)
    assert rel(pipe._get_unconditional_context()[..., ::4], gold["ctx_empty"]) <= 1e-2


def test_loops_against_reference_outputs(pipe, gold):
    noise = G.noise()
    ctx, _ = G.contexts()

    def check(p, name, lat_bar=5e-2, **kw):
        img, lat = p.generate_image(ctx, batch_size=1, diffusion_noise=noise, return_latent=True, **kw)
        r, q = rel(lat, gold[name + "_latent"]), psnr_u8(img, gold[name + "_image"])
        print(f"{name}: final latent rel {r:.4g}, image psnr {q:.2f} dB")
        assert r <= lat_bar, (name, r)
        assert q >= 28.0, (name, q)

    check(pipe, "txt2img", num_steps=4, guidance_rescale=0.7)
    check(pipe, "txt2img_norescale", num_steps=4)
    check(pipe, "img2img", num_steps=10, guidance_rescale=0.7, reference_image=G.source_image(), reference_image_strength=0.8)
    check(pipe, "inpaint", num_steps=10, guidance_rescale=0.7, reference_image=G.source_image(), reference_image_strength=0.8,
          inpaint_mask=G.mask(), mask_blur_strength=5)
    check(pipe, "controlnet", num_steps=3, control_net_image=G.edges())
    from minsdtf_b200.stable_diffusion import StableDiffusion
    tcd = StableDiffusion(img_height=G.H, img_width=G.H, engine=pipe.engine, synthetic=True, active_tcd=True)
    np.random.seed(123456)
    check(tcd, "tcd", num_steps=4, unconditional_guidance_scale=0.0)


def test_entry_points_with_string_prompts_against_reference_outputs(pipe, gold):
    noise = G.noise()
    pipe._get_initial_diffusion_noise = lambda batch_size, seed: noise
    img = pipe.text_to_image(G.PROMPT, negative_prompt=G.NEGATIVE, batch_size=1, num_steps=4, seed=7)
    q = psnr_u8(img, gold["text_to_image_image"])
    print(f"text_to_image: {q:.2f} dB")
    assert q >= 28.0
    # 154-token prompt against the 77-token empty negative prompt: two passes per step (sdtf_denoise)
    img = pipe.text_to_image(G.LONG_PROMPT, batch_size=1, num_steps=3, seed=7)
    q = psnr_u8(img, gold["text_to_image_long_image"])
    print(f"text_to_image (long prompt): {q:.2f} dB")
    assert q >= 28.0
    seen = []
    img = pipe.inpaint(G.PROMPT, negative_prompt=G.NEGATIVE, batch_size=1, num_steps=10, seed=7, reference_image=G.source_image(),
                       inpaint_mask=G.mask(), mask_blur_strength=5, callback=seen.append)
    assert seen == list(gold["inpaint_entry_callbacks"])
    q = psnr_u8(img, gold["inpaint_entry_image"])
    print(f"inpaint: {q:.2f} dB")
    assert q >= 28.0


def test_lora_merged_model_against_reference_outputs(gold, vocab):
    """a second engine: the LoRA-merged UNet / text encoder must not disturb the session's shared one"""
    from minsdtf_b200.engine import Engine
    from minsdtf_b200.stable_diffusion import StableDiffusion
    eng = Engine(0)
    try:
        sd = StableDiffusion(img_height=G.H, img_width=G.H, engine=eng, synthetic=True, bpe_vocab=vocab,
                             lora_path=synth.make_lora_state_dict())
        lat, ctx = synth.latents(1, G.h, G.h, seed=1), synth.context(1, 77, seed=2)
        te = timestep_embedding(500)[None]
        eps = sd.diffusion_model.predict_on_batch([lat, te, ctx])
        assert rel(eps, gold["lora_unet_eps"]) <= BAR
        assert rel(eps, gold["unet_eps"][:1]) > 1e-3
        assert rel(sd.encode_text(G.PROMPT)[..., ::4], gold["lora_ctx_prompt"]) <= 1e-2
        # re-finalising a component frees the previous packed weights (one device pool per component)
        import torch
        free0 = torch.cuda.mem_get_info()[0]
        for _ in range(2):
            eng.load_state_dict(synth.make_state_dict("unet"), "unet")
        assert free0 - torch.cuda.mem_get_info()[0] < 512 << 20
    finally:
        eng.close()
