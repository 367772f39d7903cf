"""The reference's own code, executed (VERDICT r01 item 2 / "what's missing" 6).

`/root/reference/stable_diffusion/*.py` runs UNMODIFIED on the torch-backed Keras stand-in of oracle/keras_shim
(tests/ref_harness.py): its model builders, its positional checkpoint loader, its tokenizer, its prompt weighting and
its `generate_image` loop.  Checked here:
  * the stand-in's `model.weights` order reproduces the reference's CKPT_MAPPING tables (shape by shape, and name by
    name where the reference names its layers) — the positional loader depends on it;
  * the CPU oracle (oracle/sd15_oracle.py, text_oracle.py — what the GPU parity tests compare the engine with) agrees
    with the reference's graphs to float32 rounding, model by model and over whole denoising loops;
  * the product's host side (`minsdtf_b200.StableDiffusion` + prompt_weighting + lora + bpe), driven with oracle-backed
    models, reproduces the reference's loop: txt2img, img2img, inpaint, ControlNet, TCD, string prompts with attention
    syntax, long prompts, textual inversion, LoRA, per-iteration callbacks;
  * tests/golden/reference_run.npz (what the GPU tests use, the reference tree does not travel) is what the reference
    produces today.
Tests that need /root/reference are skipped where it does not exist; the ones that only need the committed fixtures run
everywhere."""
import os

import numpy as np
import pytest

import golden_cases as G
import ref_harness as RH
from minsdtf_b200 import synth
from minsdtf_b200.scheduler import timestep_embedding
from oracle import sd15_oracle as O, text_oracle as TO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_run.npz")
needs_ref = pytest.mark.skipif(not RH.available(), reason="/root/reference is not on this machine")
TOL = 2e-4  # float32 graphs of ~100 layers evaluated in two different operation orders


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def ck(tmp_path_factory):
    return RH.write_checkpoints(str(tmp_path_factory.mktemp("ckpt")))


@pytest.fixture(scope="module")
def ref(ck):
    return RH.reference_pipeline(ck[0])


@pytest.fixture(scope="module")
def ref_control(ck):
    return RH.reference_pipeline(ck[0], control=True)


def _ours(ck, **kw):
    """the product pipeline with oracle-backed models (tests/oracle_engine.py)"""
    from minsdtf_b200.stable_diffusion import StableDiffusion
    from oracle_engine import OracleEngine
    paths, sds = ck
    eng = OracleEngine()
    sd = StableDiffusion(img_height=G.H, img_width=G.H, unet_ckpt=sds["unet"], vae_ckpt=sds["vae"], text_encoder_ckpt=sds["text_encoder"],
                         controlnet_path=sds["controlnet"], engine=eng, bpe_vocab=paths["vocab"], **kw)
    return sd, eng


# ------------------------------------------------------------------------------------------------ weight layout (a12)
@needs_ref
def test_shim_weight_order_matches_reference_tables(ref, ref_control):
    """model.weights of every graph, in order, has the shapes (and for the UNet the names) of the reference's
    CKPT_MAPPING entries after their permutation — i.e. the reference's positional loader fills the right variables"""
    RH.import_reference()
    from stable_diffusion.ckpt_loader import CKPT_MAPPING, UNET_KEY_MAPPING
    from minsdtf_b200 import keys as K
    shapes = {}
    for gen in (K.unet_keys, K.controlnet_keys, K.hintnet_keys, K.vae_decoder_keys, K.vae_encoder_keys):
        shapes.update(gen())
    models = {"civitai_model": ref.diffusion_model, "decoder": ref.image_decoder.model, "encoder": RH.quiet(lambda: ref.image_encoder),
              "controlnet": RH.quiet(lambda: ref_control.control_net), "hintnet": RH.quiet(lambda: ref_control.hint_net)}
    for table, model in models.items():
        mapping = CKPT_MAPPING[table]
        ws = model.weights
        assert len(ws) == len(mapping), table
        owner = {}
        for layer in model.layers:
            for v in layer.weights:
                owner[id(v)] = layer.name
        for v, (key, perm) in zip(ws, mapping):
            shp = tuple(shapes[key])
            if perm is not None:
                shp = tuple(shp[p] for p in perm)
            assert v.shape == shp, (table, key, v.name)
            if table == "civitai_model":  # the UNet's layers carry diffusers-style names: the alias of the key starts with it
                assert UNET_KEY_MAPPING[key].startswith(owner[id(v)] + "."), (key, UNET_KEY_MAPPING[key], owner[id(v)])


# ------------------------------------------------------------------------------------------------ graphs (a5, a9-a11, f1)
@needs_ref
def test_oracle_equals_reference_graphs(ref, ref_control, ck, gold):
    sds = ck[1]
    lat, ctx = synth.latents(2, G.h, G.h, seed=1), synth.context(2, 77, seed=2)
    te = np.repeat(timestep_embedding(500)[None], 2, 0)
    eps = ref.diffusion_model.predict_on_batch([lat, te, ctx])
    assert rel(O.unet_forward(sds["unet"], lat, te, ctx), eps) <= TOL
    assert rel(gold["unet_eps"], eps) <= 1e-5
    img = (G.edges().astype(np.float32) / 255.0)[None]
    hint = ref_control.hint_net.predict_on_batch(img)
    assert rel(O.hintnet_forward(sds["controlnet"], img), hint) <= TOL
    res = ref_control.control_net.predict_on_batch([lat[:1], te[:1], ctx[:1], hint])
    mine = O.controlnet_forward(sds["controlnet"], lat[:1], te[:1], ctx[:1], hint)
    assert len(res) == 13 and max(rel(a, b) for a, b in zip(mine, res)) <= TOL
    assert all(rel(gold[f"control_{i}"], r[..., ::8]) <= 1e-5 for i, r in enumerate(res))
    eps_c = ref_control.diffusion_model.predict_on_batch([lat[:1], te[:1], ctx[:1]] + list(res))
    assert rel(O.unet_forward(sds["unet"], lat[:1], te[:1], ctx[:1], res), eps_c) <= TOL
    l2 = synth.latents(1, G.h, G.h, seed=3) * 0.18215 * 3.0
    dec = ref.image_decoder.predict_on_batch(l2)
    assert rel(O.vae_decode(sds["vae"], l2), dec) <= TOL and rel(gold["decoded"], dec) <= 1e-5
    src = G.source_image().astype(np.float32)[None] / 127.5 - 1.0
    enc = ref.image_encoder.predict_on_batch(src)
    assert rel(O.vae_encode(sds["vae"], src), enc) <= TOL and rel(gold["encoded"], enc) <= 1e-5
    tok = synth.prompt_tokens(2)
    pos = np.asarray([list(range(77))], np.int32)
    c = ref.text_encoder.predict_on_batch(ref.text_clip_embedding.predict_on_batch([tok, pos]))
    assert rel(TO.text_encode(sds["text_encoder"], tok, -1), c) <= TOL


# ------------------------------------------------------------------------------------------------ the loop (a1, a4, a6-a8)
@needs_ref
def test_oracle_loop_equals_reference_loop(ref, ck, gold):
    """O.generate_image (the GPU tests' loop oracle) against the reference's generate_image"""
    sds = ck[1]
    noise = G.noise()
    ctx, _ = G.contexts()
    unc = ref._get_unconditional_context()
    weights = {"unet": sds["unet"], "vae": sds["vae"]}
    img = RH.quiet(ref.generate_image, ctx, batch_size=1, num_steps=4, diffusion_noise=noise, guidance_rescale=0.7)
    lat = np.asarray(ref._image_decoder.last_input, np.float32)
    mine = O.generate_image(weights, ctx, unc, noise, num_steps=4, guidance_scale=7.5, guidance_rescale=0.7, decode=False)
    assert rel(mine, lat) <= TOL
    assert rel(gold["txt2img_latent"], lat) <= 1e-5 and np.abs(gold["txt2img_image"].astype(int) - img.astype(int)).max() <= 1
    # inpaint (covers img2img slicing, re-noising at the current t, pixel blend)
    img = RH.quiet(ref.generate_image, ctx, batch_size=1, num_steps=10, diffusion_noise=noise, guidance_rescale=0.7,
                   reference_image=G.source_image(), reference_image_strength=0.8, inpaint_mask=G.mask(), mask_blur_strength=5)
    lat = np.asarray(ref._image_decoder.last_input, np.float32)
    in_arr, in_t = ref.preprocessed_image(G.source_image())
    m_arr, m_lat = ref.preprocessed_mask(G.mask(), 5)
    init = O.vae_encode(sds["vae"], in_t)
    mine = O.generate_image(weights, ctx, unc, noise, num_steps=10, guidance_scale=7.5, guidance_rescale=0.7, init_latent=init,
                            strength=0.8, latent_mask=m_lat, input_image_array=in_arr, input_mask_array=m_arr, decode=False)
    assert rel(mine, lat) <= TOL
    assert rel(gold["inpaint_latent"], lat) <= 1e-5


# ------------------------------------------------------------------------------------------------ host side of the product
def test_product_host_loop_reproduces_reference_fixtures(ck, gold):
    """minsdtf_b200.StableDiffusion with oracle-backed models against the outputs recorded from the reference's loop
    (needs only the committed fixtures): txt2img, img2img, inpaint, ControlNet, TCD"""
    sd, eng = _ours(ck)
    noise = G.noise()
    ctx, _ = G.contexts()

    def check(name, **kw):
        img, lat = sd.generate_image(ctx, batch_size=1, diffusion_noise=noise, return_latent=True, **kw)
        assert rel(lat, gold[name + "_latent"]) <= TOL, name
        assert np.mean(np.abs(img.astype(int) - gold[name + "_image"].astype(int)) > 1) < 1e-3, name

    check("txt2img", num_steps=4, guidance_rescale=0.7)
    check("txt2img_norescale", num_steps=4)
    check("img2img", num_steps=10, guidance_rescale=0.7, reference_image=G.source_image(), reference_image_strength=0.8)
    check("inpaint", num_steps=10, guidance_rescale=0.7, reference_image=G.source_image(), reference_image_strength=0.8,
          inpaint_mask=G.mask(), mask_blur_strength=5)
    check("controlnet", num_steps=3, control_net_image=G.edges())
    sd_t, _ = _ours(ck, active_tcd=True)
    np.random.seed(123456)
    img, lat = sd_t.generate_image(ctx, batch_size=1, num_steps=4, diffusion_noise=noise, unconditional_guidance_scale=0.0,
                                   return_latent=True)
    assert rel(lat, gold["tcd_latent"]) <= TOL


def test_product_text_path_reproduces_reference_fixtures(ck, gold):
    """string prompts: tokenizer, attention syntax, 75-token windows, mean-preserving weights, textual inversion, the
    public entry points with their defaults, per-iteration callbacks — against the reference's recorded outputs"""
    sd, eng = _ours(ck)
    assert rel(sd.encode_text(G.PROMPT)[..., ::4], gold["ctx_prompt"]) <= TOL
    assert rel(sd.encode_text(G.NEGATIVE)[..., ::4], gold["ctx_negative"]) <= TOL
    long_ctx = sd.encode_text(G.LONG_PROMPT)
    assert long_ctx.shape == (1, 154, 768) and rel(long_ctx[..., ::4], gold["ctx_long"]) <= TOL
    assert rel(sd.encode_text(G.PROMPT, G.ti_embedding())[..., ::4], gold["ctx_ti"]) <= TOL
    assert rel(sd._get_unconditional_context()[..., ::4], gold["ctx_empty"]) <= TOL
    noise = G.noise()
    sd._get_initial_diffusion_noise = lambda batch_size, seed: noise
    img = sd.text_to_image(G.PROMPT, negative_prompt=G.NEGATIVE, batch_size=1, num_steps=4, seed=7)
    assert np.mean(np.abs(img.astype(int) - gold["text_to_image_image"].astype(int)) > 1) < 1e-3
    # a two-window prompt against the one-window empty negative prompt: separate passes per branch, as the reference
    img = sd.text_to_image(G.LONG_PROMPT, batch_size=1, num_steps=3, seed=7)
    assert eng.calls[-1]["T"] == 154 and eng.calls[-1]["Tu"] == 77
    assert np.mean(np.abs(img.astype(int) - gold["text_to_image_long_image"].astype(int)) > 1) < 1e-3
    seen = []
    img = sd.inpaint(G.PROMPT, negative_prompt=G.NEGATIVE, batch_size=1, num_steps=10, seed=7, reference_image=G.source_image(),
                     inpaint_mask=G.mask(), mask_blur_strength=5, callback=seen.append)
    assert seen == list(gold["inpaint_entry_callbacks"]) == list(range(1, 9))
    assert np.mean(np.abs(img.astype(int) - gold["inpaint_entry_image"].astype(int)) > 1) < 1e-3


def test_product_lora_merge_reproduces_reference_fixtures(ck, gold):
    lora_sd = synth.make_lora_state_dict()
    sd, eng = _ours(ck, lora_path=lora_sd)
    lat, ctx = synth.latents(1, G.h, G.h, seed=1), synth.context(1, 77, seed=2)
    te = timestep_embedding(500)[None]
    eps = sd.diffusion_model.predict_on_batch([lat, te, ctx])
    assert rel(eps, gold["lora_unet_eps"]) <= TOL
    assert rel(eps, gold["unet_eps"][:1]) > 1e-3  # the adapters really change the model
    assert rel(sd.encode_text(G.PROMPT)[..., ::4], gold["lora_ctx_prompt"]) <= TOL


@needs_ref
def test_prompt_weighting_and_tokenizer_equal_reference(ck):
    RH.import_reference()
    from stable_diffusion import long_prompt_weighting as RL
    from stable_diffusion.clip_tokenizer import SimpleTokenizer
    from minsdtf_b200 import prompt_weighting as PW
    from minsdtf_b200.bpe import ClipBPE
    paths, sds = ck
    rt, mt = SimpleTokenizer(bpe_path=paths["vocab"]), ClipBPE(paths["vocab"])
    prompts = ["normal text", "an (important) word", "(unbalanced", r"\(literal\]", "(unnecessary)(parens)",
               "a (((house:1.3)) [on] a (hill:0.5), sun, (((sky))).", "", "[[a]] (b:2) c:d) :", "x (y:1.e) \\\\ z", G.PROMPT,
               G.LONG_PROMPT, G.TRUNCATED_PROMPT, "Héllo wörld &amp; 123 it's"]
    for p in prompts:
        assert PW.parse_prompt_attention(p) == RL.parse_prompt_attention(p), p
        assert mt.encode(p) == rt.encode(p), p
    for n_ti in (0, 2):
        a = PW.tokens_and_weights(mt, prompts, 300, n_ti)
        b = RH.quiet(RL.get_prompts_with_weights, rt, prompts, 300, n_ti)
        assert a == b
    emb_fn = lambda x: TO.text_embed(sds["text_encoder"], x[0], x[1])  # noqa: E731
    enc_fn = lambda e: TO.encode_embedded(sds["text_encoder"], e, -1)  # noqa: E731
    ti = G.ti_embedding()[None]
    for prompt, emb, n in ((G.PROMPT, None, 0), ([G.LONG_PROMPT, "short"], None, 0), (G.PROMPT, ti, 3), (G.TRUNCATED_PROMPT, None, 0)):
        mine = RH.quiet(PW.weighted_text_embeddings, mt, emb_fn, enc_fn, prompt, embedding=emb, embedding_tokens_count=n)
        theirs = RH.quiet(RL.get_weighted_text_embeddings, rt, type("M", (), {"predict_on_batch": staticmethod(emb_fn)}),
                          type("M", (), {"predict_on_batch": staticmethod(enc_fn)}), prompt, model_max_length=77, embedding=emb,
                          embedding_tokens_count=n, pad_token_id=49407)
        assert mine.shape == theirs.shape and np.array_equal(mine, theirs)


@needs_ref
def test_lora_deltas_equal_reference(tmp_path):
    RH.import_reference()
    from stable_diffusion.ckpt_loader import load_weights_from_lora
    from minsdtf_b200 import lora
    path = str(tmp_path / "lora.safetensors")
    synth.save_safetensors(synth.make_lora_state_dict(), path)
    te_ref, un_ref = RH.quiet(load_weights_from_lora, path)
    te, un = lora.load_lora(path)
    assert set(te) == set(te_ref) and set(un) == set(un_ref) and len(un) > 40 and len(te) == 12
    for mine, theirs in ((te, te_ref), (un, un_ref)):
        for k in mine:
            assert mine[k].shape == theirs[k].shape and np.allclose(mine[k], theirs[k], rtol=1e-6, atol=1e-8), k
