"""CLIP text tower on the engine against the oracle restatement of text_encoder.py (itself pinned
to transformers.CLIPTextModel in tests/test_cpu_text_oracle.py), and the text -> image path through the public API."""
import numpy as np
import pytest

from minsdtf_b200 import synth
from oracle import text_oracle as TO

pytestmark = pytest.mark.gpu
TEXT_BAR = 1e-2  # max |d| / max |ref| of the (B,77,768) context: bf16 GEMM operands, fp32 residual stream, vs the fp32 oracle


@pytest.fixture(scope="module")
def text_sd():
    return synth.make_state_dict("text_encoder")


@pytest.fixture(scope="module")
def engine_text(engine, text_sd):
    if "text_encoder" not in engine.loaded:
        engine.load_state_dict(text_sd, "text_encoder")
    return engine


@pytest.mark.parametrize("B,T,clip_skip", [(1, 77, -1), (3, 77, -2), (2, 77, -12), (2, 40, -1), (16, 77, -1)])
def test_text_encoder_parity(engine_text, text_sd, B, T, clip_skip):
    tokens = synth.prompt_tokens(B)[:, :T]
    got = engine_text.text_encode(tokens, clip_skip)
    ref = TO.text_encode(text_sd, tokens, clip_skip)
    err = float(np.abs(got - ref).max() / np.abs(ref).max())
    print(f"text tower B={B} T={T} clip_skip={clip_skip}: rel err {err:.4g}")
    assert got.shape == (B, T, 768) and np.isfinite(got).all()
    assert err <= TEXT_BAR, err


def test_text_encoder_is_causal_and_batch_invariant(engine_text):
    """a token's encoding depends on the tokens before it only (text_encoder.py:78-81), and not on its batch neighbours"""
    tok = synth.prompt_tokens(2)
    a = engine_text.text_encode(tok, -1)
    tok2 = tok.copy()
    tok2[0, 40:] = 1234
    b = engine_text.text_encode(tok2, -1)
    assert np.array_equal(a[0, :40], b[0, :40]) and not np.array_equal(a[0, 40:], b[0, 40:])
    assert np.array_equal(a[1], engine_text.text_encode(tok[1:], -1)[0])


def test_rejects_bad_tokens(engine_text):
    from minsdtf_b200.engine import EngineError
    with pytest.raises(EngineError):
        engine_text.text_encode(np.zeros((1, 78), np.int32), -1)  # longer than the position table
    with pytest.raises(EngineError):
        engine_text.text_encode(np.zeros((1, 77), np.int32), 0)   # clip_skip out of range


def test_token_ids_to_image_through_public_api(engine_text, engine_unet, engine_vae):
    """text_to_image-style call with token ids: tokens -> text tower -> denoise -> decode, all on the engine"""
    from minsdtf_b200.stable_diffusion import StableDiffusion
    sd = StableDiffusion(img_height=128, img_width=128, synthetic=True, engine=engine_text)
    tokens = synth.prompt_tokens(1)[0]
    ctx = sd.encode_text(tokens)
    assert ctx.shape == (77, 768)
    img = sd.generate_image(ctx, batch_size=1, num_steps=3, unconditional_guidance_scale=7.5,
                            diffusion_noise=synth.latents(1, 16, 16), guidance_rescale=0.7)
    assert img.shape == (1, 128, 128, 3) and img.dtype == np.uint8
    unc = sd._get_unconditional_context()
    assert unc.shape == (1, 77, 768)
    # the reference's two-model composition (stable_diffusion.py:700-725) gives the same context
    two = sd.text_encoder.predict_on_batch(sd.text_clip_embedding.predict_on_batch([tokens[None], sd._get_pos_ids()]))
    assert np.array_equal(two[0], ctx)


def test_long_prompt_windows_and_string_prompts(engine_text, engine_unet, engine_vae, tmp_path):
    """154 token ids = two 77-token windows encoded separately and concatenated; a string goes through model.tokenizer"""
    import gzip
    from minsdtf_b200.bpe import ClipBPE
    from minsdtf_b200.stable_diffusion import StableDiffusion
    sd = StableDiffusion(img_height=128, img_width=128, synthetic=True, engine=engine_text)
    tok = synth.prompt_tokens(2)
    long_ids = np.concatenate([tok[0], tok[1]])
    ctx = sd.encode_text(long_ids)
    assert ctx.shape == (154, 768)
    assert np.array_equal(ctx[:77], sd.encode_text(tok[0])) and np.array_equal(ctx[77:], sd.encode_text(tok[1]))
    # guidance on: the 77-token unconditional context is continued with an empty-prompt window to 154 tokens
    img = sd.generate_image(ctx, batch_size=1, num_steps=2, diffusion_noise=synth.latents(1, 16, 16), unconditional_guidance_scale=7.5)
    assert img.shape == (1, 128, 128, 3) and img.dtype == np.uint8
    gz = tmp_path / "v.txt.gz"
    with gzip.open(gz, "wb") as f:
        f.write(b"#version\nh e\nl l\nhe ll\nhell o</w>\n")
    sd.tokenizer = ClipBPE(str(gz))
    c = sd.encode_text("hello hello")
    ids = sd.tokenizer.encode("hello hello")
    assert c.shape == (77, 768) and len(ids) == 4
