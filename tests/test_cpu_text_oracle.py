"""The CLIP text-tower oracle (oracle/text_oracle.py, a restatement of the reference's text_encoder.py) against an
independent implementation of the same architecture: transformers.CLIPTextModel with the same seeded weights."""
import numpy as np
import pytest
import torch

from minsdtf_b200 import keys as K
from minsdtf_b200 import synth
from oracle import text_oracle as T


def test_text_encoder_key_table():
    keys = K.text_encoder_keys()
    assert len(keys) == 2 + 12 * 16 + 2  # the reference's two mappings: text_encoder.py:110-111 and :137-157
    assert K.n_params(keys) == 123_060_480  # SD1.5 CLIP ViT-L/14 text tower


@pytest.mark.parametrize("clip_skip", [-1, -2])
def test_oracle_matches_transformers_clip(clip_skip):
    transformers = pytest.importorskip("transformers")
    sd = synth.make_state_dict("text_encoder")
    cfg = transformers.CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                                      num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu",
                                      layer_norm_eps=1e-5, eos_token_id=49407, bos_token_id=49406, pad_token_id=49407)
    model = transformers.CLIPTextModel(cfg).eval()
    model_sd = model.state_dict()
    missing = [k for k in sd if k not in model_sd]
    assert not missing, missing[:4]
    model.load_state_dict({k: v for k, v in sd.items()}, strict=False)
    tokens = synth.prompt_tokens(2)
    with torch.no_grad():
        out = model(input_ids=torch.as_tensor(tokens, dtype=torch.long), output_hidden_states=True)
        ref = out.last_hidden_state if clip_skip == -1 else model.text_model.final_layer_norm(out.hidden_states[clip_skip])
    got = T.text_encode(sd, tokens, clip_skip)
    err = np.abs(got - ref.numpy()).max()
    assert got.shape == (2, 77, 768) and err < 1e-4, err


def test_unconditional_tokens_shape():
    tok = synth.prompt_tokens(3)
    assert tok.shape == (3, 77) and tok.dtype == np.int32 and (tok[:, 0] == 49406).all() and (tok[:, -1] == 49407).all()


def test_oracle_matches_committed_transformers_golden():
    """the same check against the committed fixture (no transformers import needed where the tests run)"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "text_tower.npz"))
    sd = synth.make_state_dict("text_encoder")
    for skip, key in ((-1, "ctx_skip1"), (-2, "ctx_skip2")):
        got = T.text_encode(sd, g["tokens"], skip)[..., ::4]
        assert np.abs(got - g[key].astype(np.float32)).max() <= 2e-3 * float(g["max_abs"])  # float16 storage of the fixture
