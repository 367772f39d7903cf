"""TEST INFRASTRUCTURE — CPU oracle, never imported by the product path.

torch-fp32 CPU restatement of the reference's CLIP text tower (stable_diffusion/text_encoder.py): TextClipEmbedding
(:106-122) followed by TextEncoder (:125-170).  Weights: {HF checkpoint key: torch tensor in PyTorch layout}, the names
of the reference's own ckpt mappings (:110-111, 137-157).

PINNED against an independent implementation: tests/test_cpu_text_oracle.py loads the same state dict into
`transformers.CLIPTextModel` (transformers 5.5, the architecture the SD1.5 text_encoder/model.safetensors was written
for) and requires agreement to 1e-4 for clip_skip -1 and -2.  (The reference's own Keras graph cannot be run offline.)
"""
import numpy as np
import torch
import torch.nn.functional as F

NUM_LAYERS, NUM_HEADS, DIM = 12, 12, 768


def quick_gelu(x):
    """text_encoder.py:102-103"""
    return x * torch.sigmoid(1.702 * x)


def _linear(sd, key, x):
    return F.linear(x, sd[key + ".weight"].float(), sd[key + ".bias"].float())


def clip_attention(sd, p, x):
    """CLIPAttention.call, text_encoder.py:76-99: q scaled AFTER its bias (:84), additive -inf mask above the diagonal
    (:78-81), softmax over keys, heads = 12, head_dim = 64."""
    B, T, C = x.shape
    d = C // NUM_HEADS
    q = _linear(sd, p + ".q_proj", x) * d ** -0.5
    k = _linear(sd, p + ".k_proj", x)
    v = _linear(sd, p + ".v_proj", x)
    q, k, v = (t.view(B, T, NUM_HEADS, d).transpose(1, 2) for t in (q, k, v))
    w = q @ k.transpose(-1, -2)
    mask = torch.triu(torch.full((T, T), float("-inf")), diagonal=1)
    w = torch.softmax(w + mask, dim=-1)
    o = (w @ v).transpose(1, 2).reshape(B, T, C)
    return _linear(sd, p + ".out_proj", o)


def encoder_layer(sd, p, x):
    """CLIPEncoderLayer.call, text_encoder.py:46-55"""
    h = F.layer_norm(x, (DIM,), sd[p + ".layer_norm1.weight"].float(), sd[p + ".layer_norm1.bias"].float(), eps=1e-5)
    x = x + clip_attention(sd, p + ".self_attn", h)
    h = F.layer_norm(x, (DIM,), sd[p + ".layer_norm2.weight"].float(), sd[p + ".layer_norm2.bias"].float(), eps=1e-5)
    h = quick_gelu(_linear(sd, p + ".mlp.fc1", h))
    return x + _linear(sd, p + ".mlp.fc2", h)


def text_embed(sd, tokens, positions=None):
    """TextClipEmbedding (text_encoder.py:22-33,106-122): token + position embedding -> (B,T,768)."""
    tokens = torch.as_tensor(np.asarray(tokens), dtype=torch.long)
    if tokens.ndim == 1:
        tokens = tokens[None]
    T = tokens.shape[1]
    pos = torch.arange(T)[None] if positions is None else torch.as_tensor(np.asarray(positions), dtype=torch.long)
    x = sd["text_model.embeddings.token_embedding.weight"].float()[tokens] + \
        sd["text_model.embeddings.position_embedding.weight"].float()[pos]
    return x.numpy().astype(np.float32)


def encode_embedded(sd, embedding, clip_skip=-1):
    """TextEncoder (text_encoder.py:125-135) on a (B,T,768) embedding: layers 0 .. 12+clip_skip, then the final
    LayerNorm of out[clip_skip] (:128-133)."""
    with torch.no_grad():
        x = torch.as_tensor(np.asarray(embedding), dtype=torch.float32)
        for i in range(NUM_LAYERS + clip_skip + 1):
            x = encoder_layer(sd, f"text_model.encoder.layers.{i}", x)
        x = F.layer_norm(x, (DIM,), sd["text_model.final_layer_norm.weight"].float(), sd["text_model.final_layer_norm.bias"].float(),
                         eps=1e-5)
    return x.numpy().astype(np.float32)


def text_encode(sd, tokens, clip_skip=-1):
    """tokens (B,T) int -> context (B,T,768) float32: TextEncoder(TextClipEmbedding([tokens, 0..T-1]))."""
    return encode_embedded(sd, text_embed(sd, tokens), clip_skip)
