"""Prompt attention syntax and long prompts — behaviour of the reference's
`stable_diffusion/long_prompt_weighting.py`:

* `parse_prompt_attention` (:26-109): `(text)` multiplies the attention of `text` by 1.1, `[text]` divides it by 1.1,
  `(text:1.3)` sets an explicit multiplier, brackets nest multiplicatively, `\\(` `\\)` `\\[` `\\]` `\\\\` are literals,
  unbalanced opening brackets apply to the end of the prompt, neighbouring runs of equal weight are merged;
* `tokens_and_weights` (:112-150): every run is tokenised on its own (stripped), the run's weight is repeated per token,
  textual-inversion placeholders (`*` tokens) go first, at most 75 * 4 tokens are kept;
* `pad_tokens_and_weights` (:153-176, no_boseos_middle=False): <start> tokens <pad...> <end> to 75 m + 2 entries; the
  weights get a 1.0 around every 75-token window and 1.0 padding to 77 m entries;
* `weighted_text_embeddings` (:179-333): windows `[75 i, 75 i + 77)` of the padded sequence, first and last entry
  overwritten by <start> / <end>, are encoded one by one and concatenated to (B, 77 m, 768); textual-inversion vectors
  replace embedding rows 1..n of the first window between the embedding lookup and the encoder; the result is
  multiplied by the per-token weights and rescaled so that each prompt keeps its mean.

The encoders are passed in as the two callables the reference uses (`text_clip_embedding.predict_on_batch([tokens,
positions])` and `text_encoder.predict_on_batch(embedding)`), so the engine-backed models and the oracle plug in alike.
"""
from __future__ import annotations

import re

import numpy as np

UP, DOWN = 1.1, 1 / 1.1
WINDOW = 77          # what the text encoder takes
BODY = WINDOW - 2    # prompt tokens per window
MAX_WINDOWS = 4

# one token of the attention syntax per match: an escaped character, a bracket, an explicit ":weight)", plain text, or a
# stray colon
_LEX = re.compile(r"\\[()\[\]\\]|\\|[(\[]|:([+-]?[.\d]+)\)|[)\]]|[^\\()\[\]:]+|:")


def parse_prompt_attention(text: str) -> list:
    """-> [[text, weight], ...] (long_prompt_weighting.py:26-109)."""
    runs = []          # [text, weight]
    open_round, open_square = [], []   # index of the first run each open bracket covers

    def scale_from(start, factor):
        for r in runs[start:]:
            r[1] *= factor

    for m in _LEX.finditer(text):
        piece, explicit = m.group(0), m.group(1)
        if piece[0] == "\\":
            runs.append([piece[1:], 1.0])
        elif piece == "(":
            open_round.append(len(runs))
        elif piece == "[":
            open_square.append(len(runs))
        elif explicit is not None and open_round:
            scale_from(open_round.pop(), float(explicit))
        elif piece == ")" and open_round:
            scale_from(open_round.pop(), UP)
        elif piece == "]" and open_square:
            scale_from(open_square.pop(), DOWN)
        else:
            runs.append([piece, 1.0])
    for start in open_round:
        scale_from(start, UP)
    for start in open_square:
        scale_from(start, DOWN)
    if not runs:
        return [["", 1.0]]
    merged = [runs[0]]
    for t, w in runs[1:]:
        if w == merged[-1][1]:
            merged[-1][0] += t
        else:
            merged.append([t, w])
    return merged


def tokens_and_weights(tokenizer, prompts, limit, embedding_tokens_count=0, embedding_tokens_weight=1.0):
    """Per prompt: token ids (no <start>/<end>) and one weight per token, cut at `limit` (:112-150)."""
    all_tokens, all_weights, cut = [], [], False
    for text in prompts:
        toks, wts = [], []
        if embedding_tokens_count > 0:
            toks += list(tokenizer.encode("*")[1:-1]) * embedding_tokens_count
            wts += [embedding_tokens_weight] * embedding_tokens_count
        for run, weight in parse_prompt_attention(text):
            ids = list(tokenizer.encode(run.strip())[1:-1])
            toks += ids
            wts += [weight] * len(ids)
            if len(toks) > limit:
                break
        if len(toks) > limit:
            cut = True
            toks, wts = toks[:limit], wts[:limit]
        all_tokens.append(toks)
        all_weights.append(wts)
    if cut:
        print("Prompt was truncated. Try to shorten the prompt or increase max_embeddings_multiples")
    return all_tokens, all_weights


def pad_tokens_and_weights(tokens, weights, windows, bos, eos, pad):
    """-> (int32 (B, 75 m + 2), float32 (B, 77 m)) (:153-176 with no_boseos_middle=False)."""
    length = BODY * windows + 2
    out_t, out_w = [], []
    for toks, wts in zip(tokens, weights):
        out_t.append([bos] + toks + [pad] * (length - 2 - len(toks)) + [eos])
        w = []
        if wts:
            for j in range(windows):
                w += [1.0] + wts[j * BODY:(j + 1) * BODY] + [1.0]
        w += [1.0] * (windows * WINDOW - len(w))
        out_w.append(w)
    return np.asarray(out_t, np.int32), np.asarray(out_w, np.float32)


def encode_windows(embed_fn, encode_fn, padded_tokens, embedding=None, embedding_tokens_count=0):
    """(B, 75 m + 2) ids -> (B, 77 m, 768) (:179-237 with no_boseos_middle=False)."""
    windows = (padded_tokens.shape[1] - 2) // BODY
    use_embedding = embedding_tokens_count > 0 and embedding is not None
    positions = np.arange(WINDOW, dtype=np.int32)[None]
    outs = []
    for i in range(windows):
        chunk = padded_tokens[:, i * BODY:i * BODY + WINDOW].copy()
        if windows > 1:
            chunk[:, 0] = padded_tokens[0, 0]
            chunk[:, -1] = padded_tokens[0, -1]
        emb = np.asarray(embed_fn([chunk, positions]))
        if use_embedding and i == 0:
            n = embedding_tokens_count
            emb = np.concatenate([emb[:, :1], np.tile(embedding, (emb.shape[0], 1, 1)).astype(emb.dtype), emb[:, n + 1:]], axis=1)
        outs.append(np.asarray(encode_fn(emb)))
    return outs[0] if windows == 1 else np.concatenate(outs, axis=1)


def weighted_text_embeddings(tokenizer, embed_fn, encode_fn, prompt, pad_token_id=49407, embedding=None,
                             embedding_tokens_count=0, embedding_tokens_weight=1.0):
    """get_weighted_text_embeddings(...) with the reference's defaults (:240-333): up to 4 windows, <start>/<end> kept in
    every window, brackets parsed, weights applied with the mean preserved."""
    if embedding_tokens_count > 0 and embedding is None:
        embedding_tokens_count = 0
    prompts = [prompt] if isinstance(prompt, str) else list(prompt)
    tokens, weights = tokens_and_weights(tokenizer, prompts, BODY * MAX_WINDOWS, embedding_tokens_count, embedding_tokens_weight)
    longest = max(len(t) for t in tokens)
    windows = max(1, min(MAX_WINDOWS, (longest - 1) // BODY + 1))
    bos = getattr(tokenizer, "start_of_text", None)
    eos = getattr(tokenizer, "end_of_text", None)
    padded, w = pad_tokens_and_weights(tokens, weights, windows, bos, eos, pad_token_id)
    ctx = np.array(encode_windows(embed_fn, encode_fn, padded, embedding, embedding_tokens_count), copy=True)
    w = w.astype(ctx.dtype)
    before = ctx.mean(axis=(-2, -1))
    ctx *= w[:, :, None]
    ctx *= (before / ctx.mean(axis=(-2, -1)))[:, None, None]
    return ctx
