"""CLIP byte-pair tokenizer: same vocabulary construction and merge semantics as the reference's
`clip_tokenizer.py:78-190` (OpenAI's published CLIP BPE), written against a user-supplied vocabulary file — the reference
downloads `bpe_simple_vocab_16e6.txt.gz` (`clip_tokenizer.py:79-82`), which is impossible offline.

    tok = ClipBPE("/path/to/bpe_simple_vocab_16e6.txt.gz")
    ids = tok.encode("a photo of an astronaut")          # [49406, ..., 49407]
    model.tokenizer = tok; model.text_to_image("a photo of an astronaut", ...)

Vocabulary ids: 256 byte symbols, the same 256 with the end-of-word mark, one id per merge rule (in file order), then
<|startoftext|> and <|endoftext|> — 49408 entries for the 48894 merges of the published file."""
from __future__ import annotations

import gzip
import html
from functools import lru_cache

import regex

EOW = "</w>"
START, END = "<|startoftext|>", "<|endoftext|>"
_SPLIT = regex.compile(r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+", regex.IGNORECASE)


@lru_cache(maxsize=1)
def byte_symbols() -> tuple:
    """one printable unicode character per byte value: printable latin-1 bytes map to themselves, the other 68 bytes to
    U+0100.. in increasing byte order (the published byte<->unicode table)"""
    keep = set(range(ord("!"), ord("~") + 1)) | set(range(0xA1, 0xAD)) | set(range(0xAE, 0x100))
    table, extra = [None] * 256, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + extra)
            extra += 1
    return tuple(table)


class ClipBPE:
    def __init__(self, vocab_path: str, max_merges: int = 49152 - 256 - 2):
        opener = gzip.open if str(vocab_path).endswith(".gz") else open
        with opener(vocab_path, "rb") as f:
            lines = f.read().decode("utf-8").split("\n")
        rules = [tuple(ln.split()) for ln in lines[1:1 + max_merges]]  # the first line is a header
        rules = [r for r in rules if len(r) == 2]
        sym = byte_symbols()
        # symbol order of the published vocabulary: printable bytes first ('!'..'~', 0xA1..0xAC, 0xAE..0xFF), then the rest
        ordered = [c for c in sym if ord(c) < 256] + [c for c in sym if ord(c) >= 256]
        vocab = ordered + [c + EOW for c in ordered] + [a + b for a, b in rules] + [START, END]
        self.ids = {tok: i for i, tok in enumerate(vocab)}
        self.rank = {rule: i for i, rule in enumerate(rules)}
        self.start_id, self.end_id = self.ids[START], self.ids[END]
        self._memo = {}

    # ------------------------------------------------------------------------------------------------------------
    def _merge_word(self, word: str) -> list:
        """greedy BPE: repeatedly fuse the adjacent pair with the lowest merge rank, everywhere it occurs, left to right"""
        hit = self._memo.get(word)
        if hit is not None:
            return hit
        parts = list(word[:-1]) + [word[-1] + EOW]
        while len(parts) > 1:
            best, best_rank = None, None
            for pair in zip(parts, parts[1:]):
                r = self.rank.get(pair)
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = pair, r
            if best is None:
                break
            fused, i = [], 0
            while i < len(parts):
                if i + 1 < len(parts) and parts[i] == best[0] and parts[i + 1] == best[1]:
                    fused.append(parts[i] + parts[i + 1])
                    i += 2
                else:
                    fused.append(parts[i])
                    i += 1
            parts = fused
        self._memo[word] = parts
        return parts

    # names the reference's SimpleTokenizer exposes (clip_tokenizer.py:120-126), used by the prompt-weighting code
    @property
    def start_of_text(self) -> int:
        return self.start_id

    @property
    def end_of_text(self) -> int:
        return self.end_id

    @staticmethod
    def clean(text: str) -> str:
        """html-unescape twice, collapse whitespace, lower-case (clip_tokenizer.py:66-75, 178-180; the optional ftfy
        pass of the original is not part of the reference either)"""
        text = html.unescape(html.unescape(text)).strip()
        return " ".join(text.split()).lower()

    def encode(self, text: str) -> list:
        """-> [<start>] + token ids + [<end>] (unpadded, untruncated: StableDiffusionBase.encode_text pads to 77)"""
        sym = byte_symbols()
        out = [self.start_id]
        for piece in _SPLIT.findall(self.clean(text)):
            if piece in (START, END):
                out.append(self.ids[piece])
                continue
            word = "".join(sym[b] for b in piece.encode("utf-8"))
            out.extend(self.ids[p] for p in self._merge_word(word))
        out.append(self.end_id)
        return out
