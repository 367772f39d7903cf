"""CLIP BPE tokenizer (minsdtf_b200/bpe.py) on a synthetic vocabulary file (the published one is a download): vocabulary
layout, merge order, word boundaries, cleaning, special tokens."""
import gzip

import pytest

from minsdtf_b200.bpe import ClipBPE, byte_symbols


@pytest.fixture()
def tok(tmp_path):
    rules = ["h e", "l l", "he ll", "hell o</w>", "w o", "r l", "wo rl", "worl d</w>", "a b", "ab c</w>"]
    p = tmp_path / "vocab.txt.gz"
    with gzip.open(p, "wb") as f:
        f.write(("#version: synthetic\n" + "\n".join(rules) + "\n").encode())
    return ClipBPE(str(p)), rules


def test_byte_table_is_a_bijection_onto_printable_characters():
    sym = byte_symbols()
    assert len(sym) == 256 and len(set(sym)) == 256
    assert sym[ord("a")] == "a" and sym[ord("!")] == "!" and sym[0xFF] == "\xff"
    assert sym[0] == chr(256) and sym[ord(" ")] == chr(256 + 32) and all(not c.isspace() for c in sym)


def test_vocabulary_layout(tok):
    t, rules = tok
    assert t.ids["!"] == 0 and t.ids["!</w>"] == 256  # 256 byte symbols, then the same with the end-of-word mark
    assert t.ids["he"] == 512 and t.ids["hello</w>"] == 512 + 3  # one id per merge rule, in file order
    assert t.start_id == 512 + len(rules) and t.end_id == t.start_id + 1


def test_merges_apply_in_rank_order_and_respect_word_ends(tok):
    t, _ = tok
    ids = t.encode("Hello  WORLD")
    assert ids == [t.start_id, t.ids["hello</w>"], t.ids["world</w>"], t.end_id]
    # 'o' inside a word is not 'o</w>': "hellos" cannot use the hell+o</w> rule
    assert t.encode("hellos")[1:-1] == [t.ids["hell"], t.ids["o"], t.ids["s</w>"]]
    # lowest rank first: "abc" -> (a b) then (ab c</w>)
    assert t.encode("abc")[1:-1] == [t.ids["abc</w>"]]
    # digits are split one by one, punctuation runs stay together
    assert t.encode("11")[1:-1] == [t.ids["1</w>"], t.ids["1</w>"]]
    assert t.encode("!!")[1:-1] == [t.ids["!"], t.ids["!</w>"]]


def test_cleaning_and_special_tokens(tok):
    t, _ = tok
    assert t.encode("  hello&amp;amp;  ")[1:-1] == [t.ids["hello</w>"], t.ids["&</w>"]]
    assert t.encode("") == [t.start_id, t.end_id]
    assert t.encode("<|endoftext|>")[1:-1] == [t.end_id]
    # non-ASCII text goes through its UTF-8 bytes
    e = t.encode("é")[1:-1]
    assert e == [t.ids[byte_symbols()[0xC3]], t.ids[byte_symbols()[0xA9] + "</w>"]]


def test_matches_transformers_clip_tokenizer_on_the_same_vocabulary(tmp_path):
    """independent implementation of the same published algorithm: transformers.CLIPTokenizer built from the same merges"""
    transformers = pytest.importorskip("transformers")
    import json
    rules = ["h e", "l l", "he ll", "hell o</w>", "w o", "r l", "wo rl", "worl d</w>", "a b", "ab c</w>", "t h", "th e</w>", "c a",
             "ca t</w>", "i n", "in g</w>", "' s</w>"]
    gz = tmp_path / "v.txt.gz"
    with gzip.open(gz, "wb") as f:
        f.write(("#version\n" + "\n".join(rules) + "\n").encode())
    t = ClipBPE(str(gz))
    (tmp_path / "vocab.json").write_text(json.dumps(t.ids))
    (tmp_path / "merges.txt").write_text("#version: 0.2\n" + "\n".join(rules) + "\n")
    hf = transformers.CLIPTokenizer(str(tmp_path / "vocab.json"), str(tmp_path / "merges.txt"))
    for s in ["Hello world", "the cat abc hellos", "hello, world!! 123", "  The   CAT's  ", "é ü", "sing-ing in the (rain)", ""]:
        assert t.encode(s) == hf(s)["input_ids"], s
