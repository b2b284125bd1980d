"""Tokenizer parity: the C++ WordPiece / CLIP-BPE tokenizers behind the C ABI against
oracle/tokenizers.py — HF `tokenizers` itself for WordPiece, the restated open_clip SimpleTokenizer for CLIP BPE.
Host-side integer work: ids must be identical.  Runs without a GPU."""
import html

import numpy as np
import pytest

ASCII_WORDS = ["the", "quick", "brown", "fox", "jump", "over", "lazy", "dog", "hello", "world", "play", "un", "aff",
               "able", "token", "search", "vector", "marqo", "is", "it", "s", "t", "re", "ve", "m", "ll", "d", "cafe",
               "naive", "uber", "istanbul", "resume", "e", "a", "o", "u", "i", "n", "c", "x", "y", "z", "b"]
PIECES = ["##s", "##ing", "##ed", "##er", "##able", "##aff", "##ly", "##e", "##a", "##b", "##x", "##n", "##1", "##2"]
PUNCT = list("!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~") + ["—", "“", "”", "«", "»", "。", "・", "§"]
CJK = ["中", "文", "日", "本", "語", "\U00020000", "あ", "ア", "ᄀ", "ᅡ", "ᆨ"]
DIGITS = list("0123456789") + ["##0", "##3"]
VOCAB = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + ASCII_WORDS + PIECES + PUNCT + CJK + DIGITS + \
        ["σ", "α", "##σ", "ß", "क", "ก", "##ำ", "ı"]


def _vocab_bytes(words):
    return ("\n".join(words) + "\n").encode("utf-8")


@pytest.fixture(scope="module")
def wp(native_lib):
    from marqo_b200.tokenizers import WordPieceTokenizer
    from oracle import tokenizers as OT
    return WordPieceTokenizer(_vocab_bytes(VOCAB)), OT.bert_wordpiece(VOCAB), OT


def _same_wordpiece(wp, texts, max_length=64):
    mine, ref, OT = wp
    got = mine(list(texts), padding=True, truncation=True, max_length=max_length)
    ids, mask = OT.bert_encode_batch(ref, texts, max_length)
    np.testing.assert_array_equal(got["input_ids"], ids)
    np.testing.assert_array_equal(got["attention_mask"], mask)
    assert not got["token_type_ids"].any()


def test_wordpiece_handpicked(wp):
    _same_wordpiece(wp, [
        "The quick brown fox jumps over the lazy dog.",
        "Hello, WORLD!  unaffable playing players",
        "café naïve Über résumé İstanbul İ ẞ ΣΑΣ",     # accents, dotted I, sharp s, sigma
        "中文日本語hello\U00020000x あア 각",                                  # CJK spacing, kana, Hangul NFD
        "tab\tnew\nline\rcr nbsp em　ideographic lsnel",
        "zero​width soft­hyphen bell\x07 del\x7f null\x00 repl� pua unassigned͸",
        "hello [SEP] world [MASK] [UNK] [PAD] [CLS] [sep] [ SEP ] hello[SEP]world [MASK]hello",
        "it's don't we're I've I'm you'll he'd $5+3=8 a_b <a> `x` ~^|",
        "x" * 100, "x" * 101, "hello" + "x" * 96, "", " ", "กำ क़ क़ ゙か",
        "a" * 500 + " b " * 200,
    ])


def test_wordpiece_truncation_and_padding(wp):
    texts = ["the " * 40, "the", "", "hello world", "quick " * 7]
    for L in (2, 3, 5, 8, 16, 512):
        _same_wordpiece(wp, texts, max_length=L)
    mine = wp[0]
    one = mine("hello world", padding=True, truncation=True, max_length=8)      # str == [str]
    assert one["input_ids"].shape == (1, 4)
    with pytest.raises(ValueError):
        mine(["x"], padding=False)


def test_wordpiece_random_sentences(wp):
    rng = np.random.default_rng(0)
    alphabet = ASCII_WORDS + [w.upper() for w in ASCII_WORDS[:10]] + ["ing", "ed", "s", "1", "23", "4567"] + \
        [p for p in PUNCT] + CJK + [" ", "  ", "\t", " ", "​", "́", "̈", "é", "Å", "Å",
                                    "[SEP]", "[MASK]", "Σ", "İ", "\x01", "﻿"]
    texts = []
    for _ in range(400):
        n = int(rng.integers(0, 40))
        parts = [alphabet[int(i)] for i in rng.integers(0, len(alphabet), size=n)]
        seps = [" " if rng.random() < 0.6 else "" for _ in parts]
        texts.append("".join(p + s for p, s in zip(parts, seps)))
    _same_wordpiece(wp, texts, max_length=48)


def test_wordpiece_every_code_point_class(wp):
    """'a' + chr(cp) + 'b' over the BMP and the supplementary planes that carry text: the id pattern shows whether
    the code point was removed (a ##b), a space (a b), punctuation (a [UNK] b), CJK-spaced, decomposed to an ASCII base
    letter, or an ordinary letter ([UNK])."""
    cps = [cp for cp in range(1, 0x110000) if not (0xD800 <= cp <= 0xDFFF)]
    mine, ref, OT = wp
    mismatches = []
    for i in range(0, len(cps), 50000):
        chunk = cps[i:i + 50000]
        texts = ["a" + chr(cp) + "b" for cp in chunk]
        got = mine(texts, padding=True, truncation=True, max_length=16)["input_ids"]
        want, _ = OT.bert_encode_batch(ref, texts, 16)
        if got.shape != want.shape:
            w = max(got.shape[1], want.shape[1])
            got = np.pad(got, ((0, 0), (0, w - got.shape[1])))
            want = np.pad(want, ((0, 0), (0, w - want.shape[1])))
        bad = np.nonzero((got != want).any(axis=1))[0]
        mismatches.extend(chunk[int(b)] for b in bad)
    # the property tables are probed out of the `tokenizers` library itself (tools/gen_unicode_tables.py), so every
    # code point must agree, assigned or not
    assert not mismatches, [hex(m) for m in mismatches[:80]]


def test_wordpiece_cased(native_lib):
    from marqo_b200.tokenizers import WordPieceTokenizer
    from oracle import tokenizers as OT
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "Hello", "hello", "World", "##s", "café", "cafe", "É"]
    mine, ref = WordPieceTokenizer(_vocab_bytes(vocab), do_lower_case=False), OT.bert_wordpiece(vocab, lowercase=False)
    texts = ["Hello Worlds hello world café cafe É É"]
    ids, mask = OT.bert_encode_batch(ref, texts, 32)
    np.testing.assert_array_equal(mine(texts, max_length=32)["input_ids"], ids)


def test_wordpiece_requires_special_tokens(native_lib):
    from marqo_b200 import _native as N
    from marqo_b200.tokenizers import WordPieceTokenizer
    with pytest.raises(N.NativeError):
        WordPieceTokenizer(b"hello\nworld\n")


# ---------------------------------------------------------------------------------------------------- CLIP BPE
CORPUS = [
    "a photo of a cat", "a photo of a dog", "the quick brown fox jumps over the lazy dog", "hello world hello there",
    "it's a beautiful day, isn't it? we're here; they've gone. i'm sure you'll like what he'd done",
    "café naïve résumé über straße", "中文 日本語 こんにちは", "price: $12.50 (50% off!!!) #sale @store",
    "1234567890 3.14159 2024-09-22", "emoji \U0001f600\U0001f680 mixed—dash “quotes”", "photo photos photograph photographer",
] * 3


@pytest.fixture(scope="module")
def clip(native_lib):
    from marqo_b200.tokenizers import ClipBpeTokenizer
    from oracle import tokenizers as OT
    merges = OT.train_toy_merges(CORPUS, 400)
    return ClipBpeTokenizer(merges.encode("utf-8")), OT.SimpleTokenizerOracle(merges), OT


def test_clip_vocab_layout(clip):
    mine, ref, _ = clip
    assert mine.vocab_size == 512 + len(ref.bpe_ranks) + 2 and ref.eot == mine.vocab_size - 1
    out = mine([""])
    assert out.shape == (1, 77) and out[0, 0] == ref.sot and out[0, 1] == ref.eot and not out[0, 2:].any()


def test_clip_handpicked(clip):
    mine, ref, _ = clip
    texts = CORPUS[:11] + [
        "A PHOTO of a CAT", "  leading and trailing   \n\t spaces  　 ", "it's 'tis 'twas don't 'S 'RE x'll",
        "&amp; &lt;b&gt;bold&lt;/b&gt; &amp;amp; &#38; &#x26; &#X41; &#65 &notit; &ampfoo &unknown; &; &# &#x; & a",
        "&#0; &#x80; &#x9f; &#xD800; &#1114112; &#x1F600; &#11; &#xfdd0; &#99999999999999999999;",
        "&quot;quoted&quot; &apos;single&apos; &nbsp;nbsp &NotEqualTilde; &nvlt; &bne;",
        "<start_of_text> inside <end_of_text> text <START_OF_TEXT>", "İstanbul ǅ ẞ Ω K K",
        "a\x1cb\x1dc\x1ed\x1fe \x1c", "word " * 200, "supercalifragilisticexpialidocious" * 5, "1234567890" * 12,
        "ſtop 'k K",
    ]
    np.testing.assert_array_equal(mine(texts), ref(texts))
    np.testing.assert_array_equal(mine(texts, context_length=16), ref(texts, context_length=16))
    np.testing.assert_array_equal(mine("a photo of a cat"), ref(["a photo of a cat"]))       # str == [str]


def test_clip_random_strings(clip):
    mine, ref, _ = clip
    rng = np.random.default_rng(1)
    alphabet = list("abcdefghijklmnopqrstuvwxyzABCXYZ0123456789 .,!?'\"-_()&#;:/\\\n\t") + \
        ["photo", "cat", "the", "'s", "'ll", "&amp;", "&#39;", "&lt;", "é", "ü", "中", "文", "\U0001f600", "—", " ",
         "́", "٣", "Ⅷ", "½", "İ", "ẞ", "<end_of_text>"]
    texts = []
    for _ in range(500):
        n = int(rng.integers(0, 60))
        texts.append("".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), size=n)))
    np.testing.assert_array_equal(mine(texts), ref(texts))


def test_clip_every_code_point_class(clip):
    """'a' + chr(cp) + '1' + chr(cp) over all planes with text: pins \\p{L} / \\p{N} / \\s / lower-casing per code point."""
    mine, ref, _ = clip
    cps = [cp for cp in range(1, 0x110000) if not (0xD800 <= cp <= 0xDFFF)]
    for i in range(0, len(cps), 20000):
        texts = ["a" + chr(cp) + "1" + chr(cp) for cp in cps[i:i + 20000]]
        got, want = mine(texts, context_length=24), ref(texts, context_length=24)
        bad = np.nonzero((got != want).any(axis=1))[0]
        # Greek capital sigma is context-sensitive in str.lower() (final sigma); the C++ path lower-cases per code point
        bad = [cps[i + int(b)] for b in bad if cps[i + int(b)] != 0x3A3]
        assert not bad, [hex(b) for b in bad[:40]]


def test_html_unescape_matches_cpython(clip):
    """html.unescape is applied twice before anything else: check it in isolation through single-symbol outputs."""
    mine, ref, _ = clip
    names = sorted(html.entities.html5)
    texts = ["x&" + n + "y" for n in names] + ["x&" + n.rstrip(";") + "zz;y" for n in names[::7]]
    np.testing.assert_array_equal(mine(texts, context_length=32), ref(texts, context_length=32))


def test_clip_merges_validation(native_lib):
    from marqo_b200 import _native as N
    from marqo_b200.tokenizers import ClipBpeTokenizer
    with pytest.raises(N.NativeError):
        ClipBpeTokenizer(b"#version\nonlyone\n")
    t = ClipBpeTokenizer(b"#version\n")                       # no merges: bytes only
    assert t.vocab_size == 514
    assert t(["ab"])[0, :4].tolist()[0] == 512


def test_clip_oracle_agrees_with_hf_cliptokenizer(clip, tmp_path):
    """Pins the restated SimpleTokenizer against an independent implementation of the same published algorithm:
    transformers' (slow) CLIPTokenizer on the same vocabulary, for text both clean the same way."""
    import json
    transformers = pytest.importorskip("transformers")
    if not hasattr(transformers, "CLIPTokenizer"):
        pytest.skip("this transformers build has no slow CLIPTokenizer")
    mine, ref, OT = clip
    vocab = dict(ref.encoder)
    vocab["<|startoftext|>"] = vocab.pop("<start_of_text>")
    vocab["<|endoftext|>"] = vocab.pop("<end_of_text>")
    (tmp_path / "vocab.json").write_text(json.dumps(vocab))
    (tmp_path / "merges.txt").write_text(OT.train_toy_merges(CORPUS, 400))
    try:
        hf = transformers.CLIPTokenizer(str(tmp_path / "vocab.json"), str(tmp_path / "merges.txt"))
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"CLIPTokenizer cannot be built offline here: {e!r}")
    for t in ["a photo of a cat", "it's a beautiful day, isn't it?", "price: $12.50 (50% off!!!)",
              "photographer photos 123", "the quick brown fox jumps over the lazy dog"]:
        want = hf(t)["input_ids"]
        assert [ref.sot] + ref.encode(t) + [ref.eot] == want
        assert mine([t])[0, :len(want)].tolist() == want


def test_tokenizer_handles_are_thread_safe(wp, clip):
    """Marqo calls vectorise from up to 16 request threads with no lock: one immutable handle, many callers."""
    from concurrent.futures import ThreadPoolExecutor
    mine_wp, ref_wp, OT = wp
    mine_bpe, ref_bpe, _ = clip
    texts = [f"hello world {i} the quick brown fox's photo #{i} café" for i in range(64)]
    want_wp = mine_wp(texts, max_length=32)["input_ids"]
    want_bpe = mine_bpe(texts)

    def work(_):
        a = mine_wp(texts, max_length=32)["input_ids"]
        b = mine_bpe(texts)
        return np.array_equal(a, want_wp) and np.array_equal(b, want_bpe)

    with ThreadPoolExecutor(8) as pool:
        assert all(pool.map(work, range(32)))


def test_property_random_unicode_text(wp, clip):
    """hypothesis: arbitrary Unicode strings (any plane, any category, surrogates excluded) tokenise identically."""
    from hypothesis import given, settings, strategies as st
    mine_wp, ref_wp, OT = wp
    mine_bpe, ref_bpe, _ = clip
    piece = st.one_of(st.characters(blacklist_categories=("Cs",)),
                      st.sampled_from(list(" \t\n.,!?'&#;[]<>_-") + ["[SEP]", "[MASK]", "&amp;", "&#x41;", "'s", "Σ"]))
    text = st.lists(piece, max_size=40).map("".join)

    @settings(max_examples=int(__import__('os').environ.get('TOKENIZER_FUZZ_EXAMPLES', '300')), deadline=None, derandomize=True)
    @given(st.lists(text, min_size=1, max_size=6))
    def check(texts):
        got = mine_wp(texts, max_length=24)
        ids, mask = OT.bert_encode_batch(ref_wp, texts, 24)
        np.testing.assert_array_equal(got["input_ids"], ids)
        np.testing.assert_array_equal(got["attention_mask"], mask)
        np.testing.assert_array_equal(mine_bpe(texts, context_length=20), ref_bpe(texts, context_length=20))

    check()
