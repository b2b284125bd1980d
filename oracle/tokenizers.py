"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by marqo_b200/).

Tokenizer checkers for SURVEY §8 f2.  Both tokenizers are third-party code the reference only CALLS
(hugging_face_model.py:125-130,179-185; open_clip_model.py:211-222,277-279):

* WordPiece: the checker is the real thing — HF `tokenizers` (installed in this image; the reference pins
  transformers==4.41.2 which depends on it) — instantiated on a synthetic vocabulary.  `bert_wordpiece()` builds
  it exactly as BertTokenizerFast does (BertNormalizer / BertPreTokenizer / WordPiece / "[CLS] A [SEP]").
* CLIP BPE: open_clip_torch 2.24.0 is NOT installed and not vendored, so `SimpleTokenizerOracle` restates its
  published `SimpleTokenizer` (open_clip/tokenizer.py) with the same third-party pieces it uses (`regex`, `html`);
  `ftfy.fix_text` is absent from the image and is left out on both sides (parity for ftfy-repairable mojibake:
  unpinned).  The cleaning functions are also in the reference tree: hf_tokenizer.py:9-17.
"""
from __future__ import annotations

import html
from typing import Dict, List, Sequence

import numpy as np
import regex as re


def bert_wordpiece(vocab_words: Sequence[str], lowercase: bool = True):
    from tokenizers import BertWordPieceTokenizer
    return BertWordPieceTokenizer(vocab={w: i for i, w in enumerate(vocab_words)}, lowercase=lowercase)


def bert_encode_batch(tok, texts: Sequence[str], max_length: int):
    """tokenizer(texts, padding=True, truncation=True, max_length=max_length) -> (ids, mask) int arrays."""
    tok.enable_truncation(max_length=max_length)
    tok.enable_padding(pad_id=tok.token_to_id("[PAD]"), pad_token="[PAD]")
    enc = tok.encode_batch(list(texts))
    return (np.asarray([e.ids for e in enc], dtype=np.int64), np.asarray([e.attention_mask for e in enc], dtype=np.int64))


# ------------------------------------------------------------------------------------------------ CLIP BPE
def bytes_to_unicode() -> Dict[int, str]:
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(2 ** 8):
        if b not in bs:
            bs.append(b)
            cs.append(2 ** 8 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


def _pairs(word):
    return set(zip(word[:-1], word[1:]))


def basic_clean(text: str) -> str:          # hf_tokenizer.py:14-17 without ftfy
    return html.unescape(html.unescape(text)).strip()


def whitespace_clean(text: str) -> str:     # hf_tokenizer.py:9-12
    return re.sub(r"\s+", " ", text).strip()


class SimpleTokenizerOracle:
    def __init__(self, merges_text: str, context_length: int = 77):
        self.byte_encoder = bytes_to_unicode()
        lines = merges_text.split("\n")
        lines = lines[1:49152 - 256 - 2 + 1]
        merges = [tuple(m.split()) for m in lines if m.strip()]     # empty lines ignored (the shipped file has none)
        vocab = list(bytes_to_unicode().values())
        vocab = vocab + [v + "</w>" for v in vocab]
        for m in merges:
            vocab.append("".join(m))
        vocab.extend(["<start_of_text>", "<end_of_text>"])
        self.encoder = dict(zip(vocab, range(len(vocab))))
        self.bpe_ranks = dict(zip(merges, range(len(merges))))
        self.cache = {"<start_of_text>": "<start_of_text>", "<end_of_text>": "<end_of_text>"}
        self.pat = re.compile(r"<start_of_text>|<end_of_text>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+",
                              re.IGNORECASE)
        self.sot = self.encoder["<start_of_text>"]
        self.eot = self.encoder["<end_of_text>"]
        self.context_length = context_length

    def bpe(self, token: str) -> str:
        if token in self.cache:
            return self.cache[token]
        word = tuple(token[:-1]) + (token[-1] + "</w>",)
        pairs = _pairs(word)
        if not pairs:
            return token + "</w>"
        while True:
            bigram = min(pairs, key=lambda p: self.bpe_ranks.get(p, float("inf")))
            if bigram not in self.bpe_ranks:
                break
            first, second = bigram
            new_word, i = [], 0
            while i < len(word):
                try:
                    j = word.index(first, i)
                    new_word.extend(word[i:j])
                    i = j
                except ValueError:
                    new_word.extend(word[i:])
                    break
                if word[i] == first and i < len(word) - 1 and word[i + 1] == second:
                    new_word.append(first + second)
                    i += 2
                else:
                    new_word.append(word[i])
                    i += 1
            word = tuple(new_word)
            if len(word) == 1:
                break
            pairs = _pairs(word)
        return " ".join(word)

    def encode(self, text: str) -> List[int]:
        out = []
        text = whitespace_clean(basic_clean(text)).lower()
        for token in re.findall(self.pat, text):
            token = "".join(self.byte_encoder[b] for b in token.encode("utf-8"))
            out.extend(self.encoder[t] for t in self.bpe(token).split(" "))
        return out

    def __call__(self, texts: Sequence[str], context_length: int = None) -> np.ndarray:
        L = context_length or self.context_length
        res = np.zeros((len(texts), L), dtype=np.int64)
        for i, t in enumerate(texts):
            toks = [self.sot] + self.encode(t) + [self.eot]
            if len(toks) > L:
                toks = toks[:L]
                toks[-1] = self.eot
            res[i, :len(toks)] = toks
        return res


def train_toy_merges(corpus: Sequence[str], n_merges: int) -> str:
    """A small BPE training run (standard greedy most-frequent-pair) -> the text of a merges file."""
    be = bytes_to_unicode()
    pat = re.compile(r"'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+", re.IGNORECASE)
    words: Dict[tuple, int] = {}
    for line in corpus:
        for tok in re.findall(pat, whitespace_clean(basic_clean(line)).lower()):
            sym = [be[b] for b in tok.encode("utf-8")]
            sym[-1] += "</w>"
            words[tuple(sym)] = words.get(tuple(sym), 0) + 1
    merges = []
    for _ in range(n_merges):
        counts: Dict[tuple, int] = {}
        for w, c in words.items():
            for p in zip(w[:-1], w[1:]):
                counts[p] = counts.get(p, 0) + c
        if not counts:
            break
        best = max(sorted(counts), key=lambda p: counts[p])
        merges.append(best)
        new_words = {}
        for w, c in words.items():
            nw, i = [], 0
            while i < len(w):
                if i + 1 < len(w) and (w[i], w[i + 1]) == best:
                    nw.append(w[i] + w[i + 1])
                    i += 2
                else:
                    nw.append(w[i])
                    i += 1
            new_words[tuple(nw)] = new_words.get(tuple(nw), 0) + c
        words = new_words
    return "#version: toy\n" + "\n".join(f"{a} {b}" for a, b in merges) + "\n"
