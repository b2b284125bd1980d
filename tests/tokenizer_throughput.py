"""Checker-side throughput comparison, kept under tests/ because it runs the oracle tokenizers (host only): texts/s of the C++ tokenizers vs the checkers (HF `tokenizers` for WordPiece, the Python
restatement of open_clip's SimpleTokenizer for CLIP BPE) on synthetic English-like text.  Not a bench line."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from marqo_b200.tokenizers import ClipBpeTokenizer, WordPieceTokenizer  # noqa: E402
from oracle import tokenizers as OT  # noqa: E402

rng = np.random.default_rng(0)
syll = ["ta", "ko", "mi", "re", "sol", "an", "ber", "ing", "ed", "pho", "to", "cat", "dog", "the", "of", "qu", "ick"]
words = sorted({"".join(syll[int(i)] for i in rng.integers(0, len(syll), size=int(rng.integers(1, 4)))) for _ in range(4000)})
vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + words[:3000] + ["##" + s for s in syll] + list(".,!?'")
texts = [" ".join(words[int(i)] for i in rng.integers(0, len(words), size=int(rng.integers(5, 120)))) + "."
         for _ in range(20000)]


def rate(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return len(texts) / best


wp = WordPieceTokenizer(("\n".join(vocab) + "\n").encode())
ref = OT.bert_wordpiece(vocab)
out = {"texts": len(texts), "mean_words": float(np.mean([len(t.split()) for t in texts]))}
out["wordpiece_cxx_texts_per_s"] = rate(lambda: wp(texts, max_length=128))
out["wordpiece_hf_tokenizers_texts_per_s"] = rate(lambda: OT.bert_encode_batch(ref, texts, 128))
merges = OT.train_toy_merges(texts[:2000], 2000)
bpe = ClipBpeTokenizer(merges.encode())
oracle = OT.SimpleTokenizerOracle(merges)
out["clip_bpe_cxx_texts_per_s"] = rate(lambda: bpe(texts))
sub = texts[:2000]
t0 = time.perf_counter()
oracle(sub)
out["clip_bpe_python_restatement_texts_per_s"] = len(sub) / (time.perf_counter() - t0)
print(json.dumps(out))
