"""Generates tests/golden/cache_golden.json by running THE REFERENCE's own cached vectorise path in the build container:
marqo.s2_inference.s2_inference.vectorise(..., enable_cache=True) (s2_inference.py:48-119) over
marqo.inference.inference_cache.MarqoInferenceCache (LRU / LFU on cachetools; the reference pins cachetools 5.3.1, this
container has a newer one — recorded in the fixture), with a deterministic fake model injected into `_available_models`
exactly as the reference's unit tests do (tests/s2_inference/test_vectorise.py:15-49).

    python tests/golden/make_cache_golden.py
"""
import json
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import _reference_import  # noqa: E402

_reference_import._Stub.__enter__ = lambda self: self          # readerwriterlock is stubbed: its locks become no-ops
_reference_import._Stub.__exit__ = lambda self, *a: False
_reference_import.install()
import torchaudio  # noqa: E402

if not hasattr(torchaudio, "set_audio_backend"):
    torchaudio.set_audio_backend = lambda *a, **k: None

import cachetools  # noqa: E402
import numpy as np  # noqa: E402

import marqo.s2_inference.s2_inference as s2  # noqa: E402
from marqo.inference.inference_cache.marqo_inference_cache import MarqoInferenceCache  # noqa: E402


class FakeModel:
    """Row i encodes the text deterministically (so a cached vector is recognisable) and records every batch."""

    def __init__(self):
        self.calls = []

    def encode(self, content, normalize=True, **kwargs):
        items = [content] if isinstance(content, str) else list(content)
        self.calls.append(list(items))
        out = np.zeros((len(items), 4), np.float32)
        for i, t in enumerate(items):
            h = sum(ord(c) * (k + 1) for k, c in enumerate(str(t))) % 9973
            out[i] = [h, len(str(t)), h % 7, 1.0 if normalize else 0.0]
        return out


SEQUENCE = [
    ["a", "b", "c"],          # fill
    ["a", "d"],               # hit a; d evicts the LRU / LFU victim
    "b",                      # str input (may have been evicted)
    ["c", "a", "e", "a"],     # duplicates inside one call
    ["e", "f", "g", "h"],     # more than the cache holds
    "a",
    ["h", "a", "zz"],
]
out = {"cachetools": cachetools.__version__, "cases": []}
props = {"name": "fake", "dimensions": 4, "type": "test", "tokens": 8}
for cache_type in ("LRU", "LFU"):
    for size in (3, 1, 0):
        s2._marqo_inference_cache = MarqoInferenceCache(cache_size=size, cache_type=cache_type)
        key = s2._create_model_cache_key("fake", "cpu", props)
        model = FakeModel()
        s2._available_models.clear()
        s2._available_models[key] = {"model": model, "most_recently_used_time": __import__("datetime").datetime.now(),
                                     "model_size": 1}
        steps = []
        for content in SEQUENCE:
            before = len(model.calls)
            res = s2.vectorise("fake", content, model_properties=props, device="cpu", normalize_embeddings=True,
                               enable_cache=True)
            steps.append({"content": content, "result": res, "encoded": model.calls[before:]})
        out["cases"].append({"cache_type": cache_type, "size": size, "steps": steps})
(HERE / "cache_golden.json").write_text(json.dumps(out))
print(out["cachetools"], len(out["cases"]), out["cases"][0]["steps"][1]["encoded"], out["cases"][3]["steps"][1]["encoded"])
