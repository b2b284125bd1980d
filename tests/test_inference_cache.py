"""vectorise(enable_cache=True) (s2_inference.py:48-119) against outputs of THE REFERENCE's own cached path
(tests/golden/make_cache_golden.py -> cache_golden.json).  Host logic: runs without a GPU.

LRU is deterministic in every cachetools version, so both the vectors and WHICH texts reach the model are compared.  The
fixture was produced with a newer cachetools than the reference pins; its LFU breaks eviction ties by set iteration order
(arbitrary), the pinned 5.3.1 by insertion order — so for LFU the vectors are compared with the fixture, and the eviction
order with the pinned version's rule restated below."""
import collections
import json
from pathlib import Path

import numpy as np
import pytest

GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "cache_golden.json").read_text())


class _FakeModel:
    """Same deterministic stand-in as in make_cache_golden.py."""

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


def _run(case, monkeypatch):
    from marqo_b200 import s2_inference as s2
    from marqo_b200.inference_cache import MarqoInferenceCache
    monkeypatch.setattr(s2, "_marqo_inference_cache", MarqoInferenceCache(case["size"], case["cache_type"]))
    key, model = "cache-test-key", _FakeModel()
    s2._available_models[key] = {"model": model, "most_recently_used_time": 0, "model_size": 1}
    steps = []
    try:
        for step in case["steps"]:
            before = len(model.calls)
            if s2._marqo_inference_cache.is_enabled():
                res = s2._vectorise_with_cache(key, step["content"], True, s2.Modality.TEXT)
            else:
                res = s2._encode_without_cache(key, step["content"], True, s2.Modality.TEXT)
            steps.append((res, model.calls[before:]))
    finally:
        s2._available_models.pop(key, None)
    return steps


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"{c['cache_type']}-{c['size']}")
def test_cached_vectorise_matches_reference(case, monkeypatch):
    got = _run(case, monkeypatch)
    for (res, encoded), want in zip(got, case["steps"]):
        assert res == want["result"]                                   # same vectors, same positions, same types
        assert isinstance(res, list) and isinstance(res[0], list) and isinstance(res[0][0], float)
        if case["cache_type"] == "LRU" or case["size"] <= 1:
            assert encoded == want["encoded"]                          # exactly the reference's misses, in its batches


def test_lfu_follows_the_pinned_cachetools_rule(monkeypatch):
    """cachetools 5.3.1 LFUCache: every get / set is a use; evict the least used, ties -> first inserted."""
    from marqo_b200.inference_cache import MarqoInferenceCache
    c = MarqoInferenceCache(3, "LFU")
    uses = collections.OrderedDict()

    def ref_set(k):
        if k not in uses and len(uses) >= 3:
            victim = min(uses, key=uses.get)
            del uses[victim]
        uses[k] = uses.get(k, 0) + 1

    def ref_get(k):
        if k in uses:
            uses[k] += 1
            return True
        return False

    rng = np.random.default_rng(0)
    for _ in range(400):
        k = "t" + str(int(rng.integers(0, 7)))
        if rng.random() < 0.5:
            assert (c.get("m", k) is not None) == ref_get(k)
        else:
            c.set("m", k, [1.0])
            ref_set(k)
        assert all(("m", x) in c for x in uses) and c.currsize == len(uses)


def test_cache_configuration_errors():
    from marqo_b200.inference_cache import EnvVarError, MarqoInferenceCache
    assert not MarqoInferenceCache(0).is_enabled()
    with pytest.raises(EnvVarError):
        MarqoInferenceCache(-1)
    with pytest.raises(EnvVarError):
        MarqoInferenceCache(3, "FIFO")
    c = MarqoInferenceCache(2, "lru")
    with pytest.raises(TypeError):
        c.get("m", 5)
    with pytest.raises(ValueError):
        ("m",) in c
