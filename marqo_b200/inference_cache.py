"""Host mirror of the reference's inference cache (src/marqo/inference/inference_cache/marqo_inference_cache.py:10-100,
marqo_lru_cache.py, marqo_lfu_cache.py): a thread-safe embedding cache keyed by "<model_cache_key>||<text>", LRU or LFU,
sized by MARQO_INFERENCE_CACHE_SIZE (0 = disabled) and typed by MARQO_INFERENCE_CACHE_TYPE.

The reference builds on cachetools==5.3.1 (requirements.txt:6); the two policies are restated here with that version's
semantics: every get / set of a key counts as a use; LRU evicts the least recently used key; LFU evicts the least
frequently used key, ties going to the key that entered the cache first (Counter.most_common(1) over insertion order).
"""
from __future__ import annotations

import threading
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

CACHE_TYPES = ("LRU", "LFU")


class EnvVarError(ValueError):
    """marqo.api.exceptions.EnvVarError"""


class _LRU:
    def __init__(self, maxsize: int):
        self.maxsize, self._d = maxsize, OrderedDict()

    def get(self, key, default=None):
        if key not in self._d:
            return default
        self._d.move_to_end(key)
        return self._d[key]

    def set(self, key, value) -> None:
        if key in self._d:
            self._d.move_to_end(key)
        elif len(self._d) >= self.maxsize:
            self._d.popitem(last=False)
        self._d[key] = value

    def __contains__(self, key) -> bool:
        return key in self._d

    def __len__(self) -> int:
        return len(self._d)

    def clear(self) -> None:
        self._d.clear()


class _LFU:
    def __init__(self, maxsize: int):
        self.maxsize = maxsize
        self._d: Dict[str, List[float]] = {}
        self._uses: Dict[str, int] = {}          # insertion-ordered: ties go to the oldest entry

    def get(self, key, default=None):
        if key not in self._d:
            return default
        self._uses[key] += 1
        return self._d[key]

    def set(self, key, value) -> None:
        if key not in self._d and len(self._d) >= self.maxsize:
            victim = min(self._uses, key=self._uses.get)     # first minimum in insertion order
            del self._d[victim]
            del self._uses[victim]
        self._d[key] = value
        self._uses[key] = self._uses.get(key, 0) + 1

    def __contains__(self, key) -> bool:
        return key in self._d

    def __len__(self) -> int:
        return len(self._d)

    def clear(self) -> None:
        self._d.clear()
        self._uses.clear()


class MarqoInferenceCache:
    def __init__(self, cache_size: int = 0, cache_type: Optional[str] = "LRU"):
        if not isinstance(cache_size, int) or isinstance(cache_size, bool) or cache_size < 0:
            raise EnvVarError(f"Invalid cache size: {cache_size}. Must be a non-negative integer. Please set the "
                              f"'MARQO_INFERENCE_CACHE_SIZE' environment variable to a non-negative integer.")
        self._lock = threading.RLock()
        self._cache = None
        if cache_size > 0:
            kind = str(getattr(cache_type, "value", cache_type)).upper()
            if kind not in CACHE_TYPES:
                raise EnvVarError(f"Invalid cache type: {cache_type}. Must be one of {CACHE_TYPES}. Please set the "
                                  f"'MARQO_INFERENCE_CACHE_TYPE' environment variable to one of the valid cache types.")
            self._cache = (_LRU if kind == "LRU" else _LFU)(cache_size)

    @staticmethod
    def _generate_key(model_cache_key: str, content: str) -> str:
        if not isinstance(model_cache_key, str):
            raise TypeError(f"model_cache_key must be a string, not {type(model_cache_key)}")
        if not isinstance(content, str):
            raise TypeError(f"content must be a string, not {type(content)}")
        return f"{model_cache_key}||{content}"

    def is_enabled(self) -> bool:
        return self._cache is not None

    def get(self, model_cache_key: str, content: str, default=None) -> Optional[List[float]]:
        key = self._generate_key(model_cache_key, content)
        with self._lock:
            return self._cache.get(key, default)

    def set(self, model_cache_key: str, content: str, value: List[float]) -> None:
        key = self._generate_key(model_cache_key, content)
        with self._lock:
            self._cache.set(key, value)

    def __contains__(self, item: Tuple[str, str]) -> bool:
        if len(item) != 2:
            raise ValueError("MarqoInferenceCache received an unsupported input for 'in' operation. Expected input is a "
                             "tuple with 'model-cache-key' and 'content'.")
        with self._lock:
            return self._cache is not None and self._generate_key(*item) in self._cache

    def clear(self) -> None:
        with self._lock:
            if self._cache is not None:
                self._cache.clear()

    @property
    def maxsize(self) -> int:
        return self._cache.maxsize

    @property
    def currsize(self) -> int:
        return len(self._cache)
