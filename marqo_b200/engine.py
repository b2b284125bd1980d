"""Thin Python objects over the C ABI handles (include/marqo_b200.h).  No arithmetic happens here."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _native as N

_METRICS = {
    # names: src/marqo/core/models/marqo_index.py:63-69 (DistanceMetric)
    "prenormalized-angular": N.METRIC_PRENORMALIZED_ANGULAR,
    "angular": N.METRIC_ANGULAR,
    "dotproduct": N.METRIC_DOTPRODUCT,
    "euclidean": N.METRIC_EUCLIDEAN,
}


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _as(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


class RowStore:
    """GPU-resident fp16 embedding matrix with exact top-k search (b200_index_*)."""

    def __init__(self, dim: int, metric: str = "prenormalized-angular", device: int = 0, capacity: int = 0,
                 _handle=None):
        self._lib = N.load()
        self.dim = int(dim)
        self.metric = metric
        self.device = int(device)
        if _handle is not None:
            self._h = _handle
            return
        if metric not in _METRICS:
            raise ValueError(f"unknown distance metric {metric!r}; expected one of {sorted(_METRICS)}")
        h = C.c_void_p()
        N.check(self._lib.b200_index_create(self.device, self.dim, _METRICS[metric], int(capacity), C.byref(h)))
        self._h = h

    # -- lifetime -------------------------------------------------------------------------------
    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.b200_index_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _handle(self):
        if not self._h:
            raise RuntimeError("RowStore is closed")
        return self._h

    # -- mutation -------------------------------------------------------------------------------
    def add(self, vecs, doc_ids: Optional[Sequence[int]] = None) -> None:
        v = _as(vecs, np.float32)
        if v.ndim != 2 or v.shape[1] != self.dim:
            raise ValueError(f"expected [m, {self.dim}] embeddings, got {v.shape}")
        d = None
        if doc_ids is not None:
            d = _as(doc_ids, np.int32)
            if d.shape != (v.shape[0],):
                raise ValueError("doc_ids must have one entry per row")
        N.check(self._lib.b200_index_add(self._handle(), _ptr(v), _ptr(d), v.shape[0]))

    def add_device(self, d_vecs_ptr: int, m: int, d_doc_ids_ptr: Optional[int] = None) -> None:
        N.check(self._lib.b200_index_add_device(self._handle(), C.c_void_p(d_vecs_ptr),
                                                C.c_void_p(d_doc_ids_ptr) if d_doc_ids_ptr else None, int(m)))

    def delete_doc(self, doc_id: int) -> None:
        N.check(self._lib.b200_index_delete_doc(self._handle(), int(doc_id)))

    # -- queries --------------------------------------------------------------------------------
    def __len__(self) -> int:
        n = C.c_int64(0)
        N.check(self._lib.b200_index_num_rows(self._handle(), C.byref(n)))
        return n.value

    def get_row(self, row: int) -> np.ndarray:
        out = np.empty(self.dim, dtype=np.float32)
        N.check(self._lib.b200_index_get_row(self._handle(), int(row), _ptr(out)))
        return out

    def search(self, queries, k: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """-> (doc [nq,k] int32, row [nq,k] int32, closeness [nq,k] float64); unused slots are -1/-1/-inf."""
        q = _as(queries, np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"expected [nq, {self.dim}] queries, got {q.shape}")
        nq = q.shape[0]
        doc = np.empty((nq, k), dtype=np.int32)
        row = np.empty((nq, k), dtype=np.int32)
        score = np.empty((nq, k), dtype=np.float64)
        N.check(self._lib.b200_index_search(self._handle(), _ptr(q), nq, int(k), _ptr(doc), _ptr(row), _ptr(score)))
        return doc, row, score

    def search_device(self, d_q_ptr: int, nq: int, k: int, d_doc_ptr: int, d_row_ptr: int, d_score_ptr: int,
                      sync: bool = True) -> None:
        N.check(self._lib.b200_index_search_device(self._handle(), C.c_void_p(d_q_ptr), int(nq), int(k),
                                                   C.c_void_p(d_doc_ptr), C.c_void_p(d_row_ptr),
                                                   C.c_void_p(d_score_ptr), 1 if sync else 0))

    def last_timing(self) -> Tuple[float, float]:
        a, b = C.c_float(0), C.c_float(0)
        N.check(self._lib.b200_index_last_timing(self._handle(), C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- persistence ----------------------------------------------------------------------------
    def save(self, path: str) -> None:
        N.check(self._lib.b200_index_save(self._handle(), str(path).encode()))

    @classmethod
    def load(cls, path: str, device: int = 0) -> "RowStore":
        lib = N.load()
        h = C.c_void_p()
        N.check(lib.b200_index_load(int(device), str(path).encode(), C.byref(h)))
        d, m, dev = C.c_int(0), C.c_int(0), C.c_int(0)
        N.check(lib.b200_index_info(h, C.byref(d), C.byref(m), C.byref(dev)))
        names = {v: k for k, v in _METRICS.items()}
        return cls(dim=d.value, metric=names[m.value], device=dev.value, _handle=h)


def topk_merge(doc: np.ndarray, row: np.ndarray, score: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Merge per-shard lists [nshards, nq, k] into [nq, k] under (score desc, doc asc)."""
    doc, row, score = _as(doc, np.int32), _as(row, np.int32), _as(score, np.float64)
    ns, nq, k = doc.shape
    od = np.empty((nq, k), np.int32)
    orow = np.empty((nq, k), np.int32)
    osc = np.empty((nq, k), np.float64)
    N.check(N.load().b200_topk_merge(ns, nq, k, _ptr(doc), _ptr(row), _ptr(score), _ptr(od), _ptr(orow), _ptr(osc)))
    return od, orow, osc
