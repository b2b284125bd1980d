"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes/numpy front-end of oracle/score_oracle.c (the CPU restatement of the
Vespa exact nearest-neighbour + closeness + top-k step; reference anchors are cited in the C file's header).
Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs — never by marqo_b200/.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional, Tuple

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

METRICS = {"prenormalized-angular": 0, "angular": 1, "dotproduct": 2, "euclidean": 3}


def _lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = _HERE / "libscore_oracle.so"
        if not path.exists():
            raise FileNotFoundError(f"{path} missing: run `python -m marqo_b200.build` (build_oracle)")
        lib = C.CDLL(str(path))
        lib.oracle_search.restype = C.c_int
        lib.oracle_closeness.restype = C.c_double
        lib.oracle_closeness.argtypes = [C.c_double, C.c_int]
        _LIB = lib
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def to_half(x: np.ndarray, normalize: bool = False) -> np.ndarray:
    """fp32 rows -> fp16 bit patterns (uint16), optionally L2-normalised in the CUDA path's arithmetic order."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape, dtype=np.uint16)
    rows = x.shape[0] if x.ndim == 2 else 1
    _lib().oracle_convert_rows(_p(x), _p(out), C.c_int64(rows), C.c_int(x.shape[-1]), C.c_int(1 if normalize else 0))
    return out


def search(queries: np.ndarray, corpus: np.ndarray, k: int, metric: str = "prenormalized-angular",
           doc_of_row: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Exact top-k.  queries/corpus are fp32; they are rounded to fp16 exactly as the row store does.
    -> (doc [nq,k] int32, row [nq,k] int32, closeness [nq,k] float64)."""
    m = METRICS[metric]
    qh = to_half(np.atleast_2d(queries), normalize=(m == 1))
    ch = to_half(corpus, normalize=(m == 1)) if corpus.shape[0] else np.empty((0, qh.shape[1]), np.uint16)
    return search_half(qh, ch, k, metric, doc_of_row)


def search_half(qh: np.ndarray, ch: np.ndarray, k: int, metric: str = "prenormalized-angular",
                doc_of_row: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    qh = np.ascontiguousarray(qh, dtype=np.uint16)
    ch = np.ascontiguousarray(ch, dtype=np.uint16)
    nq, dim = qh.shape
    n = ch.shape[0]
    d = None if doc_of_row is None else np.ascontiguousarray(doc_of_row, dtype=np.int32)
    od = np.empty((nq, k), np.int32)
    orow = np.empty((nq, k), np.int32)
    osc = np.empty((nq, k), np.float64)
    st = _lib().oracle_search(_p(qh), C.c_int(nq), _p(ch), C.c_int64(n), C.c_int(dim), _p(d), C.c_int(METRICS[metric]),
                              C.c_int(k), _p(od), _p(orow), _p(osc))
    if st != 0:
        raise MemoryError("oracle_search failed to allocate")
    return od, orow, osc


def modifiers(attrs: np.ndarray, mult: list, add: list) -> np.ndarray:
    """attrs: float64 [n_cols, n_docs] (NaN = missing cell); mult/add: [(column, weight), ...] -> [n_docs, 2] (mult, add)."""
    attrs = np.ascontiguousarray(attrs, dtype=np.float64)
    n_cols, n_docs = attrs.shape
    mc = np.asarray([c for c, _ in mult], dtype=np.int32)
    mw = np.asarray([w for _, w in mult], dtype=np.float64)
    ac = np.asarray([c for c, _ in add], dtype=np.int32)
    aw = np.asarray([w for _, w in add], dtype=np.float64)
    out = np.empty((n_docs, 2), np.float64)
    _lib().oracle_modifiers(_p(attrs), C.c_int(n_cols), C.c_int64(n_docs), _p(mc), _p(mw), C.c_int(len(mult)), _p(ac),
                            _p(aw), C.c_int(len(add)), _p(out))
    return out


def search_modified(queries: np.ndarray, corpus: np.ndarray, k: int, mod: np.ndarray,
                    metric: str = "prenormalized-angular", doc_of_row: Optional[np.ndarray] = None):
    """Exact top-k under modify(closeness) (see oracle_search_modified).  mod: [n_docs, 2] from modifiers()."""
    m = METRICS[metric]
    qh = to_half(np.atleast_2d(queries), normalize=(m == 1))
    ch = to_half(corpus, normalize=(m == 1))
    nq, dim = qh.shape
    d = None if doc_of_row is None else np.ascontiguousarray(doc_of_row, dtype=np.int32)
    ndoc = (int(d.max()) + 1) if d is not None else ch.shape[0]
    mod = np.ascontiguousarray(mod, dtype=np.float64)
    assert mod.shape[0] >= ndoc and mod.shape[1] == 2
    od = np.empty((nq, k), np.int32)
    orow = np.empty((nq, k), np.int32)
    osc = np.empty((nq, k), np.float64)
    lib = _lib()
    lib.oracle_search_modified.restype = C.c_int
    st = lib.oracle_search_modified(_p(qh), C.c_int(nq), _p(ch), C.c_int64(ch.shape[0]), C.c_int(dim), _p(d), C.c_int(m),
                                    C.c_int(k), _p(mod), _p(od), _p(orow), _p(osc))
    if st != 0:
        raise MemoryError("oracle_search_modified failed to allocate")
    return od, orow, osc


def closeness(dot: float, metric: str = "prenormalized-angular") -> float:
    return float(_lib().oracle_closeness(C.c_double(dot), C.c_int(METRICS[metric])))


def half_to_float(h: np.ndarray) -> np.ndarray:
    h = np.ascontiguousarray(h, dtype=np.uint16)
    out = np.empty(h.shape, np.float32)
    _lib().oracle_half_to_float(_p(h), _p(out), C.c_int64(h.size))
    return out
