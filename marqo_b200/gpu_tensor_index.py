"""`VespaClient`-shaped adapter over the GPU row store (boundary B2).

Drop-in for `Config.vespa_client` (src/marqo/config.py:35) on the dense path: `feed_batch` / `query` / `get_batch` /
`delete_batch` keep the argument meaning and the response shapes of src/marqo/vespa/vespa_client.py:198-242,
:267-296, :405-440, :468-500 and src/marqo/vespa/models/{query_result,feed_response,get_document_response,
delete_document_response}.py.  Tensor queries (ranking == 'embedding_similarity', YQL made only of
`nearestNeighbor(...)` terms, optionally followed by the ` AND <filter>` text either index type generates) are
answered from the GPU-resident fp16 matrix by the exact score + top-k kernels; everything else (bm25, hybrid) is handed to the optional `delegate` — a real VespaClient — or rejected with VespaError (SURVEY §8b:
"delegate ... rather than answer").

Semantics implemented (from the schema generators the reference ships, executed inside Vespa today):
  score(doc) = max over searched tensor fields, max over chunks, of closeness(q, chunk)
               (unstructured_vespa_schema.py:225-230,292-294; structured_vespa_index.py:645-688)
  matchfeatures: closest(<embeddings field>) = arg-max chunk label, distance(field,<embeddings field>)
               (consumed by _extract_highlights, structured_vespa_index.py:942-1000)
  score modifiers: relevance = modify(score, query(marqo__mult_weights_tensor), query(marqo__add_weights_tensor))
               over the document's `marqo__score_modifiers` cells (unstructured_vespa_schema.py:225-230,266-271;
               vespa_index.py:106-150; unstructured_document.py:110-125), evaluated inside the scan kernel (f3)
"""
from __future__ import annotations

import math
import re
import threading
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from ._native import ERR_UNSUPPORTED, NativeError
from .engine import RowStore
from .yql_filter import FilterSyntaxError, compile_filter
from .errors import VespaError, VespaStatusError

RANK_PROFILE_EMBEDDING_SIMILARITY = "embedding_similarity"   # */common.py
RANK_PROFILE_EMBEDDING_SIMILARITY_MODIFIERS_2_9 = "embedding_similarity_modifiers"   # */common.py (index version < 2.10)
SCORE_MODIFIERS_FIELD = "marqo__score_modifiers"             # */common.py SCORE_MODIFIERS
# structured indexes split the cells over two tensors whose products / sums are multiplied / added together
# (structured_vespa_index/common.py:3-5, structured_vespa_schema.py:256-262): one sparse tensor over their union
SCORE_MODIFIER_FIELDS = (SCORE_MODIFIERS_FIELD, "marqo__score_modifiers_float", "marqo__score_modifiers_double_long")
MULT_WEIGHTS_INPUTS = ("marqo__mult_weights_tensor", "marqo__mult_weights")   # core/constants.py:22-27
ADD_WEIGHTS_INPUTS = ("marqo__add_weights_tensor", "marqo__add_weights")
MAX_ATTRIBUTE_COLUMNS = 64
MAX_MODIFIER_TERMS = 16
QUERY_INPUT_EMBEDDINGS = ("marqo__query_embedding", "embedding_query")
EMBEDDINGS_PREFIX = "marqo__embeddings"
CHUNKS_PREFIX = "marqo__chunks"
MATCH_FEATURES = "matchfeatures"

_NN_TERM = re.compile(r"\(\s*\{([^}]*)\}\s*nearestNeighbor\(\s*([A-Za-z0-9_]+)\s*,\s*([A-Za-z0-9_]+)\s*\)\s*\)")
_WHERE = re.compile(r"\bwhere\b(.*)$", re.IGNORECASE | re.DOTALL)


class _Obj:
    """Attribute bag with `.dict()` — stands in for the pydantic models where Marqo is not importable."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def dict(self, **_):
        def conv(v):
            if isinstance(v, _Obj):
                return v.dict()
            if isinstance(v, list):
                return [conv(x) for x in v]
            return v
        return {k: conv(v) for k, v in self.__dict__.items()}


def _wrap_query_result(js: dict):
    try:  # real Marqo environment: return its own pydantic model
        from marqo.vespa.models import QueryResult  # type: ignore
        return QueryResult(**js)
    except Exception:
        root = js["root"]
        children = [_Obj(id=c["id"], relevance=c["relevance"], source=c.get("source"), fields=c["fields"])
                    for c in root.get("children", [])]
        cov = root["coverage"]
        coverage = _Obj(coverage=cov["coverage"], documents=cov["documents"], full=cov["full"], nodes=cov["nodes"],
                        results=cov["results"], results_full=cov["resultsFull"], degraded=None)
        r = _Obj(id=root["id"], relevance=root["relevance"], fields=_Obj(total_count=root["fields"]["totalCount"]),
                 coverage=coverage, children=children)
        out = _Obj(root=r, timing=None, trace=None)
        out.hits = children
        out.total_count = root["fields"]["totalCount"]
        return out


def _wrap(kind: str, js: dict):
    try:
        import marqo.vespa.models as vm  # type: ignore
        return getattr(vm, kind)(**js)
    except Exception:
        resp = []
        for r in js["responses"]:
            d = dict(r)
            if "fields" in d:
                d["document"] = _Obj(id=d.get("id"), fields=d.pop("fields"))
            else:
                d.setdefault("document", None)
            d["path_id"] = d.pop("pathId", None)
            d.setdefault("message", None)
            resp.append(_Obj(**d))
        return _Obj(responses=resp, errors=js["errors"])


class _Schema:
    def __init__(self):
        self.stores: Dict[str, RowStore] = {}          # embeddings field -> row store
        self.row_chunk: Dict[str, List[Tuple[int, str]]] = {}  # embeddings field -> row -> (doc number, chunk key)
        self.doc_num: Dict[str, int] = {}              # external id -> document number
        self.doc_ids: List[Optional[str]] = []         # document number -> external id (None = deleted)
        self.fields: List[Optional[dict]] = []         # document number -> stored non-vector fields
        self.doc_rows: List[Dict[str, List[int]]] = [] # document number -> field -> rows
        self.attr_col: Dict[str, int] = {}             # score-modifier attribute name -> device column
        self.attrs: List[Dict[str, float]] = []        # document number -> its marqo__score_modifiers cells


class GpuTensorIndex:
    def __init__(self, metric: str = "prenormalized-angular", device: int = 0, delegate=None,
                 default_search_timeout_ms: int = 1000):
        self.metric = metric
        self.device = device
        self.delegate = delegate
        self.default_search_timeout_ms = default_search_timeout_ms
        self._schemas: Dict[str, _Schema] = {}
        self._lock = threading.RLock()

    def close(self) -> None:
        with self._lock:
            for s in self._schemas.values():
                for st in s.stores.values():
                    st.close()
            self._schemas.clear()

    # ------------------------------------------------------------------------------------------------ feed
    @staticmethod
    def _doc_id_and_fields(doc) -> Tuple[str, dict]:
        if isinstance(doc, dict):
            return doc["id"], doc["fields"]
        return doc.id, doc.fields

    def _tombstone(self, s: _Schema, num: int) -> None:
        for f, rows in s.doc_rows[num].items():
            if rows:
                s.stores[f].delete_doc(num)
        s.doc_rows[num] = {}
        if s.attrs[num]:
            for store in s.stores.values():
                store.set_attributes(-1, [num], None)
            s.attrs[num] = {}

    @staticmethod
    def _replay_attributes(s: _Schema, store: RowStore) -> None:
        """A row store created after documents were fed (a new tensor field) gets their attribute cells."""
        by_col: Dict[int, Tuple[List[int], List[float]]] = {}
        for num, attrs in enumerate(s.attrs):
            for name, v in attrs.items():
                ids, vals = by_col.setdefault(s.attr_col[name], ([], []))
                ids.append(num)
                vals.append(v)
        for col, (ids, vals) in by_col.items():
            store.set_attributes(col, ids, vals)

    def feed_batch(self, batch: List[Any], schema: str, concurrency: Optional[int] = None, timeout: int = 60):
        """vespa_client.py:267-296.  Embeddings arrive as fields['marqo__embeddings[_<field>]'] = {"0": [...], ...}
        (semi_structured_document.py:139-141; unstructured_add_document_handler.py:162-163)."""
        responses = []
        errors = False
        with self._lock:
            s = self._schemas.setdefault(schema, _Schema())
            for doc in batch:
                doc_id, fields = self._doc_id_and_fields(doc)
                path_id = f"/document/v1/{schema}/{schema}/docid/{doc_id}"
                full_id = f"id:{schema}:{schema}::{doc_id}"
                try:
                    emb_fields = {k: v for k, v in fields.items() if k.startswith(EMBEDDINGS_PREFIX)}
                    staged = {}
                    for f, cells in emb_fields.items():
                        if not isinstance(cells, dict):
                            raise ValueError(f"field {f}: expected a mapped tensor {{chunk: [floats]}}")
                        keys = list(cells.keys())
                        mat = np.asarray([cells[k] for k in keys], dtype=np.float32)
                        if mat.size and mat.ndim != 2:
                            raise ValueError(f"field {f}: ragged embeddings")
                        staged[f] = (keys, mat)
                    attrs: Dict[str, float] = {}
                    for tensor_field in SCORE_MODIFIER_FIELDS:
                        cells = fields.get(tensor_field) or {}
                        if isinstance(cells, dict) and "cells" in cells and isinstance(cells["cells"], (dict, list)):
                            cells = cells["cells"]          # Vespa's verbose tensor JSON form
                        if isinstance(cells, list):
                            cells = {c["address"]["p"]: c["value"] for c in cells}
                        for name, v in cells.items():
                            # structured indexes keep float-typed modifier fields in a tensor<float>: fp32 cells
                            attrs[str(name)] = float(np.float32(v)) if tensor_field.endswith("_float") else float(v)
                    for name, v in attrs.items():
                        if not math.isfinite(v):
                            raise ValueError(f"score modifier field {name}: value {v} is not finite")
                        if name not in s.attr_col:
                            if len(s.attr_col) >= MAX_ATTRIBUTE_COLUMNS - 1:     # the last column is the filter mask
                                raise ValueError(f"more than {MAX_ATTRIBUTE_COLUMNS - 1} distinct score-modifier fields")
                            s.attr_col[name] = len(s.attr_col)
                    num = s.doc_num.get(doc_id)
                    if num is None:
                        num = len(s.doc_ids)
                        s.doc_num[doc_id] = num
                        s.doc_ids.append(doc_id)
                        s.fields.append(None)
                        s.doc_rows.append({})
                        s.attrs.append({})
                    else:
                        self._tombstone(s, num)          # add_documents replaces by _id
                        s.doc_ids[num] = doc_id
                    for f, (keys, mat) in staged.items():
                        if not len(keys):
                            continue
                        store = s.stores.get(f)
                        if store is None:
                            store = RowStore(mat.shape[1], metric=self.metric, device=self.device)
                            s.stores[f] = store
                            s.row_chunk[f] = []
                            self._replay_attributes(s, store)
                        if mat.shape[1] != store.dim:
                            raise ValueError(f"field {f}: embedding dimension {mat.shape[1]} != index dimension {store.dim}")
                        row0 = len(store)
                        store.add(mat, np.full(len(keys), num, dtype=np.int32))
                        s.row_chunk[f].extend((num, k) for k in keys)
                        s.doc_rows[num][f] = list(range(row0, row0 + len(keys)))
                    s.fields[num] = {k: v for k, v in fields.items() if not k.startswith(EMBEDDINGS_PREFIX)}
                    s.attrs[num] = attrs
                    for store in s.stores.values():
                        for name, v in attrs.items():
                            store.set_attributes(s.attr_col[name], [num], [v])
                    responses.append({"status": 200, "pathId": path_id, "id": full_id, "message": None})
                except (ValueError, KeyError, TypeError) as e:
                    errors = True
                    responses.append({"status": 400, "pathId": path_id, "id": full_id, "message": str(e)})
        return _wrap("FeedBatchResponse", {"responses": responses, "errors": errors})

    # ------------------------------------------------------------------------------------------------ query
    @staticmethod
    def _split_where(yql: str) -> Tuple[bool, Optional[str]]:
        """-> (the where clause is nearestNeighbor terms [AND <filter>], filter text or None).
        unstructured_vespa_index.py:59-66: `where {tensor_term}{' AND ' + filter}`; structured_vespa_index.py:645-688 ORs
        one nearestNeighbor term per searched field inside parentheses."""
        m = _WHERE.search(yql or "")
        if not m:
            return False, None
        rest = _NN_TERM.sub("", m.group(1))
        head, sep, tail = rest.partition(" AND ")
        if re.sub(r"\bOR\b|[()\s;]", "", head):
            return False, None
        if not sep:
            return True, None
        text = tail.strip().rstrip(";").strip()
        return (True, text) if text else (False, None)

    def _is_tensor_query(self, yql: str, ranking: Optional[str], query_features: Optional[dict]) -> bool:
        if ranking not in (RANK_PROFILE_EMBEDDING_SIMILARITY, RANK_PROFILE_EMBEDDING_SIMILARITY_MODIFIERS_2_9):
            return False
        ok, filter_text = self._split_where(yql)
        if not ok:
            return False
        if filter_text is not None:
            try:
                compile_filter(filter_text)
            except FilterSyntaxError:     # text neither of the reference's filter generators emits: not ours to answer
                return False
        qf = query_features or {}
        if not any(k in qf for k in QUERY_INPUT_EMBEDDINGS):
            return False
        for k, v in qf.items():   # lexical / global modifier tensors belong to other rank profiles
            if (k.startswith("marqo__mult_weights") or k.startswith("marqo__add_weights")) and \
                    k not in MULT_WEIGHTS_INPUTS + ADD_WEIGHTS_INPUTS and v:
                return False
        n_mult = sum(len(qf.get(k) or {}) for k in MULT_WEIGHTS_INPUTS)
        n_add = sum(len(qf.get(k) or {}) for k in ADD_WEIGHTS_INPUTS)
        return n_mult <= MAX_MODIFIER_TERMS and n_add <= MAX_MODIFIER_TERMS

    def query(self, yql: str, hits: int = 10, ranking: str = None, model_restrict: str = None,
              query_features: Dict[str, Any] = None, timeout: float = None, **kwargs):
        """vespa_client.py:198-242."""
        if not self._is_tensor_query(yql, ranking, query_features):
            if self.delegate is not None:
                return self.delegate.query(yql, hits=hits, ranking=ranking, model_restrict=model_restrict,
                                           query_features=query_features, timeout=timeout, **kwargs)
            raise VespaError("GpuTensorIndex only answers exact tensor queries (ranking=embedding_similarity, "
                             "nearestNeighbor terms only); no delegate VespaClient is configured")
        schema = model_restrict
        if schema is None:
            m = re.search(r"\bfrom\s+([A-Za-z0-9_]+)", yql)
            schema = m.group(1) if m else None
        offset = int(kwargs.get("offset", 0) or 0)
        terms = _NN_TERM.findall(yql)
        fields = [t[1] for t in terms]
        qname = next(k for k in QUERY_INPUT_EMBEDDINGS if k in query_features)
        q = np.asarray(query_features[qname], dtype=np.float32)
        mult: Dict[str, float] = {}
        add: Dict[str, float] = {}
        for k in MULT_WEIGHTS_INPUTS:
            mult.update(self._weights(query_features.get(k)))
        for k in ADD_WEIGHTS_INPUTS:
            add.update(self._weights(query_features.get(k)))
        filter_text = self._split_where(yql)[1]
        keep = compile_filter(filter_text) if filter_text is not None else None
        with self._lock:
            s = self._schemas.get(schema)
            children, n_docs = [], 0
            if s is not None:
                n_docs = sum(1 for d in s.doc_ids if d is not None)
                try:
                    children = self._search(s, schema, fields, q, hits, offset, mult, add, keep)
                except NativeError as e:
                    if e.code != ERR_UNSUPPORTED:
                        raise
                    if self.delegate is not None:
                        return self.delegate.query(yql, hits=hits, ranking=ranking, model_restrict=model_restrict,
                                                   query_features=query_features, timeout=timeout, **kwargs)
                    raise VespaError(f"GpuTensorIndex cannot answer this query: {e.message}") from e
        js = {"root": {"id": "toplevel", "relevance": 1.0, "fields": {"totalCount": len(children) + offset},
                       "coverage": {"coverage": 100, "documents": n_docs, "full": True, "nodes": 1, "results": 1,
                                    "resultsFull": 1},
                       "children": children}}
        return _wrap_query_result(js)

    @staticmethod
    def _weights(tensor) -> Dict[str, float]:
        """A query tensor<double>(p{}) as Marqo sends it ({field: weight}) or in Vespa's {"cells": ...} forms."""
        if not tensor:
            return {}
        if isinstance(tensor, dict) and "cells" in tensor:
            tensor = tensor["cells"]
        if isinstance(tensor, list):
            return {c["address"]["p"]: float(c["value"]) for c in tensor}
        return {str(k): float(v) for k, v in tensor.items()}

    @staticmethod
    def _modifier_of(attrs: Dict[str, float], mult: Dict[str, float], add: Dict[str, float]) -> Tuple[float, float]:
        """(multiplier, addend) of one document — the host copy of the device table, used to recover the raw
        closeness for the distance() match-feature."""
        m, cnt = 1.0, 0
        for name, w in mult.items():
            if name in attrs:
                m *= w * attrs[name]
                cnt += 1
        if cnt == 0:
            m = 1.0
        a = 0.0
        for name, w in add.items():
            if name in attrs:
                a += w * attrs[name]
        return m, a

    MAX_FETCH = 10000   # b200_index_search's own k limit (Marqo's limit + offset cap, tensor_search.py:1568-1588)

    def _search(self, s: _Schema, schema: str, fields: List[str], q: np.ndarray, hits: int, offset: int,
                mult: Optional[Dict[str, float]] = None, add: Optional[Dict[str, float]] = None,
                keep=None) -> List[dict]:
        k = hits + offset
        if k <= 0:
            return []
        # a weight on an attribute no document has multiplies / adds nothing anywhere: drop the term
        mult_cols = [(s.attr_col[n], w) for n, w in (mult or {}).items() if n in s.attr_col]
        add_cols = [(s.attr_col[n], w) for n, w in (add or {}).items() if n in s.attr_col]
        modified = bool(mult_cols) or bool(add_cols)
        verdict: Dict[int, bool] = {}

        def allowed(num: int) -> bool:
            if num not in verdict:
                verdict[num] = bool(keep(s.fields[num] or {}))
            return verdict[num]

        best: Dict[int, Tuple[float, str, int]] = {}   # doc number -> (score, field, row)
        for f in fields:
            store = s.stores.get(f)
            if store is None or len(store) == 0:
                continue
            if q.shape[-1] != store.dim:
                raise VespaStatusError(400, f"Expected a tensor of dimension {store.dim} for query input but got "
                                            f"{q.shape[-1]}")
            # A filter is evaluated on the host against the stored fields; the exact top-k of the ALLOWED documents is
            # the first k allowed entries of the unfiltered ranking, so fetch deeper until k of them have been seen.
            fetch = k
            while True:
                if modified:
                    doc, row, score = store.search_modified(q[None, :], fetch, mult_cols, add_cols)
                else:
                    doc, row, score = store.search(q[None, :], fetch)
                found = [(int(d), int(r), float(sc)) for d, r, sc in zip(doc[0], row[0], score[0]) if d >= 0]
                if keep is not None:
                    kept = [h for h in found if allowed(h[0])]
                    exhausted = len(found) < fetch
                    if len(kept) < k and not exhausted:
                        if fetch >= self.MAX_FETCH:
                            # a very selective filter: evaluate it over the whole schema once and let the scan skip
                            # the excluded documents (one pass instead of ever deeper fetches)
                            found = self._masked_search(s, store, q, k, allowed, mult_cols, add_cols)
                            break
                        fetch = min(self.MAX_FETCH, fetch * 4)
                        continue
                    found = kept[:k]
                break
            for d, r, sc in found:
                cur = best.get(d)
                if cur is None or sc > cur[0]:
                    best[d] = (sc, f, r)
        ranked = sorted(best.items(), key=lambda kv: (-kv[1][0], kv[0]))[offset:offset + hits]
        children = []
        for num, (sc, f, r) in ranked:
            chunk_key = s.row_chunk[f][r][1]
            out_fields = dict(s.fields[num] or {})
            raw = sc
            if modified:   # relevance is the modified score; distance() stays the raw one
                m, a = self._modifier_of(s.attrs[num], mult or {}, add or {})
                raw = (sc - a) / m if m != 0 else float("nan")
            out_fields[MATCH_FEATURES] = {
                f"closest({f})": {"type": "tensor<float>(p{})", "cells": {chunk_key: 1.0}},
                f"distance(field,{f})": self._distance_from_closeness(raw),
            }
            children.append({"id": f"id:{schema}:{schema}::{s.doc_ids[num]}", "relevance": sc, "source": "content_default",
                             "fields": out_fields})
        return children

    MASK_COLUMN = MAX_ATTRIBUTE_COLUMNS - 1     # reserved attribute column: the per-query exclusion mask
    MASK_PENALTY = -1.0e30                      # addend of an excluded document: it can only rank after every kept one

    def _masked_search(self, s: _Schema, store: RowStore, q: np.ndarray, k: int, allowed, mult_cols, add_cols):
        """Exact top-k of the kept documents in one scan: excluded documents get an additive score modifier of -1e30
        through the reserved attribute column (set for this query, removed afterwards), so they sort after every kept
        document and are dropped from the result."""
        if len(s.attr_col) >= MAX_ATTRIBUTE_COLUMNS or len(add_cols) >= MAX_MODIFIER_TERMS:
            raise NativeError(ERR_UNSUPPORTED, "no free attribute column / modifier term for the filter mask")
        excluded = [num for num, doc_id in enumerate(s.doc_ids) if doc_id is not None and not allowed(num)]
        try:
            if excluded:
                store.set_attributes(self.MASK_COLUMN, excluded, [1.0] * len(excluded))
            doc, row, score = store.search_modified(q[None, :], k, mult_cols,
                                                    list(add_cols) + [(self.MASK_COLUMN, self.MASK_PENALTY)])
        finally:
            if excluded:
                store.set_attributes(self.MASK_COLUMN, excluded, None)
        return [(int(d), int(r), float(sc)) for d, r, sc in zip(doc[0], row[0], score[0])
                if d >= 0 and sc > self.MASK_PENALTY / 2]

    def _distance_from_closeness(self, closeness: float) -> float:
        if self.metric == "dotproduct":
            return -closeness
        return 1.0 / closeness - 1.0 if closeness > 0 else math.inf

    # ------------------------------------------------------------------------------------------------ get / delete
    def get_batch(self, ids: List[str], schema: str, concurrency: Optional[int] = None, timeout: int = 60,
                  fields: Optional[List[str]] = None):
        """vespa_client.py:405-440: 404 entries are returned, not raised.  Embeddings are read back from the fp16
        row store (use_existing_tensors, add_documents_handler.py:160-165)."""
        responses = []
        with self._lock:
            s = self._schemas.get(schema)
            for doc_id in ids:
                path_id = f"/document/v1/{schema}/{schema}/docid/{doc_id}"
                num = s.doc_num.get(doc_id) if s else None
                if num is None or s.doc_ids[num] is None:
                    responses.append({"status": 404, "pathId": path_id, "id": f"id:{schema}:{schema}::{doc_id}",
                                      "message": "Document not found"})
                    continue
                out = dict(s.fields[num] or {})
                for f, rows in s.doc_rows[num].items():
                    out[f] = {s.row_chunk[f][r][1]: s.stores[f].get_row(r).tolist() for r in rows}
                if fields is not None:
                    out = {k: v for k, v in out.items() if k in fields}
                responses.append({"status": 200, "pathId": path_id, "id": f"id:{schema}:{schema}::{doc_id}",
                                  "fields": out})
        return _wrap("GetBatchResponse", {"responses": responses, "errors": False})

    def delete_batch(self, ids: List[str], schema: str, concurrency: Optional[int] = None, timeout: int = 60):
        """vespa_client.py:468-500 (deleting a missing id is a 200 in Vespa's document API)."""
        responses = []
        with self._lock:
            s = self._schemas.get(schema)
            for doc_id in ids:
                num = s.doc_num.get(doc_id) if s else None
                if num is not None and s.doc_ids[num] is not None:
                    self._tombstone(s, num)
                    s.doc_ids[num] = None
                    s.fields[num] = None
                    del s.doc_num[doc_id]
                responses.append({"status": 200, "pathId": f"/document/v1/{schema}/{schema}/docid/{doc_id}",
                                  "id": f"id:{schema}:{schema}::{doc_id}", "message": None})
        return _wrap("DeleteBatchResponse", {"responses": responses, "errors": False})

    # ------------------------------------------------------------------------------------------------ persistence
    def save(self, directory: str) -> None:
        """Corpus persistence: one binary row-store snapshot per (schema, tensor field) —
        b200_index_save: fp16 rows + row -> document map + score-modifier columns — plus a JSON manifest with what
        Vespa would keep per document (ids, stored fields, chunk keys).  Restart = GpuTensorIndex.load(directory)."""
        import json
        import os
        os.makedirs(directory, exist_ok=True)
        manifest = {"format": 1, "metric": self.metric, "schemas": {}}
        with self._lock:
            for name, s in self._schemas.items():
                stores = {}
                for i, (f, st) in enumerate(sorted(s.stores.items())):
                    fname = f"{len(manifest['schemas'])}_{i}.b200idx"
                    st.save(os.path.join(directory, fname))
                    stores[f] = {"file": fname, "row_chunk": s.row_chunk[f]}
                manifest["schemas"][name] = {"stores": stores, "doc_ids": s.doc_ids, "fields": s.fields,
                                             "doc_rows": s.doc_rows, "attr_col": s.attr_col, "attrs": s.attrs}
            tmp = os.path.join(directory, "manifest.json.tmp")
            with open(tmp, "w", encoding="utf-8") as fh:
                json.dump(manifest, fh)
            os.replace(tmp, os.path.join(directory, "manifest.json"))

    @classmethod
    def load(cls, directory: str, device: int = 0, delegate=None) -> "GpuTensorIndex":
        import json
        import os
        with open(os.path.join(directory, "manifest.json"), encoding="utf-8") as fh:
            manifest = json.load(fh)
        if manifest.get("format") != 1:
            raise VespaError(f"{directory}: unknown GpuTensorIndex snapshot format {manifest.get('format')!r}")
        ix = cls(metric=manifest["metric"], device=device, delegate=delegate)
        for name, js in manifest["schemas"].items():
            s = _Schema()
            s.doc_ids = list(js["doc_ids"])
            s.doc_num = {d: i for i, d in enumerate(s.doc_ids) if d is not None}
            s.fields = list(js["fields"])
            s.doc_rows = [{f: list(r) for f, r in d.items()} for d in js["doc_rows"]]
            s.attr_col = {k: int(v) for k, v in js["attr_col"].items()}
            s.attrs = [dict(a) for a in js["attrs"]]
            for f, st in js["stores"].items():
                s.stores[f] = RowStore.load(os.path.join(directory, st["file"]), device=device)
                s.row_chunk[f] = [(int(n), str(k)) for n, k in st["row_chunk"]]
            ix._schemas[name] = s
        return ix

    def get_document_count(self, schema: str) -> int:
        with self._lock:
            s = self._schemas.get(schema)
            return 0 if s is None else sum(1 for d in s.doc_ids if d is not None)


def gather_documents_from_response(response, tensor_fields_by_embeddings_field: Optional[Dict[str, str]] = None,
                                   highlights: bool = True) -> Dict[str, Any]:
    """Hit -> Marqo document (`_id`, stored fields, `_score`, `_highlights`): the arithmetic-free part of
    src/marqo/tensor_search/tensor_search.py:1771-1791 and structured_vespa_index.py:942-1000 for this adapter's
    QueryResult (a9).  `tensor_fields_by_embeddings_field` maps 'marqo__embeddings_<f>' -> marqo field name; the
    default strips the prefix."""
    hits = []
    for child in response.hits:
        fields = child.dict()["fields"] if hasattr(child, "dict") else child["fields"]
        doc = {k: v for k, v in fields.items() if not k.startswith("marqo__") and k != MATCH_FEATURES}
        doc["_id"] = child.id.split("::")[-1]
        doc["_score"] = child.relevance
        if highlights:
            mf = fields.get(MATCH_FEATURES, {})
            best = None
            for key, val in mf.items():
                if key.startswith("closest(") and val.get("cells"):
                    emb = key[len("closest("):-1]
                    dist = mf.get(f"distance(field,{emb})")
                    if best is None or dist < best[0]:
                        best = (dist, emb, next(iter(val["cells"])))
            doc["_highlights"] = []
            if best is not None:
                _, emb, chunk_key = best
                suffix = emb[len(EMBEDDINGS_PREFIX):]
                chunks = fields.get(CHUNKS_PREFIX + suffix)
                name = (tensor_fields_by_embeddings_field or {}).get(emb, suffix.lstrip("_") or emb)
                if chunks is not None:
                    doc["_highlights"] = [{name: chunks[int(chunk_key)]}]
        hits.append(doc)
    return {"hits": hits}
