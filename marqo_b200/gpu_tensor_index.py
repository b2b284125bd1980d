"""`VespaClient`-shaped adapter over the GPU row store (boundary B2).

Drop-in for `Config.vespa_client` (src/marqo/config.py:35) on the dense path: `feed_batch` / `query` / `get_batch` /
`delete_batch` keep the argument meaning and the response shapes of src/marqo/vespa/vespa_client.py:198-242,
:267-296, :405-440, :468-500 and src/marqo/vespa/models/{query_result,feed_response,get_document_response,
delete_document_response}.py.  Tensor queries (ranking == 'embedding_similarity', YQL made only of
`nearestNeighbor(...)` terms) are answered from the GPU-resident fp16 matrix by the exact score + top-k kernels;
everything else (bm25, hybrid, filters, score modifiers) is handed to the optional `delegate` — a real VespaClient —
or rejected with VespaError ("delegate ... rather than answer").

Semantics implemented (from the schema generators the reference ships, executed inside Vespa today):
  score(doc) = max over searched tensor fields, max over chunks, of closeness(q, chunk)
               (unstructured_vespa_schema.py:225-230,292-294; structured_vespa_index.py:645-688)
  matchfeatures: closest(<embeddings field>) = arg-max chunk label, distance(field,<embeddings field>)
               (consumed by _extract_highlights, structured_vespa_index.py:942-1000)
"""
from __future__ import annotations

import math
import re
import threading
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .engine import RowStore
from .errors import VespaError, VespaStatusError

RANK_PROFILE_EMBEDDING_SIMILARITY = "embedding_similarity"   # */common.py
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
                    num = s.doc_num.get(doc_id)
                    if num is None:
                        num = len(s.doc_ids)
                        s.doc_num[doc_id] = num
                        s.doc_ids.append(doc_id)
                        s.fields.append(None)
                        s.doc_rows.append({})
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
                        if mat.shape[1] != store.dim:
                            raise ValueError(f"field {f}: embedding dimension {mat.shape[1]} != index dimension {store.dim}")
                        row0 = len(store)
                        store.add(mat, np.full(len(keys), num, dtype=np.int32))
                        s.row_chunk[f].extend((num, k) for k in keys)
                        s.doc_rows[num][f] = list(range(row0, row0 + len(keys)))
                    s.fields[num] = {k: v for k, v in fields.items() if not k.startswith(EMBEDDINGS_PREFIX)}
                    responses.append({"status": 200, "pathId": path_id, "id": full_id, "message": None})
                except (ValueError, KeyError, TypeError) as e:
                    errors = True
                    responses.append({"status": 400, "pathId": path_id, "id": full_id, "message": str(e)})
        return _wrap("FeedBatchResponse", {"responses": responses, "errors": errors})

    # ------------------------------------------------------------------------------------------------ query
    def _is_tensor_query(self, yql: str, ranking: Optional[str], query_features: Optional[dict]) -> bool:
        if ranking != RANK_PROFILE_EMBEDDING_SIMILARITY:
            return False
        m = _WHERE.search(yql or "")
        if not m:
            return False
        rest = _NN_TERM.sub("", m.group(1))
        rest = re.sub(r"\bOR\b|[()\s;]", "", rest)
        if rest:           # an `AND <filter>` suffix (unstructured_vespa_index.py:62-66) or anything else
            return False
        qf = query_features or {}
        return any(k in qf for k in QUERY_INPUT_EMBEDDINGS) and not any(k.startswith("marqo__mult_weights")
                                                                          or k.startswith("marqo__add_weights")
                                                                          for k in qf)

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
        with self._lock:
            s = self._schemas.get(schema)
            children, n_docs = [], 0
            if s is not None:
                n_docs = sum(1 for d in s.doc_ids if d is not None)
                children = self._search(s, schema, fields, q, hits, offset)
        js = {"root": {"id": "toplevel", "relevance": 1.0, "fields": {"totalCount": len(children) + offset},
                       "coverage": {"coverage": 100, "documents": n_docs, "full": True, "nodes": 1, "results": 1,
                                    "resultsFull": 1},
                       "children": children}}
        return _wrap_query_result(js)

    def _search(self, s: _Schema, schema: str, fields: List[str], q: np.ndarray, hits: int, offset: int) -> List[dict]:
        k = hits + offset
        if k <= 0:
            return []
        best: Dict[int, Tuple[float, str, int]] = {}   # doc number -> (score, field, row)
        for f in fields:
            store = s.stores.get(f)
            if store is None or len(store) == 0:
                continue
            if q.shape[-1] != store.dim:
                raise VespaStatusError(400, f"Expected a tensor of dimension {store.dim} for query input but got "
                                            f"{q.shape[-1]}")
            doc, row, score = store.search(q[None, :], k)
            for d, r, sc in zip(doc[0], row[0], score[0]):
                if d < 0:
                    continue
                cur = best.get(int(d))
                if cur is None or sc > cur[0]:
                    best[int(d)] = (float(sc), f, int(r))
        ranked = sorted(best.items(), key=lambda kv: (-kv[1][0], kv[0]))[offset:offset + hits]
        children = []
        for num, (sc, f, r) in ranked:
            chunk_key = s.row_chunk[f][r][1]
            out_fields = dict(s.fields[num] or {})
            out_fields[MATCH_FEATURES] = {
                f"closest({f})": {"type": "tensor<float>(p{})", "cells": {chunk_key: 1.0}},
                f"distance(field,{f})": self._distance_from_closeness(sc),
            }
            children.append({"id": f"id:{schema}:{schema}::{s.doc_ids[num]}", "relevance": sc, "source": "content_default",
                             "fields": out_fields})
        return children

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
