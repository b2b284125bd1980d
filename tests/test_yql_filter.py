"""Filtered tensor search, host side (no GPU): the YQL filter text comes from THE REFERENCE's own generator
(tests/golden/make_filter_golden.py -> filter_golden.json); marqo_b200.yql_filter must accept every one of them and select
the documents Vespa would, given the schema the reference deploys (whole-value, case-insensitive attribute matching;
sameElement on one map entry)."""
import json
from pathlib import Path

import pytest

GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "filter_golden.json").read_text())


def _vespa_fields(doc_id, logical):
    """Logical Marqo document -> the stored Vespa fields of an unstructured index (unstructured_document.py:95-125)."""
    f = {"marqo__id": doc_id, "marqo__short_string_fields": {}, "marqo__string_array": [], "marqo__int_fields": {},
         "marqo__float_fields": {}, "marqo__bool_fields": {}}
    for k, v in logical.items():
        if isinstance(v, bool):
            f["marqo__bool_fields"][k] = int(v)
        elif isinstance(v, str):
            f["marqo__short_string_fields"][k] = v
        elif isinstance(v, list):
            f["marqo__string_array"].extend(f"{k}::{e}" for e in v)
        elif isinstance(v, int):
            f["marqo__int_fields"][k] = v
        elif isinstance(v, float):
            f["marqo__float_fields"][k] = v
    return f


DOCS = {
    "d0": dict(color="red", price=15, in_stock=True, tags=["sale", "new"], year=2024, rating=4.5, title='say "hi"',
               a=1, b=2, c=3, d=4, **{"meta.size": "large"}),
    "d1": dict(color="blue", price=25, in_stock=False, tags=["new"], year=2023, rating=3.0, a=1, b=5, c=3, d=5),
    "d2": dict(color="dark red", price=10.5, in_stock=True, year=2024, rating=4.5, a=2),
    "d3": dict(color="Red", price=3, in_stock="true", tags=["sale"]),
    "doc7": dict(),
}
EXPECTED = {
    "color:red": {"d0", "d3"},
    "color:(dark red)": {"d2"},
    "price:[10 TO 20]": {"d0", "d2"},
    "price:[10.5 TO *]": {"d0", "d1", "d2"},
    "price:[* TO 3]": {"d3"},
    "in_stock:true": {"d0", "d2", "d3"},
    "in_stock:false AND color:red": set(),
    "color:red OR color:blue": {"d0", "d1", "d3"},
    "NOT color:red": {"d1", "d2", "doc7"},
    "(color:red OR color:blue) AND price:[0 TO 100]": {"d0", "d1", "d3"},
    "NOT (color:red AND in_stock:true)": {"d1", "d2", "doc7"},
    "tags:sale": {"d0", "d3"},
    "year:2024": {"d0", "d2"},
    "rating:4.5": {"d0", "d2"},
    "_id:doc7": {"doc7"},
    "meta.size:large": {"d0"},
    'title:(say \\"hi\\")': {"d0"},
    "a:1 AND (b:2 OR (c:3 AND NOT d:4))": {"d0", "d1"},
    "color:RED": {"d0", "d3"},
    "NOT (_id:(d0) OR _id:(d1) OR _id:(d2))": {"d3", "doc7"},
    "(color:red) AND NOT (_id:(d0) OR _id:(doc7))": {"d3"},
}


def test_every_reference_filter_is_accepted_and_selects_the_right_documents():
    from marqo_b200.yql_filter import compile_filter
    assert {g["filter"] for g in GOLD} == set(EXPECTED)
    stored = {doc_id: _vespa_fields(doc_id, logical) for doc_id, logical in DOCS.items()}
    for g in GOLD:
        pred = compile_filter(g["yql"])
        got = {doc_id for doc_id, fields in stored.items() if pred(fields)}
        assert got == EXPECTED[g["filter"]], (g["filter"], g["yql"], got)


@pytest.mark.parametrize("bad", [
    "price > 3",                                             # not an operator either generator emits
    '(marqo__id contains "x") AND',
    '((marqo__id contains "x") AND (marqo__id contains "y") OR (marqo__id contains "z"))',
    '(marqo__int_fields contains sameElement(key contains "p"))',
    '(marqo__id contains "unterminated)',
    'default contains "x"',                                  # a lexical term, not a filter attribute
    '(title contains "x")',
])
def test_other_grammars_are_rejected(bad):
    from marqo_b200.yql_filter import FilterSyntaxError, compile_filter
    with pytest.raises(FilterSyntaxError):
        compile_filter(bad)


def _nn(field="marqo__embeddings", k=10):
    return (f"({{targetHits:{k}, approximate:False, hnsw.exploreAdditionalHits:0}}"
            f"nearestNeighbor({field}, marqo__query_embedding))")


def test_adapter_recognises_filtered_tensor_queries():
    """unstructured_vespa_index.py:59-66 appends ' AND <filter>' to the nearestNeighbor term; the adapter answers those
    (and only those) whose filter text is in the grammar above."""
    from marqo_b200.gpu_tensor_index import GpuTensorIndex
    ix = GpuTensorIndex.__new__(GpuTensorIndex)          # classification is pure host logic: no device needed
    qf = {"marqo__query_embedding": [0.0] * 8}
    for g in GOLD:
        yql = f"select * from s1 where {_nn()} AND {g['yql']}"
        assert GpuTensorIndex._split_where(yql) == (True, g["yql"])
        assert ix._is_tensor_query(yql, "embedding_similarity", qf)
    multi = f"select * from s1 where ({_nn('marqo__embeddings_a')} OR {_nn('marqo__embeddings_b')}) AND {GOLD[0]['yql']}"
    assert ix._is_tensor_query(multi, "embedding_similarity", qf)
    assert ix._is_tensor_query(f"select * from s1 where {_nn()}", "embedding_similarity", qf)
    # text outside both generators' grammars, lexical terms, other rank profiles: not answered here
    assert not ix._is_tensor_query(f'select * from s1 where {_nn()} AND (title matches "x")', "embedding_similarity", qf)
    assert not ix._is_tensor_query(f'select * from s1 where {_nn()} AND default contains "x"', "embedding_similarity", qf)
    assert not ix._is_tensor_query(f"select * from s1 where {_nn()} AND {GOLD[0]['yql']}", "bm25", qf)
    assert not ix._is_tensor_query(f'select * from s1 where default contains "x" AND {GOLD[0]["yql"]}',
                                   "embedding_similarity", qf)


class _NumpyRowStore:
    """CPU stand-in with RowStore's interface and arithmetic contract (fp16 rows, closeness = 1 / (2 - q.e) in fp64,
    order (score desc, doc asc), best chunk per document, document bitset filter) — lets the adapter's host logic run
    without a device."""

    def __init__(self, dim, metric="prenormalized-angular", device=0, capacity=0):
        self.dim, self.rows, self.docs, self.attrs = dim, [], [], {}
        self.searches = 0

    def __len__(self):
        return len(self.rows)

    def add(self, vecs, doc_ids=None):
        import numpy as np
        for i, v in enumerate(np.asarray(vecs, np.float32)):
            self.rows.append(v.astype(np.float16).astype(np.float64))
            self.docs.append(int(doc_ids[i]) if doc_ids is not None else len(self.docs))

    def delete_doc(self, doc_id):
        self.docs = [-1 if d == doc_id else d for d in self.docs]

    def delete_rows(self, rows):
        for r in rows:
            self.docs[int(r)] = -1

    def compact(self):
        import numpy as np
        new_of_old = np.full(len(self.rows), -1, np.int32)
        live = [i for i, d in enumerate(self.docs) if d >= 0]
        new_of_old[live] = np.arange(len(live))
        self.rows = [self.rows[i] for i in live]
        self.docs = [self.docs[i] for i in live]
        return new_of_old

    def set_attributes(self, column, doc_ids, values):
        for i, d in enumerate(doc_ids):
            if column == -1:
                for col in self.attrs.values():
                    col.pop(int(d), None)
            elif values is None:
                self.attrs.setdefault(column, {}).pop(int(d), None)
            else:
                self.attrs.setdefault(column, {})[int(d)] = float(values[i])

    def set_attributes_multi(self, columns, doc_ids, values):
        for c, d, v in zip(columns, doc_ids, values):
            self.attrs.setdefault(int(c), {})[int(d)] = float(v)

    def _rank(self, q, k, mult=(), add=(), allowed=None):
        import numpy as np
        qh = np.asarray(q, np.float32).reshape(-1).astype(np.float16).astype(np.float64)
        best = {}
        for r, (v, d) in enumerate(zip(self.rows, self.docs)):
            if d < 0 or (allowed is not None and not allowed(d)):
                continue
            c = 1.0 / (2.0 - float(v @ qh))
            if d not in best or c > best[d][0]:
                best[d] = (c, r)
        out = []
        for d, (c, r) in best.items():
            cells = [w * self.attrs[col][d] for col, w in mult if d in self.attrs.get(col, {})]
            m = float(np.prod(cells)) if cells else 1.0
            a = sum(w * self.attrs[col][d] for col, w in add if d in self.attrs.get(col, {}))
            out.append((m * c + a, d, r))
        out.sort(key=lambda t: (-t[0], t[1]))
        doc = np.full(k, -1, np.int32)
        row = np.full(k, -1, np.int32)
        score = np.full(k, -np.inf)
        for i, (sc, d, r) in enumerate(out[:k]):
            doc[i], row[i], score[i] = d, r, sc
        return doc, row, score

    def search(self, q, k, mult=(), add=(), filter_bits=None, filter_docs=0, filter_tag=0):
        import numpy as np
        self.searches += 1
        Q = np.atleast_2d(np.asarray(q, np.float32))
        allowed = None
        if filter_bits is not None:
            bits = np.unpackbits(np.asarray(filter_bits, np.uint32).view(np.uint8), bitorder="little")
            allowed = lambda d: d < filter_docs and bool(bits[d])
        res = [self._rank(x, k, mult, add, allowed) for x in Q]
        return tuple(np.stack([r[i] for r in res]) for i in range(3))

    def search_modified(self, q, k, mult=(), add=()):
        return self.search(q, k, mult=mult, add=add)

    def get_row(self, r):
        return self.rows[r].astype("float32")

    def get_rows(self, rows):
        import numpy as np
        return np.stack([self.rows[int(r)].astype("float32") for r in rows])

    def close(self):
        pass


def test_filtered_search_scenario_on_the_cpu_stand_in(monkeypatch):
    import marqo_b200.gpu_tensor_index as gti
    from _filter_scenario import run_filtered_search_scenario
    monkeypatch.setattr(gti, "RowStore", _NumpyRowStore)
    run_filtered_search_scenario()


def test_structured_index_modifier_tensors_on_the_cpu_stand_in(monkeypatch):
    """Structured indexes feed two modifier tensors (float / double_long, structured_vespa_index.py:217-230) and rank
    with the product / sum over both (structured_vespa_schema.py:256-262): one sparse tensor over their union here."""
    import numpy as np
    import marqo_b200.gpu_tensor_index as gti
    from _filter_scenario import _doc, _yql
    monkeypatch.setattr(gti, "RowStore", _NumpyRowStore)
    rng = np.random.default_rng(3)
    vecs = rng.standard_normal((20, 64)).astype(np.float32)
    vecs /= np.linalg.norm(vecs, axis=1, keepdims=True)
    ix = gti.GpuTensorIndex()
    docs = [_doc(f"d{i}", {"marqo__id": f"d{i}", "marqo__score_modifiers_float": {"rating": 0.1 * (i % 5)},
                           "marqo__score_modifiers_double_long": {"sold": i * 3}}, {"body": (["c"], vecs[i:i + 1])})
            for i in range(20)]
    assert not ix.feed_batch(docs, "s1").errors
    q = vecs[7]
    res = ix.query(_yql("s1", ["body"], 20), hits=20, ranking="embedding_similarity", model_restrict="s1",
                   query_features={"marqo__query_embedding": q.tolist(), "marqo__mult_weights_tensor": {"rating": 2.0},
                                   "marqo__add_weights_tensor": {"sold": 0.001}})
    qh = q.astype(np.float16).astype(np.float64)
    want = {}
    for i in range(20):
        c = 1.0 / (2.0 - float(vecs[i].astype(np.float16).astype(np.float64) @ qh))
        want[f"d{i}"] = 2.0 * float(np.float32(0.1 * (i % 5))) * c + 0.001 * (i * 3)
    order = sorted(want, key=lambda d: (-want[d], int(d[1:])))
    assert [h.id.split("::")[-1] for h in res.hits] == order
    assert all(abs(h.relevance - want[h.id.split("::")[-1]]) < 1e-12 for h in res.hits)


STRUCTURED_GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "filter_golden_structured.json").read_text())


def _structured_fields(doc_id, logical):
    """Logical document -> stored fields of a structured index whose fields all carry the Filter feature
    (structured_vespa_index.py:183-215: the value goes under the field's filter_field_name; booleans as bytes)."""
    f = {"marqo__id": doc_id}
    for k, v in logical.items():
        f[f"marqo__filter_{k}"] = int(v) if isinstance(v, bool) else v
    return f


STRUCTURED_DOCS = {
    "d0": dict(color="red", price=15.0, in_stock=True, tags=["sale", "new"], year=2024, rating=4.5, title='say "hi"'),
    "d1": dict(color="blue", price=25.0, in_stock=False, tags=["new"], year=2023, rating=3.0),
    "d2": dict(color="dark red", price=10.5, in_stock=True, year=2024, rating=4.5),
    "d3": dict(color="Red", price=3.0, in_stock=True, tags=["sale"]),
    "doc7": dict(),
}
STRUCTURED_EXPECTED = {
    "color:red": {"d0", "d3"},
    "price:[10 TO 20]": {"d0", "d2"},
    "price:[10.5 TO *]": {"d0", "d1", "d2"},
    "price:[* TO 3]": {"d3"},
    "in_stock:true": {"d0", "d2", "d3"},
    "in_stock:false AND color:red": set(),
    "color:red OR color:blue": {"d0", "d1", "d3"},
    "NOT color:red": {"d1", "d2", "doc7"},
    "(color:red OR color:blue) AND price:[0 TO 100]": {"d0", "d1", "d3"},
    "NOT (color:red AND in_stock:true)": {"d1", "d2", "doc7"},
    "tags:sale": {"d0", "d3"},
    "year:2024": {"d0", "d2"},
    "rating:4.5": {"d0", "d2"},
    "_id:doc7": {"doc7"},
    'title:(say \\"hi\\")': {"d0"},
    "color in (red, blue)": {"d0", "d1", "d3"},
    "year in (2023, 2024)": {"d0", "d1", "d2"},
    "color:RED": {"d0", "d3"},
    "NOT _id in (d0, d1)": {"d2", "d3", "doc7"},
}


def test_structured_index_filters_from_the_reference_generator():
    from marqo_b200.yql_filter import compile_filter
    assert {g["filter"] for g in STRUCTURED_GOLD} == set(STRUCTURED_EXPECTED)
    stored = {doc_id: _structured_fields(doc_id, logical) for doc_id, logical in STRUCTURED_DOCS.items()}
    for g in STRUCTURED_GOLD:
        pred = compile_filter(g["yql"])
        got = {doc_id for doc_id, fields in stored.items() if pred(fields)}
        assert got == STRUCTURED_EXPECTED[g["filter"]], (g["filter"], g["yql"], got)


def test_filters_are_one_scan_through_a_document_bitset(monkeypatch):
    """A filtered query is ONE scan whatever its selectivity: the filter is compiled to a per-document bitset once per
    distinct filter string, kept current by feed / delete, and handed to the row store."""
    import numpy as np
    import marqo_b200.gpu_tensor_index as gti
    from _filter_scenario import _doc, _yql
    monkeypatch.setattr(gti, "RowStore", _NumpyRowStore)
    rng = np.random.default_rng(5)
    n = 200
    vecs = rng.standard_normal((n, 64)).astype(np.float32)
    vecs /= np.linalg.norm(vecs, axis=1, keepdims=True)
    ix = gti.GpuTensorIndex()
    docs = [_doc(f"d{i}", {"marqo__id": f"d{i}", "marqo__int_fields": {"bucket": i % 50},
                           "marqo__score_modifiers": {"pop": float(i % 3 + 1)}}, {"body": (["c"], vecs[i:i + 1])})
            for i in range(n)]
    assert not ix.feed_batch(docs, "s1").errors
    store = next(iter(ix._schemas["s1"].stores.values()))
    q = vecs[3]
    qh = q.astype(np.float16).astype(np.float64)
    c = {f"d{i}": 1.0 / (2.0 - float(vecs[i].astype(np.float16).astype(np.float64) @ qh)) for i in range(n)}
    flt = '((marqo__int_fields contains sameElement(key contains "bucket", value = 7)) OR ' \
          '(marqo__float_fields contains sameElement(key contains "bucket", value = 7)))'        # 4 of 200 documents
    before = store.searches
    res = ix.query(_yql("s1", ["body"], 10) + f" AND {flt}", hits=10, ranking="embedding_similarity", model_restrict="s1",
                   query_features={"marqo__query_embedding": q.tolist()})
    assert store.searches == before + 1                                   # one scan, no deeper-fetch loop
    want = sorted((d for d in c if int(d[1:]) % 50 == 7), key=lambda d: (-c[d], int(d[1:])))
    assert [h.id.split("::")[-1] for h in res.hits] == want and len(want) == 4
    assert all(abs(h.relevance - c[h.id.split("::")[-1]]) < 1e-12 for h in res.hits)
    entry = ix._schemas["s1"].filters[flt]
    tag = entry.tag
    # together with score modifiers; same filter string -> same cached bitset (tag unchanged)
    resm = ix.query(_yql("s1", ["body"], 3) + f" AND {flt}", hits=3, ranking="embedding_similarity", model_restrict="s1",
                    query_features={"marqo__query_embedding": q.tolist(), "marqo__add_weights_tensor": {"pop": 0.5}})
    wantm = sorted(want, key=lambda d: (-(c[d] + 0.5 * (int(d[1:]) % 3 + 1)), int(d[1:])))[:3]
    assert [h.id.split("::")[-1] for h in resm.hits] == wantm and entry.tag == tag
    plain = ix.query(_yql("s1", ["body"], 5), hits=5, ranking="embedding_similarity", model_restrict="s1",
                     query_features={"marqo__query_embedding": q.tolist()})
    assert [h.id.split("::")[-1] for h in plain.hits] == sorted(c, key=lambda d: (-c[d], int(d[1:])))[:5]
    # the bitset follows the corpus: a new matching document, an overwrite that stops matching, a delete
    extra = rng.standard_normal((1, 64)).astype(np.float32)
    ix.feed_batch([_doc("new", {"marqo__id": "new", "marqo__int_fields": {"bucket": 7}}, {"body": (["c"], extra)}),
                   _doc("d7", {"marqo__id": "d7", "marqo__int_fields": {"bucket": 8}}, {"body": (["c"], vecs[7:8])})], "s1")
    ix.delete_batch(["d57"], "s1")
    assert entry.tag != tag
    res2 = ix.query(_yql("s1", ["body"], 10) + f" AND {flt}", hits=10, ranking="embedding_similarity", model_restrict="s1",
                    query_features={"marqo__query_embedding": q.tolist()})
    assert {h.id.split("::")[-1] for h in res2.hits} == {"d107", "d157", "new"}


def test_feed_batch_is_atomic_per_document(monkeypatch):
    """a rejected document leaves no trace — an update with a wrong dimension keeps the old version
    searchable, a rejected new document is not registered, a multi-field document fails as a whole."""
    import numpy as np
    import marqo_b200.gpu_tensor_index as gti
    from _filter_scenario import _doc, _yql
    monkeypatch.setattr(gti, "RowStore", _NumpyRowStore)
    rng = np.random.default_rng(6)
    v = rng.standard_normal((4, 64)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    ix = gti.GpuTensorIndex()
    assert not ix.feed_batch([_doc("a", {"marqo__id": "a", "t": "old"}, {"body": (["c"], v[0:1]), "title": (["c"], v[1:2])})],
                             "s1").errors
    bad_dim = rng.standard_normal((1, 128)).astype(np.float32)
    r = ix.feed_batch([_doc("a", {"marqo__id": "a", "t": "new"}, {"body": (["c"], v[2:3]), "title": (["c"], bad_dim)}),
                       _doc("b", {"marqo__id": "b"}, {"body": (["c"], bad_dim)}),
                       _doc("c", {"marqo__id": "c"}, {"body": (["c"], np.array([[np.nan] * 64], np.float32))}),
                       _doc("d", {"marqo__id": "d"}, {"body": (["c"], v[3:4])})], "s1")
    assert [x.status for x in r.responses] == [400, 400, 400, 200] and r.errors
    assert ix.get_document_count("s1") == 2
    got = ix.get_batch(["a", "b", "c", "d"], "s1")
    assert [x.status for x in got.responses] == [200, 404, 404, 200]
    assert got.responses[0].document.fields["t"] == "old"
    res = ix.query(_yql("s1", ["body"], 5), hits=5, ranking="embedding_similarity", model_restrict="s1",
                   query_features={"marqo__query_embedding": v[0].tolist()})
    assert [h.id.split("::")[-1] for h in res.hits][0] == "a" and abs(res.hits[0].relevance - 1.0) < 2e-3
    # the same id twice in one batch: the last put wins, both answer 200
    r = ix.feed_batch([_doc("e", {"marqo__id": "e", "v": 1}, {"body": (["c"], v[1:2])}),
                       _doc("e", {"marqo__id": "e", "v": 2}, {"body": (["c"], v[2:3])})], "s1")
    assert [x.status for x in r.responses] == [200, 200]
    assert ix.get_batch(["e"], "s1").responses[0].document.fields["v"] == 2
    s = ix._schemas["s1"]
    assert len(s.doc_rows[s.doc_num["e"]]["marqo__embeddings_body"]) == 1


def test_update_heavy_feed_compacts_the_matrix(monkeypatch):
    """Replaced versions are tombstoned by row, and the dead rows are squeezed out once they are a sizeable share."""
    import numpy as np
    import marqo_b200.gpu_tensor_index as gti
    from _filter_scenario import _doc, _yql
    monkeypatch.setattr(gti, "RowStore", _NumpyRowStore)
    monkeypatch.setattr(gti.GpuTensorIndex, "COMPACT_MIN_DEAD", 8)
    rng = np.random.default_rng(7)
    v = rng.standard_normal((40, 64)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    ix = gti.GpuTensorIndex()
    ix.feed_batch([_doc(f"d{i}", {"marqo__id": f"d{i}"}, {"body": (["a", "b"], v[2 * i:2 * i + 2])}) for i in range(10)], "s1")
    s = ix._schemas["s1"]
    store = s.stores["marqo__embeddings_body"]
    for rnd in range(3):
        ix.feed_batch([_doc(f"d{i}", {"marqo__id": f"d{i}", "round": rnd}, {"body": (["a", "b"], v[2 * i + 20:2 * i + 22])})
                       for i in range(10)], "s1")
    assert s.epoch >= 1 and len(store) < 20 + 3 * 20               # compaction ran
    assert len(store) - s.dead["marqo__embeddings_body"] == 20
    res = ix.query(_yql("s1", ["body"], 3), hits=3, ranking="embedding_similarity", model_restrict="s1",
                   query_features={"marqo__query_embedding": v[27].tolist()})
    top = res.hits[0]
    assert top.id.endswith("::d3") and abs(top.relevance - 1.0) < 2e-3
    assert top.dict()["fields"]["matchfeatures"]["closest(marqo__embeddings_body)"]["cells"] == {"1": 1.0}
    got = ix.get_batch(["d3"], "s1").responses[0].document.fields
    assert got["round"] == 2 and list(got["marqo__embeddings_body"]) == ["0", "1"]


def test_concurrent_queries_share_scans(monkeypatch):
    """Marqo sends one query per request from up to 8 threads (api/configs.py:27-28): requests that are in flight
    together and agree on (row store, modifiers, filter) are gathered into one scan; every caller gets its own hits."""
    import threading
    import time
    import numpy as np
    import marqo_b200.gpu_tensor_index as gti
    from _filter_scenario import _doc, _yql

    class _Slow(_NumpyRowStore):
        def search(self, q, k, **kw):
            time.sleep(0.02)                                       # a scan in flight: later arrivals pile up behind it
            return super().search(q, k, **kw)

    monkeypatch.setattr(gti, "RowStore", _Slow)
    rng = np.random.default_rng(8)
    n = 64
    v = rng.standard_normal((n, 64)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    ix = gti.GpuTensorIndex(coalesce_window_s=0.01)
    ix.feed_batch([_doc(f"d{i}", {"marqo__id": f"d{i}"}, {"body": (["c"], v[i:i + 1])}) for i in range(n)], "s1")
    out = {}

    def ask(i):
        res = ix.query(_yql("s1", ["body"], 3 + i % 3), hits=3 + i % 3, ranking="embedding_similarity", model_restrict="s1",
                       query_features={"marqo__query_embedding": v[i].tolist()})
        out[i] = [h.id.split("::")[-1] for h in res.hits]

    threads = [threading.Thread(target=ask, args=(i,)) for i in range(24)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    for i in range(24):
        assert out[i][0] == f"d{i}" and len(out[i]) == 3 + i % 3
    st = ix.coalescer_stats()
    assert st["queries"] == 24 and st["batches"] < 24              # at least some requests shared a scan
    # a lone caller is not delayed by the gathering window and still gets the right answer
    ask(5)
    assert out[5][0] == "d5"


def test_index_scenarios_on_the_cpu_stand_in(monkeypatch):
    """The GPU adapter tests' scenarios (feed / query / highlights / offset / overwrite / get_batch / delete / bad documents;
    score modifiers incl. the refused negative multiplier) against the numpy stand-in."""
    import marqo_b200.gpu_tensor_index as gti
    from marqo_b200 import _native as N
    from _filter_scenario import run_feed_query_scenario, run_score_modifier_scenario

    class _Store(_NumpyRowStore):
        def search(self, q, k, mult=(), add=(), **kw):
            # the engine refuses negative multipliers on indexes with explicit document ids (b200_index_search_ex)
            if any(w * v < 0 for col, w in mult for v in self.attrs.get(col, {}).values()):
                raise N.NativeError(N.ERR_UNSUPPORTED, "negative multiplier")
            return super().search(q, k, mult=mult, add=add, **kw)

    monkeypatch.setattr(gti, "RowStore", _Store)
    run_feed_query_scenario()
    run_score_modifier_scenario()


def test_device_chunks_feed_coalesces_into_one_append(monkeypatch):
    """add_documents fast path host logic: DeviceChunks of one batch that are consecutive in device memory become ONE
    row append per tensor field; mixed host / device documents keep feed order; rejected documents leave no rows."""
    import numpy as np
    import marqo_b200.gpu_tensor_index as gti
    from marqo_b200.gpu_tensor_index import DeviceChunks
    from _filter_scenario import _doc, _yql

    class _DevStore(_NumpyRowStore):
        """'device memory' = a dict of fake pointers -> numpy rows"""
        heap = {}
        calls = []

        def add_device_docs(self, ptr, doc_ids):
            base, off = max((b, ptr - b) for b in self.heap if b <= ptr)
            rows = self.heap[base][off // (self.dim * 4):][:len(doc_ids)]
            assert len(rows) == len(doc_ids)
            type(self).calls.append(len(doc_ids))
            self.add(rows, doc_ids)

    monkeypatch.setattr(gti, "RowStore", _DevStore)
    rng = np.random.default_rng(9)
    dim = 64
    emb = rng.standard_normal((6, dim)).astype(np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    _DevStore.heap = {1 << 20: emb}
    _DevStore.calls = []
    ix = gti.GpuTensorIndex()
    batch = [{"id": f"d{i}", "fields": {"marqo__id": f"d{i}", "n": i,
                                       "marqo__embeddings_img": DeviceChunks(["0"], (1 << 20) + i * dim * 4, dim, emb)}}
             for i in range(4)]
    batch.append({"id": "two", "fields": {"marqo__id": "two",
                                         "marqo__embeddings_img": DeviceChunks(["0", "1"], (1 << 20) + 4 * dim * 4, dim, emb)}})
    r = ix.feed_batch(batch, "s1")
    assert not r.errors and _DevStore.calls == [6]                    # ONE device append for the whole batch
    res = ix.query(_yql("s1", ["img"], 2), hits=2, ranking="embedding_similarity", model_restrict="s1",
                   query_features={"marqo__query_embedding": emb[5].tolist()})
    assert res.hits[0].id.endswith("::two")
    assert res.hits[0].dict()["fields"]["matchfeatures"]["closest(marqo__embeddings_img)"]["cells"] == {"1": 1.0}
    # wrong dimension on the device path is rejected up front, like a host document
    bad = ix.feed_batch([{"id": "bad", "fields": {"marqo__embeddings_img": DeviceChunks(["0"], 1 << 20, 128, emb)}}], "s1")
    assert bad.errors and bad.responses[0].status == 400 and ix.get_document_count("s1") == 5
    # host and device documents interleaved: rows stay in feed order
    _DevStore.calls = []
    mixed = [_doc("h1", {"marqo__id": "h1"}, {"img": (["c"], emb[0:1])}),
             {"id": "g1", "fields": {"marqo__id": "g1", "marqo__embeddings_img": DeviceChunks(["0"], (1 << 20) + dim * 4, dim, emb)}},
             _doc("h2", {"marqo__id": "h2"}, {"img": (["c"], emb[2:3])})]
    assert not ix.feed_batch(mixed, "s1").errors
    s = ix._schemas["s1"]
    order = [s.doc_ids[n] for n, _ in s.row_chunk["marqo__embeddings_img"][-3:]]
    assert order == ["h1", "g1", "h2"] and _DevStore.calls == [1]
