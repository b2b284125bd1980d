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
    order (score desc, doc asc), best chunk per document) — lets the adapter's host logic run without a device."""

    def __init__(self, dim, metric="prenormalized-angular", device=0, capacity=0):
        self.dim, self.rows, self.docs, self.attrs = dim, [], [], {}

    def __len__(self):
        return len(self.rows)

    def add(self, vecs, doc_ids=None):
        import numpy as np
        for i, v in enumerate(np.asarray(vecs, np.float32)):
            self.rows.append(v.astype(np.float16).astype(np.float64))
            self.docs.append(int(doc_ids[i]) if doc_ids is not None else len(self.docs))

    def delete_doc(self, doc_id):
        self.docs = [-1 if d == doc_id else d for d in self.docs]

    def set_attributes(self, column, doc_ids, values):
        for i, d in enumerate(doc_ids):
            if column == -1:
                for col in self.attrs.values():
                    col.pop(int(d), None)
            elif values is None:
                self.attrs.setdefault(column, {}).pop(int(d), None)
            else:
                self.attrs.setdefault(column, {})[int(d)] = float(values[i])

    def _rank(self, q, k, mult=(), add=()):
        import numpy as np
        qh = np.asarray(q, np.float32).reshape(-1).astype(np.float16).astype(np.float64)
        best = {}
        for r, (v, d) in enumerate(zip(self.rows, self.docs)):
            if d < 0:
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
        doc = np.full((1, k), -1, np.int32)
        row = np.full((1, k), -1, np.int32)
        score = np.full((1, k), -np.inf)
        for i, (sc, d, r) in enumerate(out[:k]):
            doc[0, i], row[0, i], score[0, i] = d, r, sc
        return doc, row, score

    def search(self, q, k):
        return self._rank(q, k)

    def search_modified(self, q, k, mult=(), add=()):
        return self._rank(q, k, mult, add)

    def get_row(self, r):
        return self.rows[r].astype("float32")

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


def test_very_selective_filters_switch_to_one_masked_scan(monkeypatch):
    """When deeper fetches hit their cap the adapter evaluates the filter over the whole schema and excludes documents
    through the reserved attribute column — same exact answer, one scan.  The cap is lowered to force that path."""
    import numpy as np
    import marqo_b200.gpu_tensor_index as gti
    from _filter_scenario import _doc, _yql
    monkeypatch.setattr(gti, "RowStore", _NumpyRowStore)
    monkeypatch.setattr(gti.GpuTensorIndex, "MAX_FETCH", 16)
    rng = np.random.default_rng(5)
    n = 200
    vecs = rng.standard_normal((n, 64)).astype(np.float32)
    vecs /= np.linalg.norm(vecs, axis=1, keepdims=True)
    ix = gti.GpuTensorIndex()
    docs = [_doc(f"d{i}", {"marqo__id": f"d{i}", "marqo__int_fields": {"bucket": i % 50},
                           "marqo__score_modifiers": {"pop": float(i % 3 + 1)}}, {"body": (["c"], vecs[i:i + 1])})
            for i in range(n)]
    assert not ix.feed_batch(docs, "s1").errors
    q = vecs[3]
    qh = q.astype(np.float16).astype(np.float64)
    c = {f"d{i}": 1.0 / (2.0 - float(vecs[i].astype(np.float16).astype(np.float64) @ qh)) for i in range(n)}
    flt = '((marqo__int_fields contains sameElement(key contains "bucket", value = 7)) OR ' \
          '(marqo__float_fields contains sameElement(key contains "bucket", value = 7)))'        # 4 of 200 documents
    res = ix.query(_yql("s1", ["body"], 10) + f" AND {flt}", hits=10, ranking="embedding_similarity", model_restrict="s1",
                   query_features={"marqo__query_embedding": q.tolist()})
    want = sorted((d for d in c if int(d[1:]) % 50 == 7), key=lambda d: (-c[d], int(d[1:])))
    assert [h.id.split("::")[-1] for h in res.hits] == want and len(want) == 4
    assert all(abs(h.relevance - c[h.id.split("::")[-1]]) < 1e-12 for h in res.hits)
    # together with score modifiers; and the mask column is gone afterwards (an unfiltered query sees every document)
    resm = ix.query(_yql("s1", ["body"], 3) + f" AND {flt}", hits=3, ranking="embedding_similarity", model_restrict="s1",
                    query_features={"marqo__query_embedding": q.tolist(), "marqo__add_weights_tensor": {"pop": 0.5}})
    wantm = sorted(want, key=lambda d: (-(c[d] + 0.5 * (int(d[1:]) % 3 + 1)), int(d[1:])))[:3]
    assert [h.id.split("::")[-1] for h in resm.hits] == wantm
    plain = ix.query(_yql("s1", ["body"], 5), hits=5, ranking="embedding_similarity", model_restrict="s1",
                     query_features={"marqo__query_embedding": q.tolist()})
    assert [h.id.split("::")[-1] for h in plain.hits] == sorted(c, key=lambda d: (-c[d], int(d[1:])))[:5]
    store = next(iter(ix._schemas["s1"].stores.values()))
    assert not store.attrs.get(gti.GpuTensorIndex.MASK_COLUMN)


def test_index_scenarios_on_the_cpu_stand_in(monkeypatch):
    """The GPU adapter tests' scenarios (feed / query / highlights / offset / overwrite / get_batch / delete / bad documents;
    score modifiers incl. the refused negative multiplier) against the numpy stand-in."""
    import marqo_b200.gpu_tensor_index as gti
    from marqo_b200 import _native as N
    from _filter_scenario import run_feed_query_scenario, run_score_modifier_scenario

    class _Store(_NumpyRowStore):
        def search_modified(self, q, k, mult=(), add=()):
            # the engine refuses negative multipliers on indexes with explicit document ids (b200_index_search_modified)
            if any(w * v < 0 for col, w in mult for v in self.attrs.get(col, {}).values()):
                raise N.NativeError(N.ERR_UNSUPPORTED, "negative multiplier")
            return super().search_modified(q, k, mult, add)

    monkeypatch.setattr(gti, "RowStore", _Store)
    run_feed_query_scenario()
    run_score_modifier_scenario()
