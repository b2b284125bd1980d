"""Shared scenario for filtered tensor search through GpuTensorIndex: run against the real row store on the GPU
(tests/test_adapters_gpu.py) and against a numpy stand-in on the CPU (tests/test_yql_filter.py), so the adapter's host
logic — filter recognition, deeper fetches, merge, offset, modifiers — is checked even where no device exists."""
import numpy as np


def _doc(doc_id, fields, embs):
    f = dict(fields)
    for name, (chunks, vecs) in embs.items():
        f[f"marqo__chunks_{name}"] = chunks
        f[f"marqo__embeddings_{name}"] = {str(i): v.tolist() for i, v in enumerate(vecs)}
    return {"id": doc_id, "fields": f}


def _yql(schema, fields, k):
    terms = " OR ".join(f"({{targetHits:{k}, approximate:False, hnsw.exploreAdditionalHits:0}}"
                        f"nearestNeighbor(marqo__embeddings_{f}, marqo__query_embedding))" for f in fields)
    return f"select * from {schema} where ({terms})"


def run_filtered_search_scenario():
    """Tensor search with a filter (tensor_search.py -> unstructured_vespa_index.py:59-66,135-226): exact top-k among
    the documents the filter keeps, filter text taken from the reference's generator (tests/golden/filter_golden.json)."""
    import json
    from pathlib import Path
    from marqo_b200.errors import VespaError
    import pytest
    from marqo_b200.gpu_tensor_index import GpuTensorIndex
    gold = {g["filter"]: g["yql"] for g in
            json.loads((Path(__file__).resolve().parent / "golden" / "filter_golden.json").read_text())}
    rng = np.random.default_rng(31)
    D, n = 64, 300

    def unit(m):
        x = rng.standard_normal((m, D)).astype(np.float32)
        return x / np.linalg.norm(x, axis=1, keepdims=True)

    ix = GpuTensorIndex()
    vecs, meta, docs = {}, {}, []
    for i in range(n):
        v = unit(1)
        color = ["red", "blue", "green", "Red"][i % 4]
        price = int(rng.integers(0, 40))
        fields = {"marqo__id": f"d{i}", "marqo__short_string_fields": {"color": color},
                  "marqo__int_fields": {"price": price}, "marqo__bool_fields": {"in_stock": int(i % 3 == 0)},
                  "marqo__string_array": [f"tags::{'sale' if i % 5 == 0 else 'full'}"],
                  "marqo__score_modifiers": {"pop": float(i % 7 + 1)}}
        vecs[f"d{i}"], meta[f"d{i}"] = v, dict(color=color.lower(), price=price, in_stock=i % 3 == 0, sale=i % 5 == 0, pop=i % 7 + 1)
        docs.append(_doc(f"d{i}", fields, {"body": (["c"], v)}))
    ix.feed_batch(docs, "s1")
    q = unit(1)[0]
    qh = q.astype(np.float16).astype(np.float64)
    score = {d: float(1.0 / (2.0 - vecs[d][0].astype(np.float16).astype(np.float64) @ qh)) for d in vecs}

    def ask(filter_yql, hits=10, **qf):
        yql = _yql("s1", ["body"], hits) + f" AND {filter_yql}"
        return ix.query(yql, hits=hits, ranking="embedding_similarity", model_restrict="s1",
                        query_features=dict({"marqo__query_embedding": q.tolist()}, **qf))

    cases = {
        "color:red": lambda m: m["color"] == "red",                       # case-insensitive: 'Red' matches too
        "(color:red OR color:blue) AND price:[0 TO 100]": lambda m: m["color"] in ("red", "blue"),
        "price:[10 TO 20]": lambda m: 10 <= m["price"] <= 20,
        "NOT (color:red AND in_stock:true)": lambda m: not (m["color"] == "red" and m["in_stock"]),
        "tags:sale": lambda m: m["sale"],
        "_id:doc7": lambda m: False,
    }
    for f, keep in cases.items():
        want = sorted((d for d in vecs if keep(meta[d])), key=lambda d: (-score[d], int(d[1:])))[:10]
        res = ask(gold[f])
        assert [h.id.split("::")[-1] for h in res.hits] == want, f
        for h, d in zip(res.hits, want):
            assert abs(h.relevance - score[d]) < 1e-9
    # a selective filter forces deeper fetches: in_stock AND sale AND price <= 3
    sel = f"(({gold['in_stock:true']} AND {gold['tags:sale']}) AND {gold['price:[* TO 3]']})"
    want = sorted((d for d in vecs if meta[d]["in_stock"] and meta[d]["sale"] and meta[d]["price"] <= 3),
                  key=lambda d: (-score[d], int(d[1:])))[:10]
    assert [h.id.split("::")[-1] for h in ask(sel).hits] == want
    # filter + score modifiers + offset
    wantm = sorted((d for d in vecs if meta[d]["color"] == "red"), key=lambda d: (-(0.5 * meta[d]["pop"] * score[d]), int(d[1:])))
    yql = _yql("s1", ["body"], 5) + f" AND {gold['color:red']}"
    resm = ix.query(yql, hits=5, offset=3, ranking="embedding_similarity", model_restrict="s1",
                    query_features={"marqo__query_embedding": q.tolist(), "marqo__mult_weights_tensor": {"pop": 0.5}})
    assert [h.id.split("::")[-1] for h in resm.hits] == wantm[3:8]
    # another grammar: not answered
    with pytest.raises(VespaError):
        ix.query(_yql("s1", ["body"], 5) + ' AND (title matches "x")', hits=5, ranking="embedding_similarity",
                 model_restrict="s1", query_features={"marqo__query_embedding": q.tolist()})
    ix.close()


