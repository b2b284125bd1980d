"""Shared scenarios for GpuTensorIndex (feed / query / highlights / overwrite / delete, score modifiers, filters): run against the real row store on the GPU
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


def run_feed_query_scenario():
    from marqo_b200.gpu_tensor_index import GpuTensorIndex, gather_documents_from_response
    rng = np.random.default_rng(3)
    D, n = 64, 40

    def unit(m):
        x = rng.standard_normal((m, D)).astype(np.float32)
        return x / np.linalg.norm(x, axis=1, keepdims=True)

    ix = GpuTensorIndex()
    docs, title_vecs, body_vecs = [], {}, {}
    for i in range(n):
        tv, bv = unit(1), unit(int(rng.integers(1, 4)))
        title_vecs[f"d{i}"], body_vecs[f"d{i}"] = tv, bv
        docs.append(_doc(f"d{i}", {"marqo__id": f"d{i}", "price": i},
                         {"title": ([f"title {i}"], tv), "body": ([f"body {i} chunk {j}" for j in range(len(bv))], bv)}))
    resp = ix.feed_batch(docs, "s1")
    assert not resp.errors and len(resp.responses) == n and resp.responses[0].status == 200
    assert resp.responses[3].id == "id:s1:s1::d3"                                 # parsed with split('::')[-1]

    q = body_vecs["d7"][-1] * 0.9 + 0.1 * unit(1)[0]
    q /= np.linalg.norm(q)
    qf = {"marqo__query_embedding": q.tolist()}
    res = ix.query(_yql("s1", ["title", "body"], 5), hits=5, ranking="embedding_similarity", model_restrict="s1",
                   query_features=qf)
    # brute-force expectation on the fp16-rounded store: max over fields and chunks of 1/(2 - q.e)
    qh = q.astype(np.float16).astype(np.float64)
    exp = {}
    for did in title_vecs:
        allv = np.concatenate([title_vecs[did], body_vecs[did]]).astype(np.float16).astype(np.float64)
        exp[did] = float((1.0 / (2.0 - allv @ qh)).max())
    order = sorted(exp, key=lambda d: -exp[d])[:5]
    assert [h.id.split("::")[-1] for h in res.hits] == order
    assert abs(res.hits[0].relevance - exp[order[0]]) < 1e-9
    assert res.root.coverage.coverage == 100
    out = gather_documents_from_response(res)
    assert out["hits"][0]["_id"] == "d7" and out["hits"][0]["price"] == 7
    last = len(body_vecs["d7"]) - 1
    assert out["hits"][0]["_highlights"] == [{"body": f"body 7 chunk {last}"}]
    # single-field query only searches that field
    res_t = ix.query(_yql("s1", ["title"], 3), hits=3, ranking="embedding_similarity", model_restrict="s1",
                     query_features={"marqo__query_embedding": title_vecs["d11"][0].tolist()})
    assert res_t.hits[0].id.endswith("::d11") and abs(res_t.hits[0].relevance - 1.0) < 2e-3
    # offset
    res_o = ix.query(_yql("s1", ["title", "body"], 5), hits=3, offset=2, ranking="embedding_similarity",
                     model_restrict="s1", query_features=qf)
    assert [h.id.split("::")[-1] for h in res_o.hits] == order[2:5]
    # overwrite by id: the old vectors must stop matching
    new = unit(1)
    ix.feed_batch([_doc("d7", {"marqo__id": "d7", "price": 700}, {"title": (["new title"], new)})], "s1")
    res2 = ix.query(_yql("s1", ["title", "body"], 5), hits=5, ranking="embedding_similarity", model_restrict="s1",
                    query_features=qf)
    assert "d7" not in [h.id.split("::")[-1] for h in res2.hits[:1]]
    got = ix.get_batch(["d7", "nope"], "s1")
    assert got.responses[0].status == 200 and got.responses[1].status == 404
    f7 = got.responses[0].document.fields
    assert f7["price"] == 700 and "marqo__embeddings_body" not in f7
    np.testing.assert_array_equal(np.asarray(f7["marqo__embeddings_title"]["0"], np.float32),
                                  new[0].astype(np.float16).astype(np.float32))
    # delete
    ix.delete_batch(["d11"], "s1")
    res3 = ix.query(_yql("s1", ["title"], 3), hits=3, ranking="embedding_similarity", model_restrict="s1",
                    query_features={"marqo__query_embedding": title_vecs["d11"][0].tolist()})
    assert all(not h.id.endswith("::d11") for h in res3.hits)
    assert ix.get_document_count("s1") == n - 1
    # a bad document does not fail the batch
    bad = ix.feed_batch([{"id": "x", "fields": {"marqo__embeddings_title": {"0": [1.0] * 8}}},
                         _doc("ok", {}, {"title": (["t"], unit(1))})], "s1")
    assert bad.errors and bad.responses[0].status == 400 and bad.responses[1].status == 200
    ix.close()


def run_score_modifier_scenario():
    """Tensor search with score_modifiers (tensor_search.py -> vespa_index.py:106-150): the query carries
    marqo__mult_weights_tensor / marqo__add_weights_tensor, documents carry marqo__score_modifiers."""
    from marqo_b200.errors import VespaError
    import pytest
    from marqo_b200.gpu_tensor_index import GpuTensorIndex, gather_documents_from_response
    rng = np.random.default_rng(11)
    D, n = 64, 60

    def unit(m):
        x = rng.standard_normal((m, D)).astype(np.float32)
        return x / np.linalg.norm(x, axis=1, keepdims=True)

    ix = GpuTensorIndex()
    vecs, attrs, docs = {}, {}, []
    for i in range(n):
        v = unit(int(rng.integers(1, 3)))
        a = {}
        if i % 3:
            a["popularity"] = float(rng.uniform(0.5, 3.0))
        if i % 2:
            a["meta.rating"] = float(rng.integers(1, 6))
        vecs[f"d{i}"], attrs[f"d{i}"] = v, a
        docs.append(_doc(f"d{i}", {"marqo__id": f"d{i}", "marqo__score_modifiers": a},
                         {"body": ([f"chunk {j}" for j in range(len(v))], v)}))
    assert not ix.feed_batch(docs, "s1").errors
    q = unit(1)[0]
    qh = q.astype(np.float16).astype(np.float64)
    mult, add = {"popularity": 1.5, "nobody_has_this": 4.0}, {"meta.rating": 0.02}

    def expected(did):
        c = float((1.0 / (2.0 - vecs[did].astype(np.float16).astype(np.float64) @ qh)).max())
        a = attrs[did]
        m = 1.5 * a["popularity"] if "popularity" in a else 1.0
        return m * c + 0.02 * a.get("meta.rating", 0.0), c

    qf = {"marqo__query_embedding": q.tolist(), "marqo__mult_weights_tensor": mult, "marqo__add_weights_tensor": add}
    res = ix.query(_yql("s1", ["body"], 10), hits=10, ranking="embedding_similarity", model_restrict="s1",
                   query_features=qf)
    order = sorted(vecs, key=lambda d: (-expected(d)[0], int(d[1:])))[:10]
    assert [h.id.split("::")[-1] for h in res.hits] == order
    for h, did in zip(res.hits, order):
        assert abs(h.relevance - expected(did)[0]) < 1e-9
        dist = h.dict()["fields"]["matchfeatures"]["distance(field,marqo__embeddings_body)"]
        assert abs(dist - (1.0 / expected(did)[1] - 1.0)) < 1e-6           # distance() stays the RAW distance
    assert gather_documents_from_response(res)["hits"][0]["_score"] == res.hits[0].relevance
    # pre-2.10 index versions use another rank profile name and input names (vespa_index.py:132-136)
    res29 = ix.query(_yql("s1", ["body"], 10), hits=10, ranking="embedding_similarity_modifiers", model_restrict="s1",
                     query_features={"marqo__query_embedding": q.tolist(), "marqo__mult_weights": mult,
                                     "marqo__add_weights": add})
    assert [h.id for h in res29.hits] == [h.id for h in res.hits]
    # overwriting a document replaces its modifier cells
    best = order[0]
    ix.feed_batch([_doc(best, {"marqo__id": best, "marqo__score_modifiers": {}},
                        {"body": (["c"], vecs[best][:1])})], "s1")
    vecs[best], attrs[best] = vecs[best][:1], {}
    res2 = ix.query(_yql("s1", ["body"], 10), hits=10, ranking="embedding_similarity", model_restrict="s1",
                    query_features=qf)
    order2 = sorted(vecs, key=lambda d: (-expected(d)[0], int(d[1:])))[:10]
    assert [h.id.split("::")[-1] for h in res2.hits] == order2
    # a negative multiplier cannot be answered exactly for multi-chunk documents: delegate or refuse
    with pytest.raises(VespaError):
        ix.query(_yql("s1", ["body"], 10), hits=10, ranking="embedding_similarity", model_restrict="s1",
                 query_features={"marqo__query_embedding": q.tolist(), "marqo__mult_weights_tensor": {"popularity": -1.0}})
    # lexical modifier tensors belong to bm25 / hybrid profiles: not a tensor query
    with pytest.raises(VespaError):
        ix.query(_yql("s1", ["body"], 10), hits=10, ranking="embedding_similarity", model_restrict="s1",
                 query_features={"marqo__query_embedding": q.tolist(), "marqo__mult_weights_lexical": {"popularity": 1.0}})
    ix.close()
