"""CPU checks of the score-modifier oracle (oracle/score_oracle.c: oracle_modifiers / oracle_search_modified) against a
plain numpy statement of the reference's rank expression (unstructured_vespa_schema.py:266-271):

    if (count(mult_weights * attr) == 0, 1, reduce(mult_weights * attr, prod)) * score + reduce(add_weights * attr, sum)

and of the adapter's host-side arithmetic that must agree with it."""
import numpy as np


def _numpy_modify(attrs, mult, add, closeness):
    n_docs = attrs.shape[1]
    out = np.empty(n_docs)
    for d in range(n_docs):
        cells = [w * attrs[c, d] for c, w in mult if not np.isnan(attrs[c, d])]
        m = float(np.prod(cells)) if cells else 1.0
        a = sum(w * attrs[c, d] for c, w in add if not np.isnan(attrs[c, d]))
        out[d] = m * closeness[d] + a
    return out


def test_oracle_modifiers_formula(score_oracle):
    rng = np.random.default_rng(0)
    attrs = rng.uniform(-2, 2, size=(4, 500))
    attrs[rng.random(attrs.shape) > 0.6] = np.nan
    mult, add = [(0, 1.5), (2, -0.5)], [(1, 0.25), (3, 2.0), (0, -1.0)]
    mod = score_oracle.modifiers(attrs, mult, add)
    c = rng.uniform(0.3, 1.0, size=500)
    np.testing.assert_allclose(mod[:, 0] * c + mod[:, 1], _numpy_modify(attrs, mult, add, c), rtol=1e-14, atol=1e-14)
    none = np.isnan(attrs[0]) & np.isnan(attrs[2])
    assert np.all(mod[none, 0] == 1.0)                 # count == 0 -> 1
    assert np.all(score_oracle.modifiers(attrs, [], [])[:, 0] == 1.0)
    assert np.all(score_oracle.modifiers(attrs, [], [])[:, 1] == 0.0)


def test_oracle_search_modified_matches_bruteforce(score_oracle):
    rng = np.random.default_rng(1)
    n, d, k = 600, 64, 10
    corpus = rng.standard_normal((n, d)).astype(np.float32)
    corpus /= np.linalg.norm(corpus, axis=1, keepdims=True)
    q = corpus[:3] + 0.1 * rng.standard_normal((3, d)).astype(np.float32)
    doc_of_row = (np.arange(n) // 2).astype(np.int32)
    attrs = rng.uniform(0, 2, size=(2, n // 2))
    attrs[rng.random(attrs.shape) > 0.7] = np.nan
    mult, add = [(0, 2.0)], [(1, 0.05)]
    mod = score_oracle.modifiers(attrs, mult, add)
    doc, row, score = score_oracle.search_modified(q, corpus, k, mod, "prenormalized-angular", doc_of_row)
    ch = score_oracle.half_to_float(score_oracle.to_half(corpus)).astype(np.float64)
    qh = score_oracle.half_to_float(score_oracle.to_half(q)).astype(np.float64)
    for i in range(3):
        dots = ch @ qh[i]
        best = np.maximum(dots[0::2], dots[1::2])          # closeness of the best chunk, THEN modify
        clos = 1.0 / (1.0 + (1.0 - best))
        want = _numpy_modify(attrs, mult, add, clos)
        order = np.lexsort((np.arange(n // 2), -want))[:k]
        np.testing.assert_array_equal(doc[i], order)
        np.testing.assert_allclose(score[i], want[order], rtol=0, atol=1e-9)
        np.testing.assert_array_equal(row[i] // 2, order)


def test_adapter_host_modifier_arithmetic():
    from marqo_b200.gpu_tensor_index import GpuTensorIndex
    m, a = GpuTensorIndex._modifier_of({"pop": 2.0, "price": 10.0}, {"pop": 1.5, "absent": 9.0}, {"price": -0.01})
    assert m == 3.0 and a == -0.1
    assert GpuTensorIndex._modifier_of({}, {"pop": 1.5}, {"price": 1.0}) == (1.0, 0.0)
    assert GpuTensorIndex._weights({"cells": [{"address": {"p": "f"}, "value": 2}]}) == {"f": 2.0}
    assert GpuTensorIndex._weights({"cells": {"f": 2}}) == {"f": 2.0}
    assert GpuTensorIndex._weights(None) == {}
