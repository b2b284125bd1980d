/* ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product path
 * (marqo_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use it, and only as the checker / the reported CPU baseline.
 *
 * CPU restatement of the tensor-search score step that the reference delegates to Vespa 8.332.5
 * (vespa/pom.xml:147; NOT vendored under /root/reference, so its published semantics are restated):
 *
 *   - query: `{targetHits:..., approximate:false} nearestNeighbor(field, marqo__query_embedding)`
 *     src/marqo/core/unstructured_vespa_index/unstructured_vespa_index.py:109-133
 *   - rank profile `embedding_similarity`: first-phase `closeness(field, embeddings)`; embeddings are
 *     `tensor<float>(p{}, x[D])` (one mapped cell block per chunk) so closeness is the MAX over chunks;
 *     match-features `closest(embeddings)` = arg-max chunk
 *     src/marqo/core/unstructured_vespa_index/unstructured_vespa_schema.py:155-166,225-230,292-294
 *   - closeness = 1 / (1 + distance); distance by `distance-metric`
 *     (src/marqo/core/models/marqo_index.py:63-69): prenormalized-angular 1 - q.e,
 *     angular acos(cos), dotproduct: closeness = q.e (raw), euclidean |q - e|
 *   - pinned by the reference's tests: identical vector => _score == 1.0
 *     (tests/tensor_search/integ_tests/test_custom_vector_field.py:592,602)
 *
 * Parity status: the *ordering* semantics (exact scan, max over chunks, top-k) are pinned by the reference's
 * schema + the `_score == 1.0` tests; Vespa's own tie-break between equal scores is not specified, so this
 * oracle DEFINES the total order (score desc, doc id asc) — "tie order: defined here, not pinned".
 *
 * Storage contract shared with the CUDA path: rows are fp16 (BASELINE.json configs[4]); the score is the
 * fp64 dot product of the fp16-rounded operands accumulated in the fixed order (one warp of the CUDA path, 16-byte
 * loads per lane):
 *     partial[l] = sum over j ascending, then t = 0..7, of q[i] * c[i] with i = 256 j + 8 l + t   (l = 0..31)
 *     dot        = xor-butterfly of the partials: for o in 16, 8, 4, 2, 1: partial[l] += partial[l ^ o]
 * (every product of two fp16 values is exact in fp64, so fused / unfused multiply-add agree; IEEE addition is
 * commutative, so every lane of the butterfly ends with the same bits).
 *
 * TOTAL ORDER (defined here; the CUDA path proves it reproduces it — score.cu header, "GUARD"):
 *   per document: best row = max exact dot, ties -> lowest row;  across documents: (key desc, doc id asc), where
 *   key = exact dot (or minus squared distance, or the modified score).  The CUDA path selects candidates with an
 *   approximate tensor-core key but only returns a result after checking, per query, that the exact key of the k-th
 *   document exceeds (bound of every unexamined row's approximate key) + (bound of the approximation error);
 *   otherwise it re-collects every row above a safe threshold and re-scores all of them.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- IEEE binary16 <-> binary32, round-to-nearest-even ---- */
static float half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu;
    uint32_t man = h & 0x3FFu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {
            int e = -1;
            do {
                man <<= 1;
                ++e;
            } while (!(man & 0x400u));
            man &= 0x3FFu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7F800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

static uint16_t float_to_half(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t abs = x & 0x7FFFFFFFu;
    if (abs >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | (abs > 0x7F800000u ? 0x200u : 0));
    if (abs >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u); /* rounds to inf */
    if (abs < 0x33000001u) return (uint16_t)sign;               /* rounds to zero */
    int32_t e = (int32_t)(abs >> 23) - 127;
    uint32_t m = (abs & 0x7FFFFFu) | 0x800000u;
    if (e < -14) {
        int shift = -14 - e + 13;
        uint32_t hm = m >> shift;
        uint32_t rem = m & ((1u << shift) - 1u);
        uint32_t half = 1u << (shift - 1);
        if (rem > half || (rem == half && (hm & 1u))) ++hm;
        return (uint16_t)(sign | hm);
    }
    uint32_t hm = (m >> 13) & 0x3FFu;
    uint32_t rem = m & 0x1FFFu;
    uint32_t he = (uint32_t)(e + 15);
    uint32_t h = (he << 10) | hm;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
    return (uint16_t)(sign | h);
}

/* fp32 -> fp16 rows; normalize != 0 applies the angular-metric L2 normalisation in the CUDA path's order:
 * lane-strided fp32 FMA partial sums, xor-butterfly combine, 1/sqrt, multiply. */
void oracle_convert_rows(const float* src, uint16_t* dst, int64_t rows, int dim, int normalize) {
    for (int64_t r = 0; r < rows; ++r) {
        const float* s = src + r * dim;
        float scale = 1.0f;
        if (normalize) {
            float part[32];
            for (int l = 0; l < 32; ++l) {
                float ss = 0.0f;
                for (int i = l; i < dim; i += 32) ss = fmaf(s[i], s[i], ss);
                part[l] = ss;
            }
            for (int o = 16; o > 0; o >>= 1) {
                float nxt[32];
                for (int l = 0; l < 32; ++l) nxt[l] = part[l] + part[l ^ o];
                memcpy(part, nxt, sizeof(part));
            }
            scale = part[0] > 0.0f ? 1.0f / sqrtf(part[0]) : 0.0f;
        }
        for (int i = 0; i < dim; ++i) dst[r * dim + i] = float_to_half(s[i] * scale);
    }
}

/* operands pre-widened to fp32 (exact); same summation order as documented above */
static double butterfly32(double* part) {
    for (int o = 16; o > 0; o >>= 1) {
        double nxt[32];
        for (int l = 0; l < 32; ++l) nxt[l] = part[l] + part[l ^ o];
        memcpy(part, nxt, sizeof(nxt));
    }
    return part[0];
}

static double exact_dot_f(const float* q, const float* c, int dim) {
    double part[32];
    for (int l = 0; l < 32; ++l) {
        double p = 0.0;
        for (int base = 8 * l; base < dim; base += 256)
            for (int t = 0; t < 8; ++t) p += (double)q[base + t] * (double)c[base + t];
        part[l] = p;
    }
    return butterfly32(part);
}

/* euclidean: the ordering key is minus the squared distance, sum of exact (q - c)^2 terms in the same fixed order */
static double exact_neg_sqdist_f(const float* q, const float* c, int dim) {
    double part[32];
    for (int l = 0; l < 32; ++l) {
        double p = 0.0;
        for (int base = 8 * l; base < dim; base += 256)
            for (int t = 0; t < 8; ++t) {
                double d = (double)q[base + t] - (double)c[base + t];
                volatile double sq = d * d;   /* unfused, like __dmul_rn / __dsub_rn */
                p = p - sq;
            }
        part[l] = p;
    }
    return butterfly32(part);
}

double oracle_closeness(double dot, int metric) {
    switch (metric) {
        case 3: { /* euclidean: `dot` carries -|q - e|^2 */
            double d2 = -dot;
            return 1.0 / (1.0 + sqrt(d2 > 0.0 ? d2 : 0.0));
        }
        case 0: /* prenormalized-angular */
            return 1.0 / (1.0 + (1.0 - dot));
        case 1: { /* angular */
            double c = dot > 1.0 ? 1.0 : (dot < -1.0 ? -1.0 : dot);
            return 1.0 / (1.0 + acos(c));
        }
        default: /* dotproduct */
            return dot;
    }
}

typedef struct {
    double dot;
    int32_t doc;
    int32_t row;
} hit_t;

static int hit_before(const hit_t* a, const hit_t* b) {
    return a->dot > b->dot || (a->dot == b->dot && a->doc < b->doc);
}

/* Exact search.  qh [nq, dim] fp16, corpus [n, dim] fp16, doc_of_row [n] (NULL = identity; < 0 = deleted row).
 * Outputs [nq, k]: doc, arg-max row, closeness; unused slots -1 / -1 / -inf.
 * Per document: best row = max dot, ties -> lowest row.  Across documents: (dot desc, doc asc). */
static int search_impl(const uint16_t* qh, int nq, const uint16_t* corpus, int64_t n, int dim, const int32_t* doc_of_row,
                       int metric, int k, const double* mod, int32_t* out_doc, int32_t* out_row, double* out_score) {
    int32_t max_doc = -1;
    for (int64_t r = 0; r < n; ++r) {
        int32_t d = doc_of_row ? doc_of_row[r] : (int32_t)r;
        if (d > max_doc) max_doc = d;
    }
    const size_t ndoc = (size_t)(max_doc + 1);
    const int64_t BLK = 4096;
    float* qf = (float*)malloc(sizeof(float) * (size_t)nq * dim);
    float* cf = (float*)malloc(sizeof(float) * (size_t)BLK * dim);
    double* best = (double*)malloc(sizeof(double) * (ndoc ? ndoc : 1) * (size_t)nq);
    int32_t* brow = (int32_t*)malloc(sizeof(int32_t) * (ndoc ? ndoc : 1) * (size_t)nq);
    if (!qf || !cf || !best || !brow) {
        free(qf); free(cf); free(best); free(brow);
        return 1;
    }
    for (size_t i = 0; i < (size_t)nq * dim; ++i) qf[i] = half_to_float(qh[i]);
    for (size_t i = 0; i < ndoc * (size_t)nq; ++i) brow[i] = -1;
    for (int64_t r0 = 0; r0 < n; r0 += BLK) {
        const int64_t nb = n - r0 < BLK ? n - r0 : BLK;
#pragma omp parallel for
        for (int64_t i = 0; i < nb * dim; ++i) cf[i] = half_to_float(corpus[(size_t)r0 * dim + i]);
#pragma omp parallel for schedule(dynamic, 1)
        for (int q = 0; q < nq; ++q) {
            double* bq = best + (size_t)q * ndoc;
            int32_t* rq = brow + (size_t)q * ndoc;
            for (int64_t r = 0; r < nb; ++r) {
                int32_t d = doc_of_row ? doc_of_row[r0 + r] : (int32_t)(r0 + r);
                if (d < 0) continue; /* deleted row */
                double dot = metric == 3 ? exact_neg_sqdist_f(qf + (size_t)q * dim, cf + (size_t)r * dim, dim)
                                         : exact_dot_f(qf + (size_t)q * dim, cf + (size_t)r * dim, dim);
                if (rq[d] < 0 || dot > bq[d]) { /* ties keep the lowest row */
                    bq[d] = dot;
                    rq[d] = (int32_t)(r0 + r);
                }
            }
        }
    }
    int status = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int q = 0; q < nq; ++q) {
        const double* bq = best + (size_t)q * ndoc;
        const int32_t* rq = brow + (size_t)q * ndoc;
        hit_t* top = (hit_t*)malloc(sizeof(hit_t) * (size_t)(k + 1));
        if (!top) {
            status = 1;
            continue;
        }
        int cnt = 0;
        for (int32_t d = 0; d <= max_doc; ++d) {
            if (rq[d] < 0) continue;
            hit_t h = {bq[d], d, rq[d]};
            if (mod) { /* modify(closeness of the best chunk): separate multiply and add, as the CUDA merge does */
                volatile double prod = mod[2 * (size_t)d] * oracle_closeness(bq[d], metric);
                h.dot = prod + mod[2 * (size_t)d + 1];
            }
            if (cnt == k && !hit_before(&h, &top[k - 1])) continue;
            int pos = cnt < k ? cnt : k - 1;
            while (pos > 0 && hit_before(&h, &top[pos - 1])) {
                top[pos] = top[pos - 1];
                --pos;
            }
            top[pos] = h;
            if (cnt < k) ++cnt;
        }
        for (int i = 0; i < k; ++i) {
            size_t o = (size_t)q * k + i;
            if (i < cnt) {
                out_doc[o] = top[i].doc;
                out_row[o] = top[i].row;
                out_score[o] = mod ? top[i].dot : oracle_closeness(top[i].dot, metric);
            } else {
                out_doc[o] = -1;
                out_row[o] = -1;
                out_score[o] = -INFINITY;
            }
        }
        free(top);
    }
    free(qf); free(cf); free(best); free(brow);
    return status;
}

int oracle_search(const uint16_t* qh, int nq, const uint16_t* corpus, int64_t n, int dim, const int32_t* doc_of_row,
                  int metric, int k, int32_t* out_doc, int32_t* out_row, double* out_score) {
    return search_impl(qh, nq, corpus, n, dim, doc_of_row, metric, k, NULL, out_doc, out_row, out_score);
}

/* Score modifiers.  Reference: rank-profile function `modify`
 * (src/marqo/core/unstructured_vespa_index/unstructured_vespa_schema.py:266-271):
 *     if (count(mult_weights * attribute(marqo__score_modifiers)) == 0, 1,
 *         reduce(mult_weights * attribute(marqo__score_modifiers), prod)) * score
 *     + reduce(add_weights * attribute(marqo__score_modifiers), sum)
 * with score = closeness(field, marqo__embeddings) (:292-294), i.e. the closeness of the document's BEST chunk.
 * The tensors are sparse (tensor<double>(p{})): a product cell exists only where the document has the attribute.
 * attrs: [n_cols][n_docs] doubles, NaN = the document has no such cell.  Terms are evaluated in list order (Vespa
 * leaves the reduce order unspecified; fp64 keeps the difference far below any score gap the tests use).
 * out_mod: [n_docs][2] = (mult, add). */
void oracle_modifiers(const double* attrs, int n_cols, int64_t n_docs, const int32_t* mult_cols, const double* mult_w,
                      int n_mult, const int32_t* add_cols, const double* add_w, int n_add, double* out_mod) {
    (void)n_cols;
    for (int64_t d = 0; d < n_docs; ++d) {
        double m = 1.0, a = 0.0;
        int cnt = 0;
        for (int i = 0; i < n_mult; ++i) {
            double v = attrs[(size_t)mult_cols[i] * n_docs + d];
            if (v == v) {
                volatile double t = mult_w[i] * v;
                volatile double mm = m * t;
                m = mm;
                ++cnt;
            }
        }
        if (cnt == 0) m = 1.0;
        for (int i = 0; i < n_add; ++i) {
            double v = attrs[(size_t)add_cols[i] * n_docs + d];
            if (v == v) {
                volatile double t = add_w[i] * v;
                a = a + t;
            }
        }
        out_mod[2 * d] = m;
        out_mod[2 * d + 1] = a;
    }
}

/* oracle_search with modify() applied to every document's best-chunk closeness before the top-k.
 * mod: [>= max_doc + 1][2] from oracle_modifiers.  out_score is the modified score. */
int oracle_search_modified(const uint16_t* qh, int nq, const uint16_t* corpus, int64_t n, int dim,
                           const int32_t* doc_of_row, int metric, int k, const double* mod, int32_t* out_doc,
                           int32_t* out_row, double* out_score) {
    return search_impl(qh, nq, corpus, n, dim, doc_of_row, metric, k, mod, out_doc, out_row, out_score);
}

/* Raw exact dots of one query against a list of rows — used by tests to inspect near-ties. */
void oracle_dots(const uint16_t* qv, const uint16_t* corpus, int dim, const int64_t* rows, int m, double* out) {
    float* qf = (float*)malloc(sizeof(float) * (size_t)dim);
    float* cf = (float*)malloc(sizeof(float) * (size_t)dim);
    for (int i = 0; i < dim; ++i) qf[i] = half_to_float(qv[i]);
    for (int j = 0; j < m; ++j) {
        for (int i = 0; i < dim; ++i) cf[i] = half_to_float(corpus[(size_t)rows[j] * dim + i]);
        out[j] = exact_dot_f(qf, cf, dim);
    }
    free(qf);
    free(cf);
}

/* numpy-facing helpers for the fp16 codec (checked against numpy's own float16 in tests) */
void oracle_half_to_float(const uint16_t* src, float* dst, int64_t n) {
    for (int64_t i = 0; i < n; ++i) dst[i] = half_to_float(src[i]);
}
void oracle_float_to_half(const float* src, uint16_t* dst, int64_t n) {
    for (int64_t i = 0; i < n; ++i) dst[i] = float_to_half(src[i]);
}
