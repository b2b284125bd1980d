// Score + top-k over a GPU-resident fp16 embedding matrix.
//
// Reference semantics (executed inside Vespa today, specified by
// src/marqo/core/unstructured_vespa_index/unstructured_vespa_index.py:59-133 and the rank profile in
// src/marqo/core/unstructured_vespa_index/unstructured_vespa_schema.py:225-230,292-294):
//   score(doc) = max over the doc's chunk rows of closeness(q, row);  top-`hits` documents.
//
// Pipeline per group of <= 64 queries:
//   1. scan_kernel<SELECT>  persistent, one CTA per SM.  The corpus is streamed once from HBM by TMA (128-byte
//        swizzle, 16 KB stages) and multiplied against the smem-resident query block with tcgen05.mma (M = 64
//        queries, N = 128 rows, fp16 x fp16 -> fp32 in TMEM).  Four epilogue warps read the accumulators back
//        (TMEM lane == query) and each thread keeps a sorted register list of the KP best (key, row, doc) of its
//        query.  Rows of a document that is already listed are dropped only when they are PROVABLY not its best
//        chunk (approximate key more than tol = 2 eps below the listed one).
//   2. merge_kernel  one CTA per query: threshold-filters and sorts the per-CTA lists, RE-SCORES the best M
//        candidates exactly (fp64, fixed summation order — the order oracle/score_oracle.c restates), keeps the
//        best chunk per document, ranks documents under (key desc, doc asc) — all in parallel — and then checks
//        the GUARD:   exact key of the k-th document  >  tau + eps
//        where tau bounds the approximate key of every row that was NOT re-scored (max of the full lists' tails
//        and the best unselected list entry) and eps bounds |approximate - exact| for this query.  When the guard
//        holds no unexamined row can belong to (or reorder) the top-k, so the ids are exact.  Otherwise the query
//        is flagged with a threshold L = (k-th exact key) - eps.
//   3. scan_kernel<COLLECT> + finalize_kernel (only for flagged queries; both exit immediately otherwise):
//        a second pass appends EVERY live row whose approximate key is >= L to a per-query buffer; finalize
//        re-scores all of them exactly and repeats step 2's selection.  Excluded rows have exact key < L + eps
//        <= k-th key, so the result is exact whenever the buffer did not overflow; the host loop grows the buffer /
//        lowers L for the pathological remainder (thousands of exact ties, k larger than the per-CTA lists cover).
// The approximate tensor-core key only SELECTS candidates; ids, rows and scores returned always come from the
// exact pass, and the guard makes "the candidate set contained the true top-k" a checked property, not a hope.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace mb {
namespace score {

constexpr int TILE_N = 128;     // corpus rows per tile (UMMA_N)
constexpr int BLOCK_K = 64;     // fp16 elements per 128-byte swizzle row
constexpr int MQ = 64;          // queries per pass (UMMA_M)
constexpr int UMMA_K = 16;
constexpr int KP = 16;          // candidates kept per (CTA, query) list
constexpr int K_SMALL = 10;     // k <= K_SMALL: 64 candidates are re-scored; larger k: up to M_CAP
constexpr int M_SMALL = 64;
constexpr int M_CAP = 384;      // most candidates the merge kernel re-scores exactly
constexpr int K_MERGE_MAX = 160;  // largest k the merge kernel answers itself (needs <= KP * KP list entries)
constexpr int FIN_CAP = 4096;   // most collected rows the device finalize handles (shared memory)
constexpr int MAX_LISTS = 256;  // scan grid clamp (merge_kernel's list-head table)
constexpr int ACC_STAGES = 4;   // TMEM accumulators (4 x 128 fp32 columns = all 512 columns)
constexpr int THREADS = 192;    // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr uint32_t STAGE_BYTES = TILE_N * BLOCK_K * 2;
constexpr uint32_t QCHUNK_BYTES = MQ * BLOCK_K * 2;
constexpr int MAX_DIM = 1024;
constexpr int SMEM_LIMIT = 232448;  // 227 KB
static_assert(K_MERGE_MAX + K_MERGE_MAX / 4 + 16 <= KP * KP, "list-entry threshold trick covers at most KP*KP entries");
static_assert(M_CAP >= K_MERGE_MAX + K_MERGE_MAX / 4 + 16, "M_CAP must cover the candidates needed for K_MERGE_MAX");

enum QueryStatus : int { Q_RESOLVED = 0, Q_NEED = 1 };

// Per-group query state, resident on the device; one D2H copy tells the host everything it needs.
struct QState {
    float eps[MQ];     // bound on |approximate scan key - exact scan key| for any row
    float tol[MQ];     // 2 * eps: same-document chunks closer than this are both kept
    float L[MQ];       // collect threshold (approximate-key domain); +inf = query not flagged
    float qn2[MQ];     // |q|^2 (fp32) — euclidean + score modifiers
    double qn2x[MQ];   // |q|^2 (fp64) — euclidean guard
    double ek[MQ];     // exact scan-domain key of the k-th document (diagnostics / next threshold)
    int status[MQ];
    int cnt[MQ];       // rows appended by the collect pass
    int ndocs[MQ];     // distinct documents the last exact selection saw
    int n_need;        // queries flagged for a(nother) collect pass
    int rounds;        // collect passes run by the no-sync path that still left queries flagged (sticky)
};

struct ScanParams {
    int n_rows;
    int dim;
    int num_tiles;
    int num_stages;
    int nq;
    int metric;
    const int32_t* doc_of_row;  // used when HAS_DOCS
    const float* row_bias;      // used when HAS_BIAS: key = 2 * dot - row_bias[row]  (euclidean: |row|^2)
    const float2* mod;          // used when HAS_MOD: per-document (mult, add); key = mult * closeness + add
    const uint32_t* filter;     // optional document bitset (HAS_DOCS instantiations): bit d clear = document d is excluded
    int64_t filter_docs;        // documents covered by `filter`; documents beyond are excluded
    QState* qs;
    // SELECT mode
    float* out_score;           // [grid][MQ][KP]
    int32_t* out_row;
    int32_t* out_doc;
    // COLLECT mode
    int32_t* cbuf;              // [MQ][ccap] rows
    int ccap;
};

__host__ __device__ inline size_t scan_smem_bytes(int dim, int stages, bool has_mod = false) {
    return (size_t)(dim / BLOCK_K) * QCHUNK_BYTES + (size_t)stages * STAGE_BYTES + 2 * 4 * TILE_N * sizeof(int32_t) +
           (has_mod ? 4 * TILE_N * sizeof(float2) : 0) + (2 * 16 + 2 * ACC_STAGES + 2) * sizeof(uint64_t) +
           1024 /* alignment slack */;
}

// fp32 closeness of the scan key (dot product, or 2 dot - |row|^2 for euclidean); the exact fp64 form is
// closeness_from_dot() below.  Only evaluated when score modifiers make the ranking non-monotone in the dot product.
__device__ __forceinline__ float closeness_approx(float v, int metric, float qn2) {
    switch (metric) {
        case B200_METRIC_EUCLIDEAN:
            return __fdividef(1.0f, 1.0f + sqrtf(fmaxf(qn2 - v, 0.0f)));
        case B200_METRIC_PRENORMALIZED_ANGULAR:
            return __fdividef(1.0f, 2.0f - v);
        case B200_METRIC_ANGULAR:
            return __fdividef(1.0f, 1.0f + acosf(fminf(1.0f, fmaxf(-1.0f, v))));
        default:
            return v;
    }
}

__device__ __forceinline__ float pick32(const uint32_t (&v)[32], int j) {
    uint32_t r = 0;
#pragma unroll
    for (int t = 0; t < 32; ++t)
        if (t == j) r = v[t];
    return __uint_as_float(r);
}

// Sorted (key desc, arrival order) list insert.  With HAS_DOCS a row whose document is already listed is
//   dropped    when its key is more than tol below the listed one (provably not the document's best chunk),
//   replaces   the listed entry when it is more than tol above it (the listed one is provably not the best),
//   kept too   otherwise (a near-tie: the exact pass decides which chunk represents the document).
template <bool HAS_DOCS>
__device__ __forceinline__ void list_insert(float (&ls)[KP], int (&lr)[KP], int (&ld)[KP], float s, int row, int doc,
                                            float tol) {
    if (HAS_DOCS) {
        int pos = -1;
#pragma unroll
        for (int i = KP - 1; i >= 0; --i)
            if (lr[i] >= 0 && ld[i] == doc) pos = i;   // first (= best) listed entry of this document
        if (pos >= 0) {
            float old = 0.f;
#pragma unroll
            for (int i = 0; i < KP; ++i)
                if (i == pos) old = ls[i];
            if (s < old - tol) return;
            if (s > old + tol) {
#pragma unroll
                for (int i = 0; i < KP - 1; ++i)
                    if (i >= pos) {
                        ls[i] = ls[i + 1];
                        lr[i] = lr[i + 1];
                        ld[i] = ld[i + 1];
                    }
                ls[KP - 1] = -INFINITY;
                lr[KP - 1] = -1;
                ld[KP - 1] = -1;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < KP; ++i) {
        if (s > ls[i]) {
            float ts = ls[i];
            int tr = lr[i], td = ld[i];
            ls[i] = s;
            lr[i] = row;
            ld[i] = doc;
            s = ts;
            row = tr;
            doc = td;
        }
    }
}

template <bool HAS_DOCS, bool HAS_BIAS, bool HAS_MOD, bool COLLECT>
__global__ void __launch_bounds__(THREADS, 1)
scan_kernel(const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_q, ScanParams p) {
    if (COLLECT) {
        // the fallback pass is always enqueued by the asynchronous entry point; nothing flagged -> nothing to do
        if (*reinterpret_cast<volatile int*>(&p.qs->n_need) == 0) return;
    }
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kblocks = p.dim / BLOCK_K;
    const int S = p.num_stages;
    uint8_t* smem_q = smem;
    uint8_t* smem_c = smem_q + (size_t)kblocks * QCHUNK_BYTES;
    int32_t* smem_docs = reinterpret_cast<int32_t*>(smem_c + (size_t)S * STAGE_BYTES);
    float* smem_bias = reinterpret_cast<float*>(smem_docs + 4 * TILE_N);
    float2* smem_mod = reinterpret_cast<float2*>(smem_bias + 4 * TILE_N);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_mod + (HAS_MOD ? 4 * TILE_N : 0));
    uint64_t* empty = full + 16;
    uint64_t* tfull = empty + 16;
    uint64_t* tempty = tfull + ACC_STAGES;
    uint64_t* qfull = tempty + ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qfull + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_c);
        ptx::prefetch_tmap(&tmap_q);
        for (int i = 0; i < S; ++i) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < ACC_STAGES; ++i) {
            ptx::mbar_init(&tfull[i], 1);
            ptx::mbar_init(&tempty[i], 4);
        }
        ptx::mbar_init(qfull, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            ptx::mbar_arrive_expect_tx(qfull, kblocks * QCHUNK_BYTES);
            for (int kb = 0; kb < kblocks; ++kb)
                ptx::tma_load_2d(smem_q + (size_t)kb * QCHUNK_BYTES, &tmap_q, qfull, kb * BLOCK_K, 0, ptx::kEvictLast);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < kblocks; ++kb) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    ptx::mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
                    ptx::tma_load_2d(smem_c + (size_t)stage * STAGE_BYTES, &tmap_c, &full[stage], kb * BLOCK_K,
                                     tile * TILE_N, ptx::kEvictFirst);
                    if (++stage == S) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        ptx::mbar_wait(qfull, 0);
        ptx::tc_fence_after();
        constexpr uint32_t idesc = ptx::make_idesc_f16(0 /*fp16*/, MQ, TILE_N);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
            ptx::tc_fence_after();
            for (int kb = 0; kb < kblocks; ++kb) {
                ptx::mbar_wait(&full[stage], phase);
                ptx::tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_base = ptx::smem_u32(smem_q + (size_t)kb * QCHUNK_BYTES);
                    const uint32_t b_base = ptx::smem_u32(smem_c + (size_t)stage * STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        ptx::umma_f16(tmem_base + acc * TILE_N, ptx::make_desc_k_sw128(a_base + k * UMMA_K * 2),
                                      ptx::make_desc_k_sw128(b_base + k * UMMA_K * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(&empty[stage]);
                    if (kb == kblocks - 1) ptx::umma_commit(&tfull[acc]);
                }
                __syncwarp();
                if (++stage == S) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (++acc == ACC_STAGES) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: per-query running top-KP / collect
        const int sp = warp & 3;  // TMEM sub-partition this warp may read
        const int q = sp * 16 + lane;  // M = 64 accumulator: row m lives in lane (m % 16) of sub-partition m / 16
        const bool active = lane < 16 && q < p.nq;
        int32_t* my_docs = smem_docs + (warp - 2) * TILE_N;
        float* my_bias = smem_bias + (warp - 2) * TILE_N;
        float2* my_mod = smem_mod + (HAS_MOD ? (warp - 2) * TILE_N : 0);
        const float my_qn2 = (HAS_MOD && active) ? p.qs->qn2[q] : 0.f;
        const float tol = (HAS_DOCS && active && !COLLECT) ? p.qs->tol[q] : 0.f;
        float ls[KP];
        int lr[KP], ld[KP];
#pragma unroll
        for (int i = 0; i < KP; ++i) {
            ls[i] = -INFINITY;
            lr[i] = -1;
            ld[i] = -1;
        }
        // SELECT: a row qualifies when key > thr (the list's tail).  COLLECT: when key >= thr (= L[q]; +inf when
        // the query is not flagged).
        float thr = -INFINITY;
        if (COLLECT) thr = active ? p.qs->L[q] : INFINITY;
        const uint32_t* filt = HAS_DOCS ? p.filter : nullptr;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int row0 = tile * TILE_N;
            if (HAS_DOCS || HAS_BIAS || HAS_MOD) {
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    int r = row0 + t * 32 + lane;
                    int d = r;
                    if (HAS_DOCS) {
                        d = r < p.n_rows ? __ldg(p.doc_of_row + r) : -1;
                        if (filt != nullptr && d >= 0) {
                            const bool keep = d < p.filter_docs && ((__ldg(filt + (d >> 5)) >> (d & 31)) & 1u);
                            if (!keep) d = -1;   // a filtered-out document looks like a deleted row
                        }
                        my_docs[t * 32 + lane] = d;
                    }
                    if (HAS_BIAS) my_bias[t * 32 + lane] = r < p.n_rows ? __ldg(p.row_bias + r) : 0.f;
                    if (HAS_MOD) my_mod[t * 32 + lane] = (r < p.n_rows && d >= 0) ? __ldg(p.mod + d) : make_float2(0.f, 0.f);
                }
                __syncwarp();
            }
            ptx::mbar_wait(&tfull[acc], acc_phase);
            ptx::tc_fence_after();
            const int valid = min(TILE_N, p.n_rows - row0);
#pragma unroll 1
            for (int c = 0; c < TILE_N / 32; ++c) {
                uint32_t v[32];
                ptx::tmem_ld_32x32b_x32(tmem_base + (uint32_t(sp * 32) << 16) + acc * TILE_N + c * 32, v);
                ptx::tmem_ld_wait();
                if (active) {
                    if (HAS_BIAS) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            v[j] = __float_as_uint(fmaf(2.0f, __uint_as_float(v[j]), -my_bias[c * 32 + j]));
                    }
                    if (HAS_MOD) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float2 ma = my_mod[c * 32 + j];
                            v[j] = __float_as_uint(fmaf(ma.x, closeness_approx(__uint_as_float(v[j]), p.metric, my_qn2), ma.y));
                        }
                    }
                    uint32_t mask = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float f = __uint_as_float(v[j]);
                        mask |= (COLLECT ? (f >= thr) : (f > thr)) ? (1u << j) : 0u;
                    }
                    const int nvalid = valid - c * 32;
                    if (nvalid < 32) mask &= nvalid <= 0 ? 0u : ((1u << nvalid) - 1u);
                    while (mask) {
                        const int j = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const float s = pick32(v, j);
                        const int row = row0 + c * 32 + j;
                        if (!COLLECT && !(s > thr)) continue;
                        int doc = row;
                        if (HAS_DOCS) {
                            doc = my_docs[c * 32 + j];
                            if (doc < 0) continue;  // tombstoned or filtered-out row
                        }
                        if (COLLECT) {
                            const int slot = atomicAdd(&p.qs->cnt[q], 1);
                            if (slot < p.ccap) p.cbuf[(size_t)q * p.ccap + slot] = row;
                        } else {
                            list_insert<HAS_DOCS>(ls, lr, ld, s, row, doc, tol);
                            thr = ls[KP - 1];
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
            if (++acc == ACC_STAGES) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        if (!COLLECT && lane < 16) {
            const size_t base = ((size_t)blockIdx.x * MQ + q) * KP;
#pragma unroll
            for (int i = 0; i < KP; ++i) {
                p.out_score[base + i] = ls[i];
                p.out_row[base + i] = lr[i];
                p.out_doc[base + i] = ld[i];
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// Exact selection shared by merge_kernel (candidates from the per-CTA lists) and finalize_kernel (rows from the
// collect pass).
struct ExactParams {
    int nq;
    int k;
    int dim;
    int metric;
    int doc_offset;        // added to every returned document number (global numbering of a row-sharded corpus)
    const __half* qh;      // [MQ, dim] fp16 queries as scanned
    const __half* corpus;  // [n_rows, dim]
    const int32_t* doc_of_row;
    const double2* mod64;  // optional score modifiers: per-document (mult, add) in fp64; key = mult * closeness + add
    QState* qs;
    int32_t* out_doc;      // [nq, k]
    int32_t* out_row;
    double* out_score;
};

// `val` is the exact ordering key of the re-score pass: the dot product, or minus the squared distance (euclidean)
__device__ __forceinline__ double closeness_from_dot(double dot, int metric) {
    switch (metric) {
        case B200_METRIC_EUCLIDEAN:
            return 1.0 / (1.0 + sqrt(fmax(-dot, 0.0)));
        case B200_METRIC_PRENORMALIZED_ANGULAR:
            return 1.0 / (1.0 + (1.0 - dot));
        case B200_METRIC_ANGULAR: {
            double c = fmin(1.0, fmax(-1.0, dot));
            return 1.0 / (1.0 + acos(c));
        }
        default:
            return dot;
    }
}

// Exact fp64 dot product (or minus squared distance) of query `qv` and corpus row `cv`, computed by one warp.
// Fixed order, restated by oracle/score_oracle.c: lane l accumulates elements i = 256 j + 8 l + t (j ascending, then
// t = 0..7), lanes are combined by the xor butterfly 16, 8, 4, 2, 1.  Every product of two fp16 values is exact in
// fp64, so fused and unfused multiply-add give the same bits; the euclidean form uses explicit unfused ops.
__device__ __forceinline__ double warp_exact_dot(const __half* __restrict__ qv, const __half* __restrict__ cv, int dim,
                                                 int metric, int lane) {
    double part = 0.0;
    for (int base = 8 * lane; base < dim; base += 256) {
        const uint4 qa = *reinterpret_cast<const uint4*>(qv + base);
        const uint4 ca = __ldg(reinterpret_cast<const uint4*>(cv + base));
        const __half2* qh2 = reinterpret_cast<const __half2*>(&qa);
        const __half2* ch2 = reinterpret_cast<const __half2*>(&ca);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 qf = __half22float2(qh2[t]);
            const float2 cf = __half22float2(ch2[t]);
            if (metric == B200_METRIC_EUCLIDEAN) {
                const double d0 = (double)qf.x - (double)cf.x;   // exact in fp64
                const double d1 = (double)qf.y - (double)cf.y;
                part = __dsub_rn(part, __dmul_rn(d0, d0));
                part = __dsub_rn(part, __dmul_rn(d1, d1));
            } else {
                part = __dadd_rn(part, __dmul_rn((double)qf.x, (double)cf.x));
                part = __dadd_rn(part, __dmul_rn((double)qf.y, (double)cf.y));
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part = __dadd_rn(part, __shfl_xor_sync(0xffffffffu, part, o));
    return part;
}

// Block-wide: re-score entries [0, n) (rows in s_row) exactly, keep each document's best chunk (dot desc, row asc),
// rank the documents by (key desc, doc asc) and write the best k.  Returns (through smem) the number of distinct
// documents and the exact scan-domain key of the k-th one.
//   scan-domain key = what the scan's approximate key approximates: dot | dot + |q|^2 (euclidean: 2 q.e - |e|^2) |
//   the modified score.
template <int NTHREADS>
__device__ void exact_select(const ExactParams& p, int q, int n, const int32_t* s_row, int32_t* s_doc, double* x_dot,
                             double* x_key, uint8_t* x_rep, int* s_ndocs, double* s_ek) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const __half* qv = p.qh + (size_t)q * p.dim;
    for (int c = warp; c < n; c += NTHREADS / 32) {
        const int row = s_row[c];
        const double tot = warp_exact_dot(qv, p.corpus + (size_t)row * p.dim, p.dim, p.metric, lane);
        if (lane == 0) {
            x_dot[c] = tot;
            double key = tot;
            if (p.mod64) {   // the ordering key becomes the modified score (separate multiply and add, no fma)
                const double2 ma = p.mod64[s_doc[c]];
                key = __dadd_rn(__dmul_rn(ma.x, closeness_from_dot(tot, p.metric)), ma.y);
            }
            x_key[c] = key;
        }
    }
    if (threadIdx.x == 0) {
        *s_ndocs = 0;
        *s_ek = -INFINITY;
    }
    __syncthreads();
    // best chunk per document
    for (int i = threadIdx.x; i < n; i += NTHREADS) {
        const int d = s_doc[i], r = s_row[i];
        const double v = x_dot[i];
        bool rep = true;
        for (int j = 0; j < n; ++j)
            rep &= !(s_doc[j] == d && (x_dot[j] > v || (x_dot[j] == v && s_row[j] < r)));
        x_rep[i] = rep ? 1 : 0;
        if (rep) atomicAdd(s_ndocs, 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += NTHREADS) {
        if (!x_rep[i]) continue;
        const int d = s_doc[i];
        const double v = x_key[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += (x_rep[j] && (x_key[j] > v || (x_key[j] == v && s_doc[j] < d))) ? 1 : 0;
        if (rank < p.k) {
            const size_t o = (size_t)q * p.k + rank;
            p.out_doc[o] = d + p.doc_offset;
            p.out_row[o] = s_row[i];
            p.out_score[o] = p.mod64 ? v : closeness_from_dot(v, p.metric);
            if (rank == p.k - 1)
                *s_ek = p.mod64 ? v : (p.metric == B200_METRIC_EUCLIDEAN ? v + p.qs->qn2x[q] : v);
        }
    }
    __syncthreads();
    const int nd = *s_ndocs;
    for (int i = nd + threadIdx.x; i < p.k; i += NTHREADS) {
        const size_t o = (size_t)q * p.k + i;
        p.out_doc[o] = -1;
        p.out_row[o] = -1;
        p.out_score[o] = -INFINITY;
    }
}

// monotone float <-> uint32 map (for shared-memory atomicMax / atomicMin on scores of either sign)
__device__ __forceinline__ uint32_t flt_key(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_flt(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct MergeParams {
    int num_lists;  // scan grid size
    int sort_n;     // power of two >= num_lists * KP
    int l_only;     // k > K_MERGE_MAX: only choose the collect threshold L (the collect pass answers)
    const float* in_score;
    const int32_t* in_row;
    const int32_t* in_doc;
    ExactParams ex;
};

__device__ __forceinline__ bool approx_before(float sa, int ra, float sb, int rb) {
    return sa > sb || (sa == sb && ra < rb);
}

constexpr int MERGE_THREADS = 256;

__global__ void __launch_bounds__(MERGE_THREADS) merge_kernel(MergeParams p) {
    extern __shared__ uint8_t msmem[];
    float* s_score = reinterpret_cast<float*>(msmem);
    int32_t* s_row = reinterpret_cast<int32_t*>(s_score + p.sort_n);
    int32_t* s_doc = s_row + p.sort_n;
    __shared__ double x_dot[M_CAP], x_key[M_CAP];
    __shared__ uint8_t x_rep[M_CAP];
    __shared__ float s_head[MAX_LISTS];
    __shared__ float s_thr;
    __shared__ uint32_t s_tail_max, s_below_max, s_union_min;
    __shared__ int s_count, s_ndocs;
    __shared__ double s_ek;

    const int q = blockIdx.x;
    const int k = p.ex.k;
    QState* qs = p.ex.qs;
    const int total = p.num_lists * KP;
    // candidates wanted from the lists: enough that the k-th exact key clears the best unselected entry
    const int need = p.l_only ? INT_MAX : (k <= K_SMALL ? KP : k + k / 4 + 16);
    const int e_sel = p.l_only ? 0 : (need + KP - 1) / KP - 1;   // <= KP - 1 by the static_asserts
    if (threadIdx.x == 0) {
        s_thr = -INFINITY;
        s_tail_max = flt_key(-INFINITY);
        s_below_max = flt_key(-INFINITY);
        s_union_min = flt_key(INFINITY);
        s_count = 0;
    }
    __syncthreads();
    // Every list is sorted, so the KP-th largest of the lists' e_sel-th entries is a lower bound of the global
    // KP*(e_sel+1)-th best key: only candidates >= that bound are wanted.  Shrinks the sort from ~2.4k keys to ~need.
    for (int l = threadIdx.x; l < p.num_lists; l += blockDim.x) {
        const size_t src = ((size_t)l * MQ + q) * KP;
        s_head[l] = p.in_row[src + e_sel] >= 0 ? p.in_score[src + e_sel] : -INFINITY;
        // a FULL list may have rejected rows: its tail bounds their keys
        if (p.in_row[src + KP - 1] >= 0) atomicMax(&s_tail_max, flt_key(p.in_score[src + KP - 1]));
    }
    __syncthreads();
    if (!p.l_only && p.num_lists >= KP) {
        for (int l = threadIdx.x; l < p.num_lists; l += blockDim.x) {
            const float h = s_head[l];
            int rank = 0;
            for (int j = 0; j < p.num_lists; ++j) rank += (s_head[j] > h) || (s_head[j] == h && j < l);
            if (rank == KP - 1) s_thr = h;
        }
    }
    __syncthreads();
    const float thr = s_thr;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int list = i / KP, e = i % KP;
        const size_t src = ((size_t)list * MQ + q) * KP + e;
        const int rr = p.in_row[src];
        if (rr < 0) continue;
        const float sc = p.in_score[src];
        atomicMin(&s_union_min, flt_key(sc));
        if (sc >= thr) {
            const int slot = atomicAdd(&s_count, 1);
            s_score[slot] = sc;
            s_row[slot] = rr;
            s_doc[slot] = p.in_doc[src];
        } else {
            atomicMax(&s_below_max, flt_key(sc));
        }
    }
    __syncthreads();
    const int count = s_count;
    int n2 = 32;
    while (n2 < count) n2 <<= 1;
    for (int i = count + threadIdx.x; i < n2; i += blockDim.x) {
        s_score[i] = -INFINITY;
        s_row[i] = INT_MAX;
        s_doc[i] = -1;
    }
    __syncthreads();
    // bitonic sort, order: (key desc, row asc) — a total order, so the result does not depend on slot order
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < n2 / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const float sa = s_score[lo], sb = s_score[hi];
                const int ra = s_row[lo], rb = s_row[hi];
                const bool wrong = up ? approx_before(sb, rb, sa, ra) : approx_before(sa, ra, sb, rb);
                if (wrong) {
                    s_score[lo] = sb;
                    s_score[hi] = sa;
                    s_row[lo] = rb;
                    s_row[hi] = ra;
                    const int da = s_doc[lo];
                    s_doc[lo] = s_doc[hi];
                    s_doc[hi] = da;
                }
            }
            __syncthreads();
        }
    }
    const float eps = qs->eps[q];
    if (p.l_only) {
        // k beyond the merge kernel's reach: collect every row at least as good as the (k + slack)-th list entry
        if (threadIdx.x == 0) {
            const int want = k + k / 8 + 16;
            float L = -INFINITY;   // fewer entries than wanted and no list is full: the lists ARE the corpus
            if (count >= want) L = s_score[want - 1];
            else if (key_flt(s_tail_max) > -INFINITY) L = key_flt(s_union_min);
            qs->L[q] = L;
            qs->cnt[q] = 0;
            qs->status[q] = Q_NEED;
            atomicAdd(&qs->n_need, 1);
        }
        return;
    }
    const int mcap = k <= K_SMALL ? M_SMALL : M_CAP;
    const int M = min(count, mcap);
    exact_select<MERGE_THREADS>(p.ex, q, M, s_row, s_doc, x_dot, x_key, x_rep, &s_ndocs, &s_ek);
    __syncthreads();
    if (threadIdx.x == 0) {
        // tau: upper bound of the approximate key of every live row that was not re-scored
        const float tau_m = count > M ? s_score[M] : key_flt(s_below_max);
        const float tau = fmaxf(key_flt(s_tail_max), tau_m);
        const int nd = s_ndocs;
        const bool resolved = nd >= k ? (s_ek > (double)tau + (double)eps) : (tau == -INFINITY);
        qs->ndocs[q] = nd;
        qs->ek[q] = s_ek;
        qs->cnt[q] = 0;
        if (resolved) {
            qs->status[q] = Q_RESOLVED;
            qs->L[q] = INFINITY;
        } else {
            // every row whose exact key can reach the k-th one has approximate key >= ek - eps; with fewer than k
            // documents in hand start from the weakest list entry
            qs->L[q] = nd >= k ? __double2float_rd(s_ek - (double)eps) : key_flt(s_union_min);
            qs->status[q] = Q_NEED;
            atomicAdd(&qs->n_need, 1);
        }
    }
}

// One CTA per flagged query: exact selection over the rows appended by the collect pass.
constexpr int FIN_THREADS = 512;
struct FinalizeParams {
    const int32_t* cbuf;
    int ccap;
    int64_t live_bound;   // n_rows: a collect pass with L = -inf saw every live row
    ExactParams ex;
};

__host__ __device__ inline size_t finalize_smem_bytes(int n) { return (size_t)n * (4 + 4 + 8 + 8 + 1) + 64; }

__global__ void __launch_bounds__(FIN_THREADS) finalize_kernel(FinalizeParams p) {
    extern __shared__ __align__(16) uint8_t fsmem[];
    QState* qs = p.ex.qs;
    const int q = blockIdx.x;
    if (qs->status[q] != Q_NEED) return;
    const int cnt = qs->cnt[q];
    const int k = p.ex.k;
    __shared__ int s_ndocs;
    __shared__ double s_ek;
    if (cnt > p.ccap || cnt > FIN_CAP) {   // overflow: the host grows the buffer / finalizes on the host
        if (threadIdx.x == 0) atomicAdd(&qs->n_need, 1);
        return;
    }
    double* x_dot = reinterpret_cast<double*>(fsmem);
    double* x_key = x_dot + cnt;
    int32_t* s_row = reinterpret_cast<int32_t*>(x_key + cnt);
    int32_t* s_doc = s_row + cnt;
    uint8_t* x_rep = reinterpret_cast<uint8_t*>(s_doc + cnt);
    for (int i = threadIdx.x; i < cnt; i += FIN_THREADS) {
        const int r = p.cbuf[(size_t)q * p.ccap + i];
        s_row[i] = r;
        s_doc[i] = p.ex.doc_of_row[r];
    }
    __syncthreads();
    exact_select<FIN_THREADS>(p.ex, q, cnt, s_row, s_doc, x_dot, x_key, x_rep, &s_ndocs, &s_ek);
    __syncthreads();
    if (threadIdx.x == 0) {
        const float L = qs->L[q];
        const float eps = qs->eps[q];
        const int nd = s_ndocs;
        // rows not collected have approximate key < L, i.e. exact key < L + eps
        const bool resolved = (L == -INFINITY) || (nd >= k && s_ek >= (double)L + (double)eps);
        qs->ndocs[q] = nd;
        qs->ek[q] = s_ek;
        qs->cnt[q] = 0;
        if (resolved) {
            qs->status[q] = Q_RESOLVED;
            qs->L[q] = INFINITY;
        } else {
            float nl;
            if (nd >= k) nl = __double2float_rd(s_ek - (double)eps);
            else nl = L - fmaxf(8.0f * eps, 0.05f * fabsf(L));   // too few documents above L: look deeper
            if (!(nl < L)) nl = -INFINITY;
            qs->L[q] = nl;
            atomicAdd(&qs->n_need, 1);
        }
    }
}

// Host-finalize support: exact dot and key of every collected row of one query (warp per row).
__global__ void exact_keys_kernel(ExactParams p, int q, const int32_t* rows, int n, double* out_dot, double* out_key,
                                  int32_t* out_doc) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const int row = rows[i];
    const double tot = warp_exact_dot(p.qh + (size_t)q * p.dim, p.corpus + (size_t)row * p.dim, p.dim, p.metric, lane);
    if (lane == 0) {
        const int d = p.doc_of_row[row];
        double key = tot;
        if (p.mod64) {
            const double2 ma = p.mod64[d];
            key = __dadd_rn(__dmul_rn(ma.x, closeness_from_dot(tot, p.metric)), ma.y);
        }
        out_dot[i] = tot;
        out_key[i] = key;
        out_doc[i] = d;
    }
}

// Per-query error bound of the approximate scan key (warp per query).  fp16 x fp16 products are exact in fp32; the
// tensor core adds them (16 per instruction, dim / 16 instructions) with at most one truncation per addend, so
//   |approx dot - exact dot| <= dim * 2^-23 * sum|q_i e_i| <= dim * 2^-23 * |q| |e|.
// The bound used is twice that (c = dim * 2^-22) with |e| <= sqrt(max_n2), the largest stored row norm.
__global__ void query_prep_kernel(const __half* __restrict__ qh, int dim, int metric, int has_mod,
                                  const float* __restrict__ max_n2, const double* __restrict__ mod_max, QState* qs) {
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= MQ) return;
    double n2 = 0.0;
    for (int i = lane; i < dim; i += 32) {
        const double f = (double)__half2float(qh[(size_t)q * dim + i]);
        n2 += f * f;
    }
    for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    if (lane != 0) return;
    const float qn = sqrtf((float)n2) * 1.001f;
    const float R = sqrtf(*max_n2) * 1.001f;
    const float c = (float)dim * 2.384185791015625e-07f;   // dim * 2^-22
    float eps = c * qn * R;                                // dot-product domain
    if (metric == B200_METRIC_EUCLIDEAN)                   // key = fma(2, dot, -n2_fp32[row])
        eps = 2.0f * eps + 0.5f * c * R * R + 4.76837158203125e-07f * (2.0f * qn * R + R * R);
    if (has_mod) {
        // key = fma(mult32, closeness32(v), add32); closeness error from the error eps of v:
        float ec;
        switch (metric) {
            case B200_METRIC_PRENORMALIZED_ANGULAR: {   // 1/(2 - v): Lipschitz 1/(2 - vmax)^2
                const float room = 2.0f - qn * R - eps;
                ec = room > 0.05f ? eps / (room * room) : INFINITY;
                break;
            }
            case B200_METRIC_ANGULAR:                     // |acos a - acos b| <= 2 sqrt|a - b|, |d closeness / d theta| <= 1
                ec = 2.0f * sqrtf(eps);
                break;
            case B200_METRIC_EUCLIDEAN:                   // |sqrt a - sqrt b| <= sqrt|a - b|
                ec = sqrtf(eps + 4.76837158203125e-07f * qn * qn);
                break;
            default:
                ec = eps;
        }
        const float mmax = (float)mod_max[0] * 1.001f, amax = (float)mod_max[1] * 1.001f;
        eps = mmax * (ec + 1e-6f) + (mmax + amax) * 4.76837158203125e-07f;
    }
    eps += 1e-30f;
    qs->eps[q] = eps;
    qs->tol[q] = 2.0f * eps;
    qs->L[q] = INFINITY;
    qs->qn2[q] = (float)n2;
    qs->qn2x[q] = n2;
    qs->status[q] = Q_RESOLVED;
    qs->cnt[q] = 0;
    qs->ndocs[q] = 0;
    qs->ek[q] = 0.0;
    if (q == 0) qs->n_need = 0;
}

__global__ void reset_need_kernel(QState* qs) { qs->n_need = 0; }
__global__ void note_unresolved_kernel(QState* qs) {
    if (qs->n_need > 0) qs->rounds += 1;
}

// ------------------------------------------------------------------------------------------------
// fp32 -> fp16 row conversion (optionally L2-normalising first, for the angular metric).
// `flags`: bit 0 is set when a value is not finite or does not fit fp16 (|x| > 65504 after normalisation);
// `max_n2` accumulates the largest squared norm of the STORED rows (the error bound of the scan needs it).
__global__ void convert_rows_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t rows, int dim,
                                    int64_t dst_rows_total, int normalize, float* __restrict__ n2_out,
                                    float* __restrict__ max_n2, int* __restrict__ flags) {
    // one warp per row; rows in [rows, dst_rows_total) are zero-filled (query padding)
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= dst_rows_total) return;
    __half* d = dst + row * dim;
    if (row >= rows) {
        for (int i = lane; i < dim; i += 32) d[i] = __float2half_rn(0.f);
        return;
    }
    const float* s = src + row * dim;
    float scale = 1.f;
    if (normalize) {
        float ss = 0.f;
        for (int i = lane; i < dim; i += 32) ss = __fmaf_rn(s[i], s[i], ss);
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        scale = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
    }
    float n2 = 0.f;
    bool bad = false;
    for (int i = lane; i < dim; i += 32) {
        const float x = s[i] * scale;
        bad |= !(fabsf(x) <= 65504.0f);   // also true for NaN
        const __half hv = __float2half_rn(x);
        d[i] = hv;
        const float f = __half2float(hv);
        n2 = fmaf(f, f, n2);
    }
    for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    if (__any_sync(0xffffffffu, bad) && lane == 0 && flags) atomicOr(flags, 1);
    if (lane == 0) {
        if (n2_out) n2_out[row] = n2;   // |row|^2 of the STORED (fp16-rounded) values: the euclidean scan's per-row term
        if (max_n2 && n2 == n2 && n2 < INFINITY) atomicMax(reinterpret_cast<int*>(max_n2), __float_as_int(n2));   // n2 >= 0
    }
}

__global__ void row_norms_kernel(const __half* __restrict__ rows, int64_t n, int dim, float* __restrict__ n2_out,
                                 float* __restrict__ max_n2) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    float n2 = 0.f;
    for (int i = lane; i < dim; i += 32) {
        const float f = __half2float(rows[row * dim + i]);
        n2 = fmaf(f, f, n2);
    }
    for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    if (lane == 0) {
        if (n2_out) n2_out[row] = n2;
        if (n2 == n2 && n2 < INFINITY) atomicMax(reinterpret_cast<int*>(max_n2), __float_as_int(n2));
    }
}

__global__ void iota_kernel(int32_t* dst, int64_t n, int32_t start) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = start + (int32_t)i;
}

__global__ void tombstone_kernel(int32_t* doc_of_row, int64_t n, int32_t doc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && doc_of_row[i] == doc) doc_of_row[i] = -1;
}

__global__ void tombstone_rows_kernel(int32_t* doc_of_row, const int32_t* rows, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) doc_of_row[rows[i]] = -1;
}

// compaction: copy live rows to their new positions (one warp per OLD row; new_of_old[row] < 0 = dead)
__global__ void compact_rows_kernel(const __half* __restrict__ src, __half* __restrict__ dst,
                                    const int32_t* __restrict__ src_doc, int32_t* __restrict__ dst_doc,
                                    const float* __restrict__ src_n2, float* __restrict__ dst_n2,
                                    const int32_t* __restrict__ new_of_old, int64_t n, int dim) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const int32_t to = new_of_old[row];
    if (to < 0) return;
    const uint4* s = reinterpret_cast<const uint4*>(src + row * dim);
    uint4* d = reinterpret_cast<uint4*>(dst + (int64_t)to * dim);
    for (int i = lane; i < dim / 8; i += 32) d[i] = s[i];
    if (lane == 0) {
        dst_doc[to] = src_doc[row];
        if (src_n2) dst_n2[to] = src_n2[row];
    }
}

// Score modifiers (reference: the rank-profile function `modify`, unstructured_vespa_schema.py:266-271):
//   mult = count(mult_w * attr) == 0 ? 1 : prod(mult_w * attr);   add = sum(add_w * attr)
// over the attribute cells a document HAS (NaN = missing cell of the sparse tensor<double>(p{})), in the order the
// caller lists the columns.  Written in fp64 for the exact merge and in fp32 for the scan.
constexpr int MAX_MOD_TERMS = 16;
struct ModifierParams {
    int n_docs;
    int attr_cap;
    int n_mult, n_add;
    const double* mult_col[MAX_MOD_TERMS];
    const double* add_col[MAX_MOD_TERMS];
    double mult_w[MAX_MOD_TERMS];
    double add_w[MAX_MOD_TERMS];
    double2* out64;
    float2* out32;
    int* negative_flag;
    double* mod_max;   // [2]: max |mult|, max |add| over the documents (error bound of the modified scan key)
};

__global__ void modifier_kernel(ModifierParams p) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= p.n_docs) return;
    double m = 1.0, a = 0.0;
    int cnt = 0;
    for (int i = 0; i < p.n_mult; ++i) {
        const double v = (p.mult_col[i] && d < p.attr_cap) ? p.mult_col[i][d] : NAN;
        if (v == v) {
            m = __dmul_rn(m, __dmul_rn(p.mult_w[i], v));
            ++cnt;
        }
    }
    if (cnt == 0) m = 1.0;
    for (int i = 0; i < p.n_add; ++i) {
        const double v = (p.add_col[i] && d < p.attr_cap) ? p.add_col[i][d] : NAN;
        if (v == v) a = __dadd_rn(a, __dmul_rn(p.add_w[i], v));
    }
    p.out64[d] = make_double2(m, a);
    p.out32[d] = make_float2((float)m, (float)a);
    if (m < 0.0) atomicOr(p.negative_flag, 1);
    // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long*>(p.mod_max), (unsigned long long)__double_as_longlong(fabs(m)));
    atomicMax(reinterpret_cast<unsigned long long*>(p.mod_max + 1), (unsigned long long)__double_as_longlong(fabs(a)));
}

__global__ void fill_nan_kernel(double* dst, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = NAN;
}

__global__ void scatter_attr_kernel(double* col, const int32_t* docs, const double* vals, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) col[docs[i]] = vals ? vals[i] : NAN;
}

// (column, document, value) triples in one launch; cols_table[c] = device column pointer
__global__ void scatter_attr_multi_kernel(double* const* cols_table, const int32_t* cols, const int32_t* docs,
                                          const double* vals, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cols_table[cols[i]][docs[i]] = vals[i];
}

// clear every attribute cell of the listed documents (document overwritten or deleted)
__global__ void clear_attr_kernel(double* const* cols_table, int n_cols, const int32_t* docs, int64_t n, int64_t attr_cap) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t d = docs[i];
    if (d >= attr_cap) return;
    for (int c = 0; c < n_cols; ++c)
        if (cols_table[c]) cols_table[c][d] = NAN;
}

// Merge of all-gathered per-shard lists on the device: one warp per query, candidates strided over the lanes,
// k rounds of warp arg-max under (score desc, doc asc).  Shard s's block: doc int32 [nq,k] | row int32 [nq,k] |
// score f64 [nq,k] packed back to back (the layout b200_index_search_device writes when given one buffer).
__device__ __forceinline__ void warp_merge_shards(const uint8_t* __restrict__ gathered, size_t shard_stride, int nshards,
                                                  int nq, int k, int q, int lane, int32_t* out_doc, int32_t* out_row,
                                                  double* out_score) {
    const int total = nshards * k;
    constexpr int PER = 8;  // up to 256 candidates per query
    double sc[PER];
    int dc[PER], rw[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = lane + 32 * i;
        sc[i] = -INFINITY;
        dc[i] = INT_MAX;
        rw[i] = -1;
        if (c < total) {
            const int s = c / k, e = c % k;
            const uint8_t* base = gathered + (size_t)s * shard_stride;
            const int d = reinterpret_cast<const int32_t*>(base)[(size_t)q * k + e];
            if (d >= 0) {
                dc[i] = d;
                rw[i] = reinterpret_cast<const int32_t*>(base + (size_t)nq * k * 4)[(size_t)q * k + e];
                sc[i] = reinterpret_cast<const double*>(base + (size_t)nq * k * 8)[(size_t)q * k + e];
            }
        }
    }
    for (int r = 0; r < k; ++r) {
        double bs = -INFINITY;
        int bd = INT_MAX, bi = -1;
#pragma unroll
        for (int i = 0; i < PER; ++i)
            if (sc[i] > bs || (sc[i] == bs && dc[i] < bd)) {
                bs = sc[i];
                bd = dc[i];
                bi = i;
            }
        int br = -1;
#pragma unroll
        for (int i = 0; i < PER; ++i)
            if (i == bi) br = rw[i];
        int owner = lane;
        for (int off = 16; off > 0; off >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, bs, off);
            const int od = __shfl_xor_sync(0xffffffffu, bd, off);
            const int orr = __shfl_xor_sync(0xffffffffu, br, off);
            const int oo = __shfl_xor_sync(0xffffffffu, owner, off);
            if (os > bs || (os == bs && od < bd)) {
                bs = os;
                bd = od;
                br = orr;
                owner = oo;
            }
        }
        if (lane == 0) {
            const size_t o = (size_t)q * k + r;
            const bool ok = bd != INT_MAX;
            out_doc[o] = ok ? bd : -1;
            out_row[o] = ok ? br : -1;
            out_score[o] = ok ? bs : -INFINITY;
        }
        if (lane == owner && bi >= 0) {   // retire the winner
#pragma unroll
            for (int i = 0; i < PER; ++i)
                if (i == bi) {
                    sc[i] = -INFINITY;
                    dc[i] = INT_MAX;
                }
        }
    }
}

__global__ void merge_shards_kernel(const uint8_t* __restrict__ gathered, int nshards, int nq, int k, int32_t* out_doc,
                                    int32_t* out_row, double* out_score) {
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    warp_merge_shards(gathered, (size_t)nq * k * 16, nshards, nq, k, q, lane, out_doc, out_row, out_score);
}

// ------------------------------------------------------------------------------------------------
// Fused exchange + merge over NVLink peer memory ("peer-stores into a symmetric buffer").
// Every rank owns a symmetric exchange buffer [2 parities][world][block] + flags [2][world]; rank r's kernel
//   1. copies its packed local result block into slot r of EVERY peer's buffer (plain stores over NVLink / local),
//   2. publishes it with a system-scope release store of the call's epoch into flag[r] on every peer,
//   3. spins (acquire loads on its OWN flags) until every rank's block of this epoch has landed,
//   4. merges the world * k candidates per query — the same warp_merge_shards as the all-gather path.
// One launch replaces ncclAllGather + merge_shards_kernel; parity = epoch & 1 double-buffers consecutive calls so a
// fast rank's next block never overwrites one a slow rank is still merging.
struct ExchangeParams {
    uint8_t* peer_buf[8];        // peer_buf[s] = base of rank s's exchange buffer mapped into this process
    unsigned long long* peer_flag[8];
    int rank, world;
    int nq, k;
    size_t block_bytes;          // nq * k * 16
    size_t slot_stride;          // bytes between slots (>= block_bytes, 16-byte aligned)
    unsigned long long epoch;
    const uint8_t* local_block;  // packed {doc | row | score} of this rank
    int32_t* out_doc;
    int32_t* out_row;
    double* out_score;
};

__global__ void __launch_bounds__(256) exchange_merge_kernel(ExchangeParams p) {
    const int parity = (int)(p.epoch & 1ull);
    const size_t par_off = (size_t)parity * p.world * p.slot_stride;
    const int n16 = (int)(p.block_bytes / 16);
    const uint4* src = reinterpret_cast<const uint4*>(p.local_block);
    // 1. push: thread t of the grid copies 16-byte words; peers interleaved so every NVLink port sees traffic at once
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int gthreads = gridDim.x * blockDim.x;
    for (int i = gtid; i < n16 * p.world; i += gthreads) {
        const int s = i % p.world, w = i / p.world;
        uint4* dst = reinterpret_cast<uint4*>(p.peer_buf[s] + par_off + (size_t)p.rank * p.slot_stride);
        dst[w] = src[w];
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish: one counter per (peer, source rank); every CTA adds 1, the block is complete at epoch * gridDim.x
    if (threadIdx.x < p.world) {
        unsigned long long* f = p.peer_flag[threadIdx.x] + (size_t)parity * p.world + p.rank;
        asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(f), "l"(1ull) : "memory");
    }
    // 3. wait for every source rank's block of this epoch
    if (threadIdx.x < p.world) {
        const unsigned long long* f = p.peer_flag[p.rank] + (size_t)parity * p.world + threadIdx.x;
        const unsigned long long want = ((p.epoch >> 1) + 1ull) * (unsigned long long)gridDim.x;
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
        } while (v < want);
    }
    __syncthreads();
    // 4. merge: one warp per query
    const uint8_t* mine = p.peer_buf[p.rank] + par_off;
    const int lane = threadIdx.x & 31;
    for (int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < p.nq; q += gridDim.x * (blockDim.x >> 5))
        warp_merge_shards(mine, p.slot_stride, p.world, p.nq, p.k, q, lane, p.out_doc, p.out_row, p.out_score);
}

__global__ void fill_empty_kernel(int32_t* out_doc, int32_t* out_row, double* out_score, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        out_doc[i] = -1;
        out_row[i] = -1;
        out_score[i] = -INFINITY;
    }
}

}  // namespace score
}  // namespace mb

// ====================================================================================================
using namespace mb;
using namespace mb::score;

struct b200_exchange {
    int device = 0;
    int rank = 0, world = 1;
    size_t slot_stride = 0;      // bytes per (parity, source rank) slot
    size_t buf_bytes = 0;
    uint8_t* local = nullptr;    // cudaMalloc'ed: [2][world][slot_stride] blocks, then [2][world] u64 flags
    uint8_t* peer[8] = {nullptr};   // peer[s] = rank s's buffer mapped here (peer[rank] == local)
    bool opened[8] = {false};
    unsigned long long epoch = 0;
};

struct b200_index {
    int device = 0;
    int dim = 0;
    int metric = 0;
    int sms = 0;
    int64_t capacity = 0;
    int64_t n_rows = 0;
    int64_t dead_rows = 0;  // tombstoned rows still occupying the matrix
    bool has_docs = false;  // false while doc_of_row[i] == i for every row (identity fast path)
    int32_t doc_offset = 0; // added to returned document numbers (shard -> global numbering)
    __half* corpus = nullptr;
    int32_t* doc_of_row = nullptr;
    float* row_n2 = nullptr;       // [capacity] squared norms (euclidean metric only)
    float* max_n2 = nullptr;       // [1] largest squared norm of a stored row
    int* d_flags = nullptr;        // [2]: bit 0 of [0] = non-finite / out-of-fp16-range input; [1] = negative multiplier
    // per-search workspaces
    __half* qh = nullptr;          // [MQ, dim]
    float* q_stage = nullptr;      // [MQ, dim] fp32 staging for host queries
    float* list_score = nullptr;   // [grid][MQ][KP]
    int32_t* list_row = nullptr;
    int32_t* list_doc = nullptr;
    int out_k = 0;                 // the resident output block holds [MQ, out_k]
    int32_t* o_doc = nullptr;
    int32_t* o_row = nullptr;
    double* o_score = nullptr;
    QState* qs = nullptr;          // device
    QState* h_qs = nullptr;        // pinned host mirror
    int32_t* cbuf = nullptr;       // [MQ][ccap] rows appended by the collect pass
    int ccap = 0;
    // score modifiers: per-document numeric attributes (one device column per attribute name, NaN = missing)
    std::vector<double*> attr_cols;
    double** d_cols_table = nullptr;  // device copy of attr_cols (B200_MAX_ATTRIBUTE_COLUMNS entries)
    bool cols_table_dirty = true;
    int64_t attr_cap = 0;          // documents each column can hold
    int64_t max_doc = -1;          // largest explicit document number seen by add()
    double2* mod64 = nullptr;      // [mod_cap] (mult, add) of the current modified search
    float2* mod32 = nullptr;
    int64_t mod_cap = 0;
    double* mod_max = nullptr;     // [2] max |mult|, max |add|
    bool mod_active = false;
    // document filter of the current search (device bitset over local document numbers)
    uint32_t* filter_bits = nullptr;
    int64_t filter_cap_words = 0;
    int64_t filter_docs = 0;
    uint64_t filter_tag = 0;       // identity of the bitset held in filter_bits (0 = none cached)
    bool filter_active = false;
    // statistics of the exactness machinery (b200_index_search_stats)
    int64_t stat_groups = 0, stat_flagged = 0, stat_collect_passes = 0, stat_host_finalize = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool timing_valid = false;
    std::mutex mu;
};

namespace {

void index_free(b200_index* ix) {
    if (!ix) return;
    cudaSetDevice(ix->device);
    cudaFree(ix->corpus);
    cudaFree(ix->doc_of_row);
    cudaFree(ix->row_n2);
    cudaFree(ix->max_n2);
    cudaFree(ix->d_flags);
    cudaFree(ix->qh);
    cudaFree(ix->q_stage);
    cudaFree(ix->list_score);
    cudaFree(ix->list_row);
    cudaFree(ix->list_doc);
    cudaFree(ix->o_doc);
    cudaFree(ix->o_row);
    cudaFree(ix->o_score);
    cudaFree(ix->qs);
    if (ix->h_qs) cudaFreeHost(ix->h_qs);
    cudaFree(ix->cbuf);
    for (double* c : ix->attr_cols) cudaFree(c);
    cudaFree(ix->d_cols_table);
    cudaFree(ix->mod64);
    cudaFree(ix->mod32);
    cudaFree(ix->mod_max);
    cudaFree(ix->filter_bits);
    for (auto& e : ix->ev)
        if (e) cudaEventDestroy(e);
    if (ix->own_stream) cudaStreamDestroy(ix->own_stream);
    delete ix;
}

void cuda_alloc(void** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        fail(B200_ERR_OOM, "cudaMalloc(%zu bytes) failed: out of device memory", bytes);
    }
    MB_CUDA(e);
}

struct DevBuf {   // RAII scratch allocation
    void* p = nullptr;
    DevBuf() = default;
    explicit DevBuf(size_t bytes) { cuda_alloc(&p, bytes); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { cudaFree(p); }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

void ensure_capacity(b200_index* ix, int64_t need_rows) {
    if (need_rows <= ix->capacity) return;
    int64_t cap = std::max<int64_t>(need_rows, ix->capacity + ix->capacity / 2);
    cap = (int64_t)round_up((size_t)cap, TILE_N);
    __half* nc = nullptr;
    int32_t* nd = nullptr;
    float* nn = nullptr;
    try {
        cuda_alloc((void**)&nc, (size_t)cap * ix->dim * sizeof(__half));
        cuda_alloc((void**)&nd, (size_t)cap * sizeof(int32_t));
        if (ix->metric == B200_METRIC_EUCLIDEAN) cuda_alloc((void**)&nn, (size_t)cap * sizeof(float));
    } catch (...) {
        cudaFree(nc);
        cudaFree(nd);
        cudaFree(nn);
        throw;
    }
    if (ix->n_rows > 0) {
        MB_CUDA(cudaMemcpyAsync(nc, ix->corpus, (size_t)ix->n_rows * ix->dim * sizeof(__half), cudaMemcpyDeviceToDevice,
                                ix->stream));
        MB_CUDA(cudaMemcpyAsync(nd, ix->doc_of_row, (size_t)ix->n_rows * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                                ix->stream));
        if (nn)
            MB_CUDA(cudaMemcpyAsync(nn, ix->row_n2, (size_t)ix->n_rows * sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
    }
    MB_CUDA(cudaStreamSynchronize(ix->stream));
    cudaFree(ix->corpus);
    cudaFree(ix->doc_of_row);
    cudaFree(ix->row_n2);
    ix->corpus = nc;
    ix->doc_of_row = nd;
    ix->row_n2 = nn;
    ix->capacity = cap;
}

void ensure_out_k(b200_index* ix, int k) {
    if (k <= ix->out_k) return;
    const int nk = std::max(k, 16);
    MB_CUDA(cudaStreamSynchronize(ix->stream));
    cudaFree(ix->o_doc);
    cudaFree(ix->o_row);
    cudaFree(ix->o_score);
    ix->o_doc = ix->o_row = nullptr;
    ix->o_score = nullptr;
    ix->out_k = 0;
    cuda_alloc((void**)&ix->o_doc, (size_t)MQ * nk * sizeof(int32_t));
    cuda_alloc((void**)&ix->o_row, (size_t)MQ * nk * sizeof(int32_t));
    cuda_alloc((void**)&ix->o_score, (size_t)MQ * nk * sizeof(double));
    ix->out_k = nk;
}

void ensure_ccap(b200_index* ix, int cap) {
    if (cap <= ix->ccap) return;
    MB_CUDA(cudaStreamSynchronize(ix->stream));
    cudaFree(ix->cbuf);
    ix->cbuf = nullptr;
    ix->ccap = 0;
    cuda_alloc((void**)&ix->cbuf, (size_t)MQ * cap * sizeof(int32_t));
    ix->ccap = cap;
}

using ScanFn = void (*)(const CUtensorMap, const CUtensorMap, ScanParams);
template <bool C>
ScanFn scan_fn(int idx) {
    static const ScanFn table[8] = {scan_kernel<false, false, false, C>, scan_kernel<true, false, false, C>,
                                    scan_kernel<false, true, false, C>,  scan_kernel<true, true, false, C>,
                                    scan_kernel<false, false, true, C>,  scan_kernel<true, false, true, C>,
                                    scan_kernel<false, true, true, C>,   scan_kernel<true, true, true, C>};
    return table[idx];
}

b200_index* index_new(int device, int dim, int metric, int64_t capacity_rows) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        fail(B200_ERR_NO_DEVICE, "no CUDA device available (marqo_b200 has no CPU fallback)");
    }
    MB_CHECK_ARG(device >= 0 && device < ndev, "device %d out of range (%d devices)", device, ndev);
    int major = 0;
    MB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) fail(B200_ERR_NO_DEVICE, "device %d has compute capability %d.x; sm_100 required", device, major);
    MB_CHECK_ARG(dim > 0 && dim % BLOCK_K == 0 && dim <= MAX_DIM, "dim must be a multiple of %d and <= %d (got %d)",
                 BLOCK_K, MAX_DIM, dim);
    MB_CHECK_ARG(metric >= 0 && metric <= B200_METRIC_EUCLIDEAN, "unknown metric %d", metric);
    MB_CHECK_ARG(capacity_rows >= 0, "capacity_rows must be >= 0");
    DeviceGuard g(device);
    b200_index* ix = new b200_index();
    try {
        ix->device = device;
        ix->dim = dim;
        ix->metric = metric;
        ix->sms = std::min(sm_count(device), MAX_LISTS);   // merge_kernel's list-head table holds MAX_LISTS lists
        MB_CUDA(cudaStreamCreateWithFlags(&ix->own_stream, cudaStreamNonBlocking));
        ix->stream = ix->own_stream;
        for (auto& e : ix->ev) MB_CUDA(cudaEventCreate(&e));
        cuda_alloc((void**)&ix->qh, (size_t)MQ * dim * sizeof(__half));
        cuda_alloc((void**)&ix->q_stage, (size_t)MQ * dim * sizeof(float));
        const size_t nl = (size_t)ix->sms * MQ * KP;
        cuda_alloc((void**)&ix->list_score, nl * sizeof(float));
        cuda_alloc((void**)&ix->list_row, nl * sizeof(int32_t));
        cuda_alloc((void**)&ix->list_doc, nl * sizeof(int32_t));
        cuda_alloc((void**)&ix->qs, sizeof(QState));
        MB_CUDA(cudaMemset(ix->qs, 0, sizeof(QState)));
        MB_CUDA(cudaHostAlloc((void**)&ix->h_qs, sizeof(QState), cudaHostAllocPortable));
        cuda_alloc((void**)&ix->max_n2, sizeof(float));
        MB_CUDA(cudaMemset(ix->max_n2, 0, sizeof(float)));
        cuda_alloc((void**)&ix->d_flags, 2 * sizeof(int));
        MB_CUDA(cudaMemset(ix->d_flags, 0, 2 * sizeof(int)));
        cuda_alloc((void**)&ix->mod_max, 2 * sizeof(double));
        MB_CUDA(cudaMemset(ix->mod_max, 0, 2 * sizeof(double)));
        cuda_alloc((void**)&ix->d_cols_table, B200_MAX_ATTRIBUTE_COLUMNS * sizeof(double*));
        ensure_out_k(ix, 16);
        ensure_ccap(ix, FIN_CAP);
        ensure_capacity(ix, std::max<int64_t>(capacity_rows, TILE_N));
        const auto smem_attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
        for (int i = 0; i < 8; ++i) {
            MB_CUDA(cudaFuncSetAttribute(scan_fn<false>(i), smem_attr, SMEM_LIMIT));
            MB_CUDA(cudaFuncSetAttribute(scan_fn<true>(i), smem_attr, SMEM_LIMIT));
        }
        MB_CUDA(cudaFuncSetAttribute(merge_kernel, smem_attr, 64 * 1024));
        MB_CUDA(cudaFuncSetAttribute(finalize_kernel, smem_attr, (int)finalize_smem_bytes(FIN_CAP)));
    } catch (...) {
        index_free(ix);
        throw;
    }
    return ix;
}

// Raises B200_ERR_INVALID_ARG when the last conversion saw a non-finite / out-of-fp16-range value (needs a
// synchronised stream).
void check_input_flags(b200_index* ix, const char* what) {
    int flags = 0;
    MB_CUDA(cudaMemcpyAsync(&flags, ix->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ix->stream));
    MB_CUDA(cudaStreamSynchronize(ix->stream));
    if (flags & 1) {
        MB_CUDA(cudaMemsetAsync(ix->d_flags, 0, sizeof(int), ix->stream));
        fail(B200_ERR_INVALID_ARG,
             "%s contain a value that is not finite or does not fit the fp16 row store (|x| <= 65504 after "
             "normalisation)", what);
    }
}

void add_rows_device(b200_index* ix, const float* d_vecs, const int32_t* d_doc_ids, int64_t m) {
    ensure_capacity(ix, ix->n_rows + m);
    const int wpb = 8;
    const int64_t blocks = (m + wpb - 1) / wpb;
    convert_rows_kernel<<<(unsigned)blocks, wpb * 32, 0, ix->stream>>>(
        d_vecs, ix->corpus + (size_t)ix->n_rows * ix->dim, m, ix->dim, m, ix->metric == B200_METRIC_ANGULAR,
        ix->metric == B200_METRIC_EUCLIDEAN ? ix->row_n2 + ix->n_rows : nullptr, ix->max_n2, ix->d_flags);
    MB_CUDA(cudaGetLastError());
    if (d_doc_ids) {
        MB_CUDA(cudaMemcpyAsync(ix->doc_of_row + ix->n_rows, d_doc_ids, (size_t)m * sizeof(int32_t),
                                cudaMemcpyDeviceToDevice, ix->stream));
        ix->has_docs = true;
    } else {
        iota_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ix->stream>>>(ix->doc_of_row + ix->n_rows, m,
                                                                        (int32_t)ix->n_rows);
        MB_CUDA(cudaGetLastError());
    }
    ix->n_rows += m;
}

struct GroupOut {   // device [g, k]
    int32_t* doc;
    int32_t* row;
    double* score;
};

ExactParams exact_params(b200_index* ix, int nq, int k, const GroupOut& out) {
    ExactParams ex{};
    ex.nq = nq;
    ex.k = k;
    ex.dim = ix->dim;
    ex.metric = ix->metric;
    ex.doc_offset = ix->doc_offset;
    ex.qh = ix->qh;
    ex.corpus = ix->corpus;
    ex.doc_of_row = ix->doc_of_row;
    ex.mod64 = ix->mod_active ? ix->mod64 : nullptr;
    ex.qs = ix->qs;
    ex.out_doc = out.doc;
    ex.out_row = out.row;
    ex.out_score = out.score;
    return ex;
}

struct ScanLaunch {
    CUtensorMap tmap_c, tmap_q;
    ScanParams sp;
    int grid;
    size_t smem;
    int fn_index;
};

ScanLaunch prepare_scan(b200_index* ix, int nq) {
    ScanLaunch L{};
    const int num_tiles = (int)((ix->n_rows + TILE_N - 1) / TILE_N);
    L.grid = std::min(num_tiles, ix->sms);
    const bool mod = ix->mod_active;
    int stages = 16;
    while (stages > 2 && scan_smem_bytes(ix->dim, stages, mod) > (size_t)SMEM_LIMIT) --stages;
    L.smem = scan_smem_bytes(ix->dim, stages, mod);
    if (L.smem > (size_t)SMEM_LIMIT) fail(B200_ERR_INTERNAL, "scan kernel shared memory budget exceeded");
    L.tmap_c = make_tmap_2d(ix->corpus, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)ix->dim, (uint64_t)ix->n_rows,
                            (uint64_t)ix->dim * 2, BLOCK_K, TILE_N, CU_TENSOR_MAP_SWIZZLE_128B);
    L.tmap_q = make_tmap_2d(ix->qh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)ix->dim, (uint64_t)MQ,
                            (uint64_t)ix->dim * 2, BLOCK_K, MQ, CU_TENSOR_MAP_SWIZZLE_128B);
    ScanParams& sp = L.sp;
    sp.n_rows = (int)ix->n_rows;
    sp.dim = ix->dim;
    sp.num_tiles = num_tiles;
    sp.num_stages = stages;
    sp.nq = nq;
    sp.metric = ix->metric;
    sp.doc_of_row = ix->doc_of_row;
    sp.row_bias = ix->row_n2;
    sp.mod = mod ? ix->mod32 : nullptr;
    sp.filter = ix->filter_active ? ix->filter_bits : nullptr;
    sp.filter_docs = ix->filter_docs;
    sp.qs = ix->qs;
    sp.out_score = ix->list_score;
    sp.out_row = ix->list_row;
    sp.out_doc = ix->list_doc;
    sp.cbuf = ix->cbuf;
    sp.ccap = ix->ccap;
    const bool bias = ix->metric == B200_METRIC_EUCLIDEAN;
    const bool docs = ix->has_docs || ix->filter_active;   // the filter is applied where the document numbers are read
    L.fn_index = (docs ? 1 : 0) | (bias ? 2 : 0) | (mod ? 4 : 0);
    return L;
}

void launch_collect(b200_index* ix, ScanLaunch& L) {
    L.sp.cbuf = ix->cbuf;
    L.sp.ccap = ix->ccap;
    scan_fn<true>(L.fn_index)<<<L.grid, THREADS, L.smem, ix->stream>>>(L.tmap_c, L.tmap_q, L.sp);
    MB_CUDA(cudaGetLastError());
}

void launch_finalize(b200_index* ix, int nq, int k, const GroupOut& out) {
    FinalizeParams fp{};
    fp.cbuf = ix->cbuf;
    fp.ccap = ix->ccap;
    fp.live_bound = ix->n_rows;
    fp.ex = exact_params(ix, nq, k, out);
    reset_need_kernel<<<1, 1, 0, ix->stream>>>(ix->qs);
    finalize_kernel<<<nq, FIN_THREADS, finalize_smem_bytes(std::min(ix->ccap, FIN_CAP)), ix->stream>>>(fp);
    MB_CUDA(cudaGetLastError());
}

void fetch_qstate(b200_index* ix) {
    MB_CUDA(cudaMemcpyAsync(ix->h_qs, ix->qs, sizeof(QState), cudaMemcpyDeviceToHost, ix->stream));
    MB_CUDA(cudaStreamSynchronize(ix->stream));
}

// Exact selection of one query's collected rows on the host (more rows than the device finalize holds in shared
// memory: deep pagination, thousands of exact ties).  Keys are computed on the device (exact_keys_kernel); the
// dedup / sort is the same rule as exact_select.  Returns true when the query is resolved.
bool host_finalize(b200_index* ix, int q, int nq, int k, const GroupOut& out, int cnt) {
    ++ix->stat_host_finalize;
    DevBuf d_dot((size_t)cnt * 8), d_key((size_t)cnt * 8), d_doc((size_t)cnt * 4);
    GroupOut dummy{nullptr, nullptr, nullptr};
    ExactParams ex = exact_params(ix, nq, k, dummy);
    const int32_t* rows = ix->cbuf + (size_t)q * ix->ccap;
    exact_keys_kernel<<<(cnt + 7) / 8, 256, 0, ix->stream>>>(ex, q, rows, cnt, d_dot.as<double>(), d_key.as<double>(),
                                                            d_doc.as<int32_t>());
    MB_CUDA(cudaGetLastError());
    struct Hit {
        double dot, key;
        int32_t doc, row;
    };
    std::vector<double> h_dot(cnt), h_key(cnt);
    std::vector<int32_t> h_doc(cnt), h_row(cnt);
    MB_CUDA(cudaMemcpyAsync(h_dot.data(), d_dot.p, (size_t)cnt * 8, cudaMemcpyDeviceToHost, ix->stream));
    MB_CUDA(cudaMemcpyAsync(h_key.data(), d_key.p, (size_t)cnt * 8, cudaMemcpyDeviceToHost, ix->stream));
    MB_CUDA(cudaMemcpyAsync(h_doc.data(), d_doc.p, (size_t)cnt * 4, cudaMemcpyDeviceToHost, ix->stream));
    MB_CUDA(cudaMemcpyAsync(h_row.data(), rows, (size_t)cnt * 4, cudaMemcpyDeviceToHost, ix->stream));
    MB_CUDA(cudaStreamSynchronize(ix->stream));
    std::vector<Hit> v(cnt);
    for (int i = 0; i < cnt; ++i) v[i] = {h_dot[i], h_key[i], h_doc[i], h_row[i]};
    // best chunk per document: (dot desc, row asc); then documents by (key desc, doc asc)
    std::sort(v.begin(), v.end(), [](const Hit& a, const Hit& b) {
        return a.doc < b.doc || (a.doc == b.doc && (a.dot > b.dot || (a.dot == b.dot && a.row < b.row)));
    });
    size_t w = 0;
    for (size_t i = 0; i < v.size(); ++i)
        if (i == 0 || v[i].doc != v[i - 1].doc) v[w++] = v[i];
    v.resize(w);
    std::sort(v.begin(), v.end(), [](const Hit& a, const Hit& b) { return a.key > b.key || (a.key == b.key && a.doc < b.doc); });
    const int nd = (int)v.size();
    QState* h = ix->h_qs;
    const double L = h->L[q], eps = h->eps[q];
    double ek = -std::numeric_limits<double>::infinity();
    if (nd >= k) ek = ix->mod_active ? v[k - 1].key : (ix->metric == B200_METRIC_EUCLIDEAN ? v[k - 1].key + h->qn2x[q] : v[k - 1].key);
    const bool resolved = std::isinf(L) ? (L < 0) : (nd >= k && ek >= L + eps);
    std::vector<int32_t> o_doc(k, -1), o_row(k, -1);
    std::vector<double> o_sc(k, -std::numeric_limits<double>::infinity());
    for (int i = 0; i < k && i < nd; ++i) {
        o_doc[i] = v[i].doc + ix->doc_offset;
        o_row[i] = v[i].row;
        if (ix->mod_active) o_sc[i] = v[i].key;
        else {
            const double d = v[i].key;
            switch (ix->metric) {   // same expressions as closeness_from_dot
                case B200_METRIC_EUCLIDEAN: o_sc[i] = 1.0 / (1.0 + std::sqrt(std::max(-d, 0.0))); break;
                case B200_METRIC_PRENORMALIZED_ANGULAR: o_sc[i] = 1.0 / (1.0 + (1.0 - d)); break;
                case B200_METRIC_ANGULAR: o_sc[i] = 1.0 / (1.0 + std::acos(std::min(1.0, std::max(-1.0, d)))); break;
                default: o_sc[i] = d;
            }
        }
    }
    MB_CUDA(cudaMemcpyAsync(out.doc + (size_t)q * k, o_doc.data(), (size_t)k * 4, cudaMemcpyHostToDevice, ix->stream));
    MB_CUDA(cudaMemcpyAsync(out.row + (size_t)q * k, o_row.data(), (size_t)k * 4, cudaMemcpyHostToDevice, ix->stream));
    MB_CUDA(cudaMemcpyAsync(out.score + (size_t)q * k, o_sc.data(), (size_t)k * 8, cudaMemcpyHostToDevice, ix->stream));
    MB_CUDA(cudaStreamSynchronize(ix->stream));
    h->ndocs[q] = nd;
    h->ek[q] = ek;
    if (resolved) {
        h->status[q] = Q_RESOLVED;
        h->L[q] = INFINITY;
    } else {
        float nl;
        if (nd >= k) nl = std::nextafterf((float)(ek - eps), -INFINITY);
        else nl = (float)L - std::max(8.0f * (float)eps, 0.05f * std::fabs((float)L));
        if (!(nl < (float)L)) nl = -INFINITY;
        h->L[q] = nl;
    }
    return resolved;
}

// One group of <= MQ queries already converted into ix->qh.
//   may_sync: the caller tolerates host synchronisation — flagged queries are driven to resolution here.
//   otherwise: one collect + finalize pass is enqueued unconditionally (both exit at once when nothing is flagged);
//   queries still unresolved after it are counted in QState::rounds (b200_index_search_stats).
void search_group(b200_index* ix, int nq, int k, const GroupOut& out, bool record_timing, bool may_sync) {
    const int total = nq * k;
    ++ix->stat_groups;
    if (ix->n_rows == 0) {
        fill_empty_kernel<<<(total + 255) / 256, 256, 0, ix->stream>>>(out.doc, out.row, out.score, total);
        MB_CUDA(cudaGetLastError());
        return;
    }
    query_prep_kernel<<<MQ / 8, 256, 0, ix->stream>>>(ix->qh, ix->dim, ix->metric, ix->mod_active ? 1 : 0, ix->max_n2,
                                                     ix->mod_max, ix->qs);
    MB_CUDA(cudaGetLastError());
    ScanLaunch L = prepare_scan(ix, nq);
    if (record_timing) MB_CUDA(cudaEventRecord(ix->ev[0], ix->stream));
    scan_fn<false>(L.fn_index)<<<L.grid, THREADS, L.smem, ix->stream>>>(L.tmap_c, L.tmap_q, L.sp);
    MB_CUDA(cudaGetLastError());
    if (record_timing) MB_CUDA(cudaEventRecord(ix->ev[1], ix->stream));

    MergeParams mp{};
    mp.num_lists = L.grid;
    int sort_n = 32;
    while (sort_n < L.grid * KP) sort_n <<= 1;
    mp.sort_n = sort_n;
    mp.l_only = k > K_MERGE_MAX ? 1 : 0;
    mp.in_score = ix->list_score;
    mp.in_row = ix->list_row;
    mp.in_doc = ix->list_doc;
    mp.ex = exact_params(ix, nq, k, out);
    merge_kernel<<<nq, MERGE_THREADS, (size_t)sort_n * 12, ix->stream>>>(mp);
    MB_CUDA(cudaGetLastError());
    if (record_timing) {
        MB_CUDA(cudaEventRecord(ix->ev[2], ix->stream));
        ix->timing_valid = true;
    }
    if (!may_sync) {
        launch_collect(ix, L);
        launch_finalize(ix, nq, k, out);
        note_unresolved_kernel<<<1, 1, 0, ix->stream>>>(ix->qs);
        MB_CUDA(cudaGetLastError());
        return;
    }
    fetch_qstate(ix);
    QState* h = ix->h_qs;
    if (h->n_need == 0) return;
    ix->stat_flagged += h->n_need;
    for (int round = 0; h->n_need > 0; ++round) {
        if (round >= 64) fail(B200_ERR_INTERNAL, "exact top-k did not converge in %d collect passes", round);
        ++ix->stat_collect_passes;
        if (round >= 12)   // stop lowering L step by step: take everything
            for (int q = 0; q < nq; ++q)
                if (h->status[q] == Q_NEED) h->L[q] = -INFINITY;
        MB_CUDA(cudaMemcpyAsync(ix->qs, h, sizeof(QState), cudaMemcpyHostToDevice, ix->stream));
        launch_collect(ix, L);
        fetch_qstate(ix);
        int max_cnt = 0;
        for (int q = 0; q < nq; ++q)
            if (h->status[q] == Q_NEED) max_cnt = std::max(max_cnt, h->cnt[q]);
        if (max_cnt > ix->ccap) {   // grow the buffer and repeat the pass with the same thresholds
            if ((size_t)MQ * max_cnt * 4 > ((size_t)8 << 30))
                fail(B200_ERR_OOM, "exact top-k needs %d candidate rows per query (massive ties); not supported", max_cnt);
            ensure_ccap(ix, (int)round_up((size_t)max_cnt + max_cnt / 8, 1024));
            for (int q = 0; q < nq; ++q) h->cnt[q] = 0;
            --round;
            continue;
        }
        if (max_cnt <= FIN_CAP) {
            launch_finalize(ix, nq, k, out);
            fetch_qstate(ix);
        } else {
            int need = 0;
            for (int q = 0; q < nq; ++q) {
                if (h->status[q] != Q_NEED) continue;
                if (!host_finalize(ix, q, nq, k, out, h->cnt[q])) ++need;
                h->cnt[q] = 0;
            }
            h->n_need = need;
        }
    }
    // leave the device copy clean for the next group (status / thresholds are re-initialised by query_prep_kernel)
}

void convert_queries(b200_index* ix, const float* d_q, int g) {
    convert_rows_kernel<<<MQ / 8, 256, 0, ix->stream>>>(d_q, ix->qh, g, ix->dim, MQ, ix->metric == B200_METRIC_ANGULAR,
                                                        nullptr, nullptr, ix->d_flags);
    MB_CUDA(cudaGetLastError());
}

void search_device(b200_index* ix, const float* d_q, int nq, int k, int32_t* d_out_doc, int32_t* d_out_row,
                   double* d_out_score, bool may_sync) {
    for (int q0 = 0; q0 < nq; q0 += MQ) {
        const int g = std::min(MQ, nq - q0);
        convert_queries(ix, d_q + (size_t)q0 * ix->dim, g);
        GroupOut out{d_out_doc + (size_t)q0 * k, d_out_row + (size_t)q0 * k, d_out_score + (size_t)q0 * k};
        search_group(ix, g, k, out, q0 + MQ >= nq, may_sync);
    }
}

void check_search_args(b200_index* ix, const void* q, int nq, int k, const void* a, const void* b, const void* c) {
    MB_CHECK_ARG(ix != nullptr, "index is NULL");
    MB_CHECK_ARG(q && a && b && c, "NULL buffer");
    MB_CHECK_ARG(nq > 0, "nq must be positive (got %d)", nq);
    MB_CHECK_ARG(k > 0, "k must be positive (got %d)", k);
    MB_CHECK_ARG(k <= 11000, "k = %d exceeds 11000 (Marqo's own limit + offset cap, api/configs.py:24-25)", k);
}

// largest document number among device-resident ids (ingest path; sizes the score-modifier tables)
void track_max_doc(b200_index* ix, const int32_t* d_ids, int64_t m) {
    std::vector<int32_t> h((size_t)m);
    MB_CUDA(cudaMemcpy(h.data(), d_ids, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int32_t v : h) ix->max_doc = std::max<int64_t>(ix->max_doc, v);
}

int64_t num_docs(const b200_index* ix) { return std::max<int64_t>(ix->n_rows, ix->max_doc + 1); }

void fill_nan(b200_index* ix, double* dst, int64_t n) {
    if (n <= 0) return;
    fill_nan_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ix->stream>>>(dst, n);
    MB_CUDA(cudaGetLastError());
}

void ensure_attr_capacity(b200_index* ix, int64_t need_docs) {
    if (need_docs <= ix->attr_cap) return;
    const int64_t cap = (int64_t)round_up((size_t)std::max<int64_t>(need_docs, ix->attr_cap + ix->attr_cap / 2), 1024);
    for (double*& col : ix->attr_cols) {
        if (!col) continue;
        double* nc = nullptr;
        cuda_alloc((void**)&nc, (size_t)cap * sizeof(double));
        if (ix->attr_cap > 0)
            MB_CUDA(cudaMemcpyAsync(nc, col, (size_t)ix->attr_cap * sizeof(double), cudaMemcpyDeviceToDevice, ix->stream));
        fill_nan(ix, nc + ix->attr_cap, cap - ix->attr_cap);
        MB_CUDA(cudaStreamSynchronize(ix->stream));
        cudaFree(col);
        col = nc;
    }
    ix->attr_cap = cap;
    ix->cols_table_dirty = true;
}

double* attr_column(b200_index* ix, int column) {
    if ((int)ix->attr_cols.size() <= column) ix->attr_cols.resize(column + 1, nullptr);
    if (!ix->attr_cols[column]) {
        cuda_alloc((void**)&ix->attr_cols[column], (size_t)ix->attr_cap * sizeof(double));
        fill_nan(ix, ix->attr_cols[column], ix->attr_cap);
        ix->cols_table_dirty = true;
    }
    return ix->attr_cols[column];
}

void sync_cols_table(b200_index* ix) {
    if (!ix->cols_table_dirty) return;
    double* table[B200_MAX_ATTRIBUTE_COLUMNS] = {nullptr};
    for (size_t c = 0; c < ix->attr_cols.size(); ++c) table[c] = ix->attr_cols[c];
    MB_CUDA(cudaMemcpyAsync(ix->d_cols_table, table, sizeof(table), cudaMemcpyHostToDevice, ix->stream));
    MB_CUDA(cudaStreamSynchronize(ix->stream));   // `table` is a stack array
    ix->cols_table_dirty = false;
}

struct ModScope {  // marks the index as "searching with modifiers" for the duration of one call
    b200_index* ix;
    explicit ModScope(b200_index* i, bool on) : ix(i) { ix->mod_active = on; }
    ~ModScope() { ix->mod_active = false; }
};
struct FilterScope {
    b200_index* ix;
    explicit FilterScope(b200_index* i, bool on) : ix(i) { ix->filter_active = on; }
    ~FilterScope() { ix->filter_active = false; }
};

void prepare_modifiers(b200_index* ix, const int32_t* mult_cols, const double* mult_w, int n_mult,
                       const int32_t* add_cols, const double* add_w, int n_add) {
    const int64_t nd = num_docs(ix);
    if (nd > ix->mod_cap) {
        cudaFree(ix->mod64);
        cudaFree(ix->mod32);
        ix->mod64 = nullptr;
        ix->mod32 = nullptr;
        ix->mod_cap = 0;
        const int64_t cap = (int64_t)round_up((size_t)nd + (size_t)nd / 2, 1024);
        cuda_alloc((void**)&ix->mod64, (size_t)cap * sizeof(double2));
        cuda_alloc((void**)&ix->mod32, (size_t)cap * sizeof(float2));
        ix->mod_cap = cap;
    }
    if (nd == 0) return;
    ModifierParams mp{};
    mp.n_docs = (int)nd;
    mp.attr_cap = (int)ix->attr_cap;
    mp.n_mult = n_mult;
    mp.n_add = n_add;
    auto col = [&](int c) -> const double* {
        MB_CHECK_ARG(c >= 0 && c < B200_MAX_ATTRIBUTE_COLUMNS, "attribute column %d out of range", c);
        return c < (int)ix->attr_cols.size() ? ix->attr_cols[c] : nullptr;   // never-set column: missing everywhere
    };
    for (int i = 0; i < n_mult; ++i) {
        mp.mult_col[i] = col(mult_cols[i]);
        mp.mult_w[i] = mult_w[i];
    }
    for (int i = 0; i < n_add; ++i) {
        mp.add_col[i] = col(add_cols[i]);
        mp.add_w[i] = add_w[i];
    }
    mp.out64 = ix->mod64;
    mp.out32 = ix->mod32;
    mp.negative_flag = ix->d_flags + 1;
    mp.mod_max = ix->mod_max;
    MB_CUDA(cudaMemsetAsync(ix->d_flags + 1, 0, sizeof(int), ix->stream));
    MB_CUDA(cudaMemsetAsync(ix->mod_max, 0, 2 * sizeof(double), ix->stream));
    modifier_kernel<<<(unsigned)((nd + 255) / 256), 256, 0, ix->stream>>>(mp);
    MB_CUDA(cudaGetLastError());
    int flag = 0;
    MB_CUDA(cudaMemcpyAsync(&flag, ix->d_flags + 1, sizeof(int), cudaMemcpyDeviceToHost, ix->stream));
    MB_CUDA(cudaStreamSynchronize(ix->stream));
    // closeness(field, embeddings) is the best chunk's closeness; the scan keeps, per document, the chunk with the best
    // MODIFIED key, which is the same chunk only while the multiplier is >= 0.
    if (flag && ix->has_docs)
        fail(B200_ERR_UNSUPPORTED,
             "a negative multiplicative score modifier on a corpus with explicit document ids (multi-chunk documents) "
             "is not supported");
}

// Upload the caller's document bitset unless the device already holds the bitset with this tag.
void prepare_filter(b200_index* ix, const uint32_t* bits, int64_t n_docs, uint64_t tag) {
    MB_CHECK_ARG(bits != nullptr && n_docs >= 0, "filter_bits is NULL or filter_docs < 0");
    const int64_t words = (n_docs + 31) / 32;
    if (tag != 0 && tag == ix->filter_tag && n_docs == ix->filter_docs) return;
    if (words > ix->filter_cap_words) {
        MB_CUDA(cudaStreamSynchronize(ix->stream));
        cudaFree(ix->filter_bits);
        ix->filter_bits = nullptr;
        ix->filter_cap_words = 0;
        ix->filter_tag = 0;
        const int64_t cap = (int64_t)round_up((size_t)words + (size_t)words / 2 + 1, 256);
        cuda_alloc((void**)&ix->filter_bits, (size_t)cap * 4);
        ix->filter_cap_words = cap;
    }
    if (words > 0)
        MB_CUDA(cudaMemcpyAsync(ix->filter_bits, bits, (size_t)words * 4, cudaMemcpyHostToDevice, ix->stream));
    MB_CUDA(cudaStreamSynchronize(ix->stream));   // the caller's buffer may be pageable and short-lived
    ix->filter_docs = n_docs;
    ix->filter_tag = tag;
}

void search_host(b200_index* ix, const float* q, int nq, int k, int32_t* out_doc, int32_t* out_row, double* out_score) {
    ensure_out_k(ix, k);
    for (int q0 = 0; q0 < nq; q0 += MQ) {
        const int gq = std::min(MQ, nq - q0);
        MB_CUDA(cudaMemcpyAsync(ix->q_stage, q + (size_t)q0 * ix->dim, (size_t)gq * ix->dim * sizeof(float),
                                cudaMemcpyHostToDevice, ix->stream));
        search_device(ix, ix->q_stage, gq, k, ix->o_doc, ix->o_row, ix->o_score, true);
        MB_CUDA(cudaMemcpyAsync(out_doc + (size_t)q0 * k, ix->o_doc, (size_t)gq * k * sizeof(int32_t),
                                cudaMemcpyDeviceToHost, ix->stream));
        MB_CUDA(cudaMemcpyAsync(out_row + (size_t)q0 * k, ix->o_row, (size_t)gq * k * sizeof(int32_t),
                                cudaMemcpyDeviceToHost, ix->stream));
        MB_CUDA(cudaMemcpyAsync(out_score + (size_t)q0 * k, ix->o_score, (size_t)gq * k * sizeof(double),
                                cudaMemcpyDeviceToHost, ix->stream));
        check_input_flags(ix, "queries");   // synchronises
    }
}

void validate_opts(const b200_search_opts* o) {
    if (!o) return;
    MB_CHECK_ARG(o->n_mult >= 0 && o->n_mult <= MAX_MOD_TERMS && o->n_add >= 0 && o->n_add <= MAX_MOD_TERMS,
                 "at most %d multiplicative and %d additive modifiers per search", MAX_MOD_TERMS, MAX_MOD_TERMS);
    MB_CHECK_ARG((o->n_mult == 0 || (o->mult_cols && o->mult_w)) && (o->n_add == 0 || (o->add_cols && o->add_w)),
                 "NULL modifier list");
    for (int i = 0; i < o->n_mult; ++i) MB_CHECK_ARG(std::isfinite(o->mult_w[i]), "mult_w[%d] is not finite", i);
    for (int i = 0; i < o->n_add; ++i) MB_CHECK_ARG(std::isfinite(o->add_w[i]), "add_w[%d] is not finite", i);
    MB_CHECK_ARG(o->filter_bits != nullptr || o->filter_docs == 0, "filter_docs > 0 with filter_bits == NULL");
}

}  // namespace

extern "C" {

int b200_index_create(int device, int dim, int metric, int64_t capacity_rows, b200_index** out) {
    return guarded([&] {
        MB_CHECK_ARG(out != nullptr, "out is NULL");
        *out = nullptr;
        *out = index_new(device, dim, metric, capacity_rows);
    });
}

int b200_index_destroy(b200_index* ix) {
    return guarded([&] { index_free(ix); });
}

int b200_index_add(b200_index* ix, const float* vecs, const int32_t* doc_ids, int64_t m) {
    return guarded([&] {
        MB_CHECK_ARG(ix != nullptr, "index is NULL");
        MB_CHECK_ARG(m >= 0, "m must be >= 0");
        if (m == 0) return;
        MB_CHECK_ARG(vecs != nullptr, "vecs is NULL");
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        MB_CHECK_ARG(ix->n_rows + m < (int64_t)INT32_MAX, "row count would exceed 2^31-1");
        int64_t hi = ix->max_doc;
        if (doc_ids)
            for (int64_t i = 0; i < m; ++i) {
                MB_CHECK_ARG(doc_ids[i] >= 0, "doc_ids[%lld] is negative", (long long)i);
                hi = std::max<int64_t>(hi, doc_ids[i]);
            }
        const int64_t chunk = 1 << 16;
        DevBuf d_v((size_t)std::min(m, chunk) * ix->dim * sizeof(float));
        DevBuf d_d(doc_ids ? (size_t)std::min(m, chunk) * sizeof(int32_t) : 16);
        const int64_t rows_before = ix->n_rows;
        const bool docs_before = ix->has_docs;
        try {
            for (int64_t o = 0; o < m; o += chunk) {
                const int64_t c = std::min(chunk, m - o);
                MB_CUDA(cudaMemcpyAsync(d_v.p, vecs + (size_t)o * ix->dim, (size_t)c * ix->dim * sizeof(float),
                                        cudaMemcpyHostToDevice, ix->stream));
                if (doc_ids)
                    MB_CUDA(cudaMemcpyAsync(d_d.p, doc_ids + o, (size_t)c * sizeof(int32_t), cudaMemcpyHostToDevice,
                                            ix->stream));
                add_rows_device(ix, d_v.as<float>(), doc_ids ? d_d.as<int32_t>() : nullptr, c);
                MB_CUDA(cudaStreamSynchronize(ix->stream));
            }
            check_input_flags(ix, "embeddings");
        } catch (...) {   // nothing of a rejected batch stays searchable
            ix->n_rows = rows_before;
            ix->has_docs = docs_before;
            throw;
        }
        ix->max_doc = hi;
    });
}

int b200_index_add_device(b200_index* ix, const float* d_vecs, const int32_t* d_doc_ids, int64_t m) {
    return guarded([&] {
        MB_CHECK_ARG(ix != nullptr, "index is NULL");
        MB_CHECK_ARG(m >= 0, "m must be >= 0");
        if (m == 0) return;
        MB_CHECK_ARG(d_vecs != nullptr, "d_vecs is NULL");
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        MB_CHECK_ARG(ix->n_rows + m < (int64_t)INT32_MAX, "row count would exceed 2^31-1");
        const int64_t rows_before = ix->n_rows;
        const bool docs_before = ix->has_docs;
        add_rows_device(ix, d_vecs, d_doc_ids, m);
        try {
            check_input_flags(ix, "embeddings");
        } catch (...) {
            ix->n_rows = rows_before;
            ix->has_docs = docs_before;
            throw;
        }
        if (d_doc_ids) track_max_doc(ix, d_doc_ids, m);
    });
}

int b200_index_add_device_docs(b200_index* ix, const float* d_vecs, const int32_t* doc_ids, int64_t m) {
    return guarded([&] {
        MB_CHECK_ARG(ix != nullptr, "index is NULL");
        MB_CHECK_ARG(m >= 0, "m must be >= 0");
        if (m == 0) return;
        MB_CHECK_ARG(d_vecs != nullptr && doc_ids != nullptr, "NULL argument");
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        MB_CHECK_ARG(ix->n_rows + m < (int64_t)INT32_MAX, "row count would exceed 2^31-1");
        int64_t hi = ix->max_doc;
        for (int64_t i = 0; i < m; ++i) {
            MB_CHECK_ARG(doc_ids[i] >= 0, "doc_ids[%lld] is negative", (long long)i);
            hi = std::max<int64_t>(hi, doc_ids[i]);
        }
        DevBuf d_d((size_t)m * sizeof(int32_t));
        MB_CUDA(cudaMemcpyAsync(d_d.p, doc_ids, (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, ix->stream));
        const int64_t rows_before = ix->n_rows;
        const bool docs_before = ix->has_docs;
        add_rows_device(ix, d_vecs, d_d.as<int32_t>(), m);
        try {
            check_input_flags(ix, "embeddings");
        } catch (...) {
            ix->n_rows = rows_before;
            ix->has_docs = docs_before;
            throw;
        }
        ix->max_doc = hi;
    });
}

int b200_index_delete_doc(b200_index* ix, int32_t doc_id) {
    return guarded([&] {
        MB_CHECK_ARG(ix != nullptr, "index is NULL");
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        if (ix->n_rows == 0) return;
        tombstone_kernel<<<(unsigned)((ix->n_rows + 255) / 256), 256, 0, ix->stream>>>(ix->doc_of_row, ix->n_rows,
                                                                                      doc_id);
        MB_CUDA(cudaGetLastError());
        MB_CUDA(cudaStreamSynchronize(ix->stream));
        ix->has_docs = true;
    });
}

int b200_index_delete_rows(b200_index* ix, const int32_t* rows, int64_t n) {
    return guarded([&] {
        MB_CHECK_ARG(ix != nullptr, "index is NULL");
        MB_CHECK_ARG(n >= 0, "n must be >= 0");
        if (n == 0) return;
        MB_CHECK_ARG(rows != nullptr, "rows is NULL");
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        for (int64_t i = 0; i < n; ++i)
            MB_CHECK_ARG(rows[i] >= 0 && rows[i] < ix->n_rows, "rows[%lld] = %d out of range", (long long)i, rows[i]);
        DevBuf d_r((size_t)n * sizeof(int32_t));
        MB_CUDA(cudaMemcpyAsync(d_r.p, rows, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ix->stream));
        tombstone_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ix->stream>>>(ix->doc_of_row, d_r.as<int32_t>(), n);
        MB_CUDA(cudaGetLastError());
        MB_CUDA(cudaStreamSynchronize(ix->stream));
        ix->has_docs = true;
        ix->dead_rows += n;
    });
}

int b200_index_compact(b200_index* ix, int32_t* out_new_of_old, int64_t* out_rows) {
    return guarded([&] {
        MB_CHECK_ARG(ix != nullptr && out_new_of_old != nullptr && out_rows != nullptr, "NULL argument");
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        const int64_t n = ix->n_rows;
        std::vector<int32_t> doc((size_t)n);
        if (n > 0) {
            MB_CUDA(cudaMemcpyAsync(doc.data(), ix->doc_of_row, (size_t)n * 4, cudaMemcpyDeviceToHost, ix->stream));
            MB_CUDA(cudaStreamSynchronize(ix->stream));
        }
        int64_t live = 0;
        for (int64_t i = 0; i < n; ++i) out_new_of_old[i] = doc[i] >= 0 ? (int32_t)live++ : -1;
        *out_rows = live;
        if (live == n) {
            ix->dead_rows = 0;
            return;
        }
        const int64_t cap = (int64_t)round_up((size_t)std::max<int64_t>(live, TILE_N), TILE_N);
        __half* nc = nullptr;
        int32_t* nd = nullptr;
        float* nn = nullptr;
        DevBuf d_map((size_t)std::max<int64_t>(n, 1) * 4);
        try {
            cuda_alloc((void**)&nc, (size_t)cap * ix->dim * sizeof(__half));
            cuda_alloc((void**)&nd, (size_t)cap * sizeof(int32_t));
            if (ix->row_n2) cuda_alloc((void**)&nn, (size_t)cap * sizeof(float));
            MB_CUDA(cudaMemcpyAsync(d_map.p, out_new_of_old, (size_t)n * 4, cudaMemcpyHostToDevice, ix->stream));
            compact_rows_kernel<<<(unsigned)((n + 7) / 8), 256, 0, ix->stream>>>(ix->corpus, nc, ix->doc_of_row, nd,
                                                                                ix->row_n2, nn, d_map.as<int32_t>(), n,
                                                                                ix->dim);
            MB_CUDA(cudaGetLastError());
            MB_CUDA(cudaStreamSynchronize(ix->stream));
        } catch (...) {
            cudaFree(nc);
            cudaFree(nd);
            cudaFree(nn);
            throw;
        }
        cudaFree(ix->corpus);
        cudaFree(ix->doc_of_row);
        cudaFree(ix->row_n2);
        ix->corpus = nc;
        ix->doc_of_row = nd;
        ix->row_n2 = nn;
        ix->capacity = cap;
        ix->n_rows = live;
        ix->dead_rows = 0;
    });
}

int b200_index_num_rows(b200_index* ix, int64_t* out_rows) {
    return guarded([&] {
        MB_CHECK_ARG(ix && out_rows, "NULL argument");
        std::lock_guard<std::mutex> lk(ix->mu);
        *out_rows = ix->n_rows;
    });
}

int b200_index_info(b200_index* ix, int* out_dim, int* out_metric, int* out_device) {
    return guarded([&] {
        MB_CHECK_ARG(ix && out_dim && out_metric && out_device, "NULL argument");
        *out_dim = ix->dim;
        *out_metric = ix->metric;
        *out_device = ix->device;
    });
}

int b200_index_get_row(b200_index* ix, int64_t row, float* out_vec) {
    return b200_index_get_rows(ix, &row, 1, out_vec);
}

int b200_index_get_rows(b200_index* ix, const int64_t* rows, int64_t n, float* out_vecs) {
    return guarded([&] {
        MB_CHECK_ARG(ix && out_vecs && (rows || n == 0), "NULL argument");
        MB_CHECK_ARG(n >= 0, "n must be >= 0");
        if (n == 0) return;
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        std::vector<__half> tmp((size_t)n * ix->dim);
        for (int64_t i = 0; i < n; ++i) {
            MB_CHECK_ARG(rows[i] >= 0 && rows[i] < ix->n_rows, "row %lld out of range", (long long)rows[i]);
            MB_CUDA(cudaMemcpyAsync(tmp.data() + (size_t)i * ix->dim, ix->corpus + (size_t)rows[i] * ix->dim,
                                    (size_t)ix->dim * sizeof(__half), cudaMemcpyDeviceToHost, ix->stream));
        }
        MB_CUDA(cudaStreamSynchronize(ix->stream));
        for (size_t i = 0; i < tmp.size(); ++i) out_vecs[i] = __half2float(tmp[i]);
    });
}

int b200_index_search_ex(b200_index* ix, const float* q, int nq, int k, const b200_search_opts* opts, int32_t* out_doc,
                         int32_t* out_row, double* out_score) {
    return guarded([&] {
        check_search_args(ix, q, nq, k, out_doc, out_row, out_score);
        validate_opts(opts);
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        const bool mod = opts && (opts->n_mult > 0 || opts->n_add > 0);
        const bool filt = opts && opts->filter_bits != nullptr;
        if (mod) prepare_modifiers(ix, opts->mult_cols, opts->mult_w, opts->n_mult, opts->add_cols, opts->add_w, opts->n_add);
        if (filt) prepare_filter(ix, opts->filter_bits, opts->filter_docs, opts->filter_tag);
        ModScope ms(ix, mod);
        FilterScope fs(ix, filt);
        search_host(ix, q, nq, k, out_doc, out_row, out_score);
    });
}

int b200_index_search(b200_index* ix, const float* q, int nq, int k, int32_t* out_doc, int32_t* out_row,
                      double* out_score) {
    return b200_index_search_ex(ix, q, nq, k, nullptr, out_doc, out_row, out_score);
}

int b200_index_search_modified(b200_index* ix, const float* q, int nq, int k, const int32_t* mult_cols,
                               const double* mult_w, int n_mult, const int32_t* add_cols, const double* add_w, int n_add,
                               int32_t* out_doc, int32_t* out_row, double* out_score) {
    b200_search_opts o{};
    o.mult_cols = mult_cols;
    o.mult_w = mult_w;
    o.n_mult = n_mult;
    o.add_cols = add_cols;
    o.add_w = add_w;
    o.n_add = n_add;
    return b200_index_search_ex(ix, q, nq, k, &o, out_doc, out_row, out_score);
}

int b200_index_set_attributes(b200_index* ix, int column, const int32_t* doc_ids, const double* values, int64_t n) {
    return guarded([&] {
        MB_CHECK_ARG(ix != nullptr, "index is NULL");
        MB_CHECK_ARG(n >= 0, "n must be >= 0");
        MB_CHECK_ARG(column >= -1 && column < B200_MAX_ATTRIBUTE_COLUMNS, "attribute column %d out of range", column);
        MB_CHECK_ARG(column >= 0 || values == nullptr, "column -1 (all columns) only clears: values must be NULL");
        if (n == 0) return;
        MB_CHECK_ARG(doc_ids != nullptr, "doc_ids is NULL");
        int64_t hi = -1;
        for (int64_t i = 0; i < n; ++i) {
            MB_CHECK_ARG(doc_ids[i] >= 0, "doc_ids[%lld] is negative", (long long)i);
            hi = std::max<int64_t>(hi, doc_ids[i]);
            if (values) MB_CHECK_ARG(std::isfinite(values[i]), "values[%lld] is not finite", (long long)i);
        }
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        if (column < 0 && ix->attr_cols.empty()) return;
        ensure_attr_capacity(ix, hi + 1);
        DevBuf d_ids((size_t)n * sizeof(int32_t));
        DevBuf d_vals(values ? (size_t)n * sizeof(double) : 16);
        MB_CUDA(cudaMemcpyAsync(d_ids.p, doc_ids, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ix->stream));
        if (values)
            MB_CUDA(cudaMemcpyAsync(d_vals.p, values, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ix->stream));
        const unsigned blocks = (unsigned)((n + 255) / 256);
        if (column >= 0) {
            scatter_attr_kernel<<<blocks, 256, 0, ix->stream>>>(attr_column(ix, column), d_ids.as<int32_t>(),
                                                                values ? d_vals.as<double>() : nullptr, n);
            MB_CUDA(cudaGetLastError());
        } else {
            sync_cols_table(ix);
            clear_attr_kernel<<<blocks, 256, 0, ix->stream>>>(ix->d_cols_table, (int)ix->attr_cols.size(),
                                                              d_ids.as<int32_t>(), n, ix->attr_cap);
            MB_CUDA(cudaGetLastError());
        }
        MB_CUDA(cudaStreamSynchronize(ix->stream));
    });
}

int b200_index_set_attributes_multi(b200_index* ix, const int32_t* columns, const int32_t* doc_ids, const double* values,
                                    int64_t n) {
    return guarded([&] {
        MB_CHECK_ARG(ix != nullptr, "index is NULL");
        MB_CHECK_ARG(n >= 0, "n must be >= 0");
        if (n == 0) return;
        MB_CHECK_ARG(columns && doc_ids && values, "NULL argument");
        int64_t hi = -1;
        int max_col = -1;
        for (int64_t i = 0; i < n; ++i) {
            MB_CHECK_ARG(columns[i] >= 0 && columns[i] < B200_MAX_ATTRIBUTE_COLUMNS, "columns[%lld] = %d out of range",
                         (long long)i, columns[i]);
            MB_CHECK_ARG(doc_ids[i] >= 0, "doc_ids[%lld] is negative", (long long)i);
            MB_CHECK_ARG(std::isfinite(values[i]), "values[%lld] is not finite", (long long)i);
            hi = std::max<int64_t>(hi, doc_ids[i]);
            max_col = std::max(max_col, columns[i]);
        }
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        ensure_attr_capacity(ix, hi + 1);
        std::vector<char> used(max_col + 1, 0);
        for (int64_t i = 0; i < n; ++i) used[columns[i]] = 1;
        for (int c = 0; c <= max_col; ++c)
            if (used[c]) attr_column(ix, c);
        sync_cols_table(ix);
        DevBuf d_cols((size_t)n * 4), d_ids((size_t)n * 4), d_vals((size_t)n * 8);
        MB_CUDA(cudaMemcpyAsync(d_cols.p, columns, (size_t)n * 4, cudaMemcpyHostToDevice, ix->stream));
        MB_CUDA(cudaMemcpyAsync(d_ids.p, doc_ids, (size_t)n * 4, cudaMemcpyHostToDevice, ix->stream));
        MB_CUDA(cudaMemcpyAsync(d_vals.p, values, (size_t)n * 8, cudaMemcpyHostToDevice, ix->stream));
        scatter_attr_multi_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ix->stream>>>(
            ix->d_cols_table, d_cols.as<int32_t>(), d_ids.as<int32_t>(), d_vals.as<double>(), n);
        MB_CUDA(cudaGetLastError());
        MB_CUDA(cudaStreamSynchronize(ix->stream));
    });
}

int b200_index_search_device(b200_index* ix, const float* d_q, int nq, int k, int32_t* d_out_doc, int32_t* d_out_row,
                             double* d_out_score, int sync) {
    return guarded([&] {
        check_search_args(ix, d_q, nq, k, d_out_doc, d_out_row, d_out_score);
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        search_device(ix, d_q, nq, k, d_out_doc, d_out_row, d_out_score, sync != 0);
        if (sync) MB_CUDA(cudaStreamSynchronize(ix->stream));
    });
}

int b200_index_search_stats(b200_index* ix, int64_t* out_groups, int64_t* out_flagged, int64_t* out_collect_passes,
                            int64_t* out_unresolved) {
    return guarded([&] {
        MB_CHECK_ARG(ix != nullptr, "index is NULL");
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        fetch_qstate(ix);
        if (out_groups) *out_groups = ix->stat_groups;
        if (out_flagged) *out_flagged = ix->stat_flagged;
        if (out_collect_passes) *out_collect_passes = ix->stat_collect_passes;
        if (out_unresolved) *out_unresolved = ix->h_qs->rounds;
    });
}

int b200_index_set_stream(b200_index* ix, void* cuda_stream, int use_external) {
    return guarded([&] {
        MB_CHECK_ARG(ix != nullptr, "index is NULL");
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        MB_CUDA(cudaStreamSynchronize(ix->stream));
        ix->stream = use_external ? reinterpret_cast<cudaStream_t>(cuda_stream) : ix->own_stream;
    });
}

int b200_index_set_doc_offset(b200_index* ix, int32_t offset) {
    return guarded([&] {
        MB_CHECK_ARG(ix != nullptr, "index is NULL");
        MB_CHECK_ARG(offset >= 0, "offset must be >= 0");
        std::lock_guard<std::mutex> lk(ix->mu);
        ix->doc_offset = offset;
    });
}

int b200_topk_merge_device(b200_index* ix, const void* d_gathered, int nshards, int nq, int k, int32_t* d_out_doc,
                           int32_t* d_out_row, double* d_out_score, int sync) {
    return guarded([&] {
        MB_CHECK_ARG(ix && d_gathered && d_out_doc && d_out_row && d_out_score, "NULL argument");
        MB_CHECK_ARG(nshards > 0 && nq > 0 && k > 0, "nshards, nq, k must be positive");
        if (nshards * k > 256) fail(B200_ERR_UNSUPPORTED, "device merge handles up to 256 candidates per query");
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        merge_shards_kernel<<<(nq + 3) / 4, 128, 0, ix->stream>>>(reinterpret_cast<const uint8_t*>(d_gathered), nshards, nq,
                                                                k, d_out_doc, d_out_row, d_out_score);
        MB_CUDA(cudaGetLastError());
        if (sync) MB_CUDA(cudaStreamSynchronize(ix->stream));
    });
}

int b200_index_last_timing(b200_index* ix, float* scan_ms, float* merge_ms) {
    return guarded([&] {
        MB_CHECK_ARG(ix && scan_ms && merge_ms, "NULL argument");
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        if (!ix->timing_valid) fail(B200_ERR_INVALID_ARG, "no search has been timed yet");
        MB_CUDA(cudaEventSynchronize(ix->ev[2]));
        MB_CUDA(cudaEventElapsedTime(scan_ms, ix->ev[0], ix->ev[1]));
        MB_CUDA(cudaEventElapsedTime(merge_ms, ix->ev[1], ix->ev[2]));
    });
}

int b200_topk_merge(int nshards, int nq, int k, const int32_t* doc, const int32_t* row, const double* score,
                    int32_t* out_doc, int32_t* out_row, double* out_score) {
    return guarded([&] {
        MB_CHECK_ARG(nshards > 0 && nq > 0 && k > 0, "nshards, nq, k must be positive");
        MB_CHECK_ARG(doc && row && score && out_doc && out_row && out_score, "NULL buffer");
        struct Hit {
            double s;
            int32_t d, r;
        };
        std::vector<Hit> hits;
        for (int q = 0; q < nq; ++q) {
            hits.clear();
            for (int s = 0; s < nshards; ++s)
                for (int i = 0; i < k; ++i) {
                    const size_t o = ((size_t)s * nq + q) * k + i;
                    if (doc[o] >= 0) hits.push_back({score[o], doc[o], row[o]});
                }
            std::sort(hits.begin(), hits.end(),
                      [](const Hit& a, const Hit& b) { return a.s > b.s || (a.s == b.s && a.d < b.d); });
            for (int i = 0; i < k; ++i) {
                const size_t o = (size_t)q * k + i;
                if (i < (int)hits.size()) {
                    out_doc[o] = hits[i].d;
                    out_row[o] = hits[i].r;
                    out_score[o] = hits[i].s;
                } else {
                    out_doc[o] = -1;
                    out_row[o] = -1;
                    out_score[o] = -std::numeric_limits<double>::infinity();
                }
            }
        }
    });
}

// ---------------------------------------------------------------------------------------------------- exchange
int b200_exchange_create(int device, int rank, int world, int max_nq, int max_k, b200_exchange** out, void* out_handle) {
    return guarded([&] {
        MB_CHECK_ARG(out && out_handle, "NULL argument");
        *out = nullptr;
        MB_CHECK_ARG(world >= 1 && world <= 8 && rank >= 0 && rank < world, "bad rank/world %d/%d (world <= 8)", rank, world);
        MB_CHECK_ARG(max_nq > 0 && max_k > 0 && world * max_k <= 256, "world * max_k must be <= 256");
        DeviceGuard g(device);
        std::unique_ptr<b200_exchange> ex(new b200_exchange());
        ex->device = device;
        ex->rank = rank;
        ex->world = world;
        ex->slot_stride = round_up((size_t)max_nq * max_k * 16, 256);
        ex->buf_bytes = 2 * (size_t)world * ex->slot_stride + 2 * (size_t)world * sizeof(unsigned long long);
        cuda_alloc((void**)&ex->local, ex->buf_bytes);
        MB_CUDA(cudaMemset(ex->local, 0, ex->buf_bytes));
        MB_CUDA(cudaDeviceSynchronize());
        ex->peer[rank] = ex->local;
        cudaIpcMemHandle_t h;
        MB_CUDA(cudaIpcGetMemHandle(&h, ex->local));
        static_assert(sizeof(cudaIpcMemHandle_t) == B200_EXCHANGE_HANDLE_BYTES, "handle size");
        memcpy(out_handle, &h, sizeof(h));
        *out = ex.release();
    });
}

int b200_exchange_open(b200_exchange* ex, const void* handles) {
    return guarded([&] {
        MB_CHECK_ARG(ex && handles, "NULL argument");
        DeviceGuard g(ex->device);
        const uint8_t* hp = reinterpret_cast<const uint8_t*>(handles);
        for (int s = 0; s < ex->world; ++s) {
            if (s == ex->rank || ex->opened[s]) continue;
            cudaIpcMemHandle_t h;
            memcpy(&h, hp + (size_t)s * B200_EXCHANGE_HANDLE_BYTES, sizeof(h));
            void* p = nullptr;
            MB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            ex->peer[s] = reinterpret_cast<uint8_t*>(p);
            ex->opened[s] = true;
        }
    });
}

int b200_exchange_destroy(b200_exchange* ex) {
    return guarded([&] {
        if (!ex) return;
        cudaSetDevice(ex->device);
        cudaDeviceSynchronize();
        for (int s = 0; s < ex->world; ++s)
            if (ex->opened[s]) cudaIpcCloseMemHandle(ex->peer[s]);
        cudaFree(ex->local);
        delete ex;
    });
}

int b200_index_search_exchange(b200_index* ix, b200_exchange* ex, const float* d_q, int nq, int k, void* d_local_block,
                               int32_t* d_out_doc, int32_t* d_out_row, double* d_out_score, int sync) {
    return guarded([&] {
        MB_CHECK_ARG(ex != nullptr && d_local_block != nullptr, "NULL argument");
        check_search_args(ix, d_q, nq, k, d_out_doc, d_out_row, d_out_score);
        MB_CHECK_ARG(nq <= MQ, "one exchange call handles at most %d queries", MQ);
        MB_CHECK_ARG((size_t)nq * k * 16 <= ex->slot_stride, "nq * k exceeds the exchange buffer's block size");
        MB_CHECK_ARG(ex->world * k <= 256, "world * k must be <= 256");
        MB_CHECK_ARG(ex->device == ix->device, "exchange buffer and index live on different devices");
        for (int s = 0; s < ex->world; ++s) MB_CHECK_ARG(ex->peer[s] != nullptr, "peer %d has not been opened", s);
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        const size_t nk = (size_t)nq * k;
        uint8_t* blk = reinterpret_cast<uint8_t*>(d_local_block);
        search_device(ix, d_q, nq, k, reinterpret_cast<int32_t*>(blk), reinterpret_cast<int32_t*>(blk + nk * 4),
                      reinterpret_cast<double*>(blk + nk * 8), false);
        ExchangeParams p{};
        for (int s = 0; s < ex->world; ++s) {
            p.peer_buf[s] = ex->peer[s];
            p.peer_flag[s] = reinterpret_cast<unsigned long long*>(ex->peer[s] + 2 * (size_t)ex->world * ex->slot_stride);
        }
        p.rank = ex->rank;
        p.world = ex->world;
        p.nq = nq;
        p.k = k;
        p.block_bytes = nk * 16;
        p.slot_stride = ex->slot_stride;
        p.epoch = ex->epoch++;
        p.local_block = blk;
        p.out_doc = d_out_doc;
        p.out_row = d_out_row;
        p.out_score = d_out_score;
        exchange_merge_kernel<<<8, 256, 0, ix->stream>>>(p);   // the SAME grid on every rank (flag arithmetic)
        MB_CUDA(cudaGetLastError());
        if (sync) MB_CUDA(cudaStreamSynchronize(ix->stream));
    });
}

// ---------------------------------------------------------------------------------------------------- persistence
namespace {
struct FileCloser {
    void operator()(FILE* f) const {
        if (f) fclose(f);
    }
};
struct SnapshotHeader {
    char magic[8];
    int32_t version, dim, metric, has_docs;
    int64_t n_rows;
};
}  // namespace

int b200_index_save(b200_index* ix, const char* path) {
    return guarded([&] {
        MB_CHECK_ARG(ix && path, "NULL argument");
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        // written under a temporary name and renamed at the end: a crash or a short write never damages the previous
        // snapshot of the same name
        const std::string tmp_path = std::string(path) + ".tmp";
        std::unique_ptr<FILE, FileCloser> f(fopen(tmp_path.c_str(), "wb"));
        if (!f) fail(B200_ERR_INVALID_ARG, "cannot open %s for writing", tmp_path.c_str());
        SnapshotHeader hdr;
        memcpy(hdr.magic, "B200IDX\0", 8);
        hdr.version = 2;   // 2 = version 1 + the attribute-column trailer
        hdr.dim = ix->dim;
        hdr.metric = ix->metric;
        hdr.has_docs = ix->has_docs ? 1 : 0;
        hdr.n_rows = ix->n_rows;
        bool ok = fwrite(&hdr, sizeof(hdr), 1, f.get()) == 1;
        std::vector<uint8_t> buf((size_t)1 << 24);
        auto dump = [&](const void* dptr, size_t bytes) {
            for (size_t o = 0; o < bytes && ok; o += buf.size()) {
                const size_t c = std::min(buf.size(), bytes - o);
                MB_CUDA(cudaMemcpy(buf.data(), (const uint8_t*)dptr + o, c, cudaMemcpyDeviceToHost));
                ok = fwrite(buf.data(), 1, c, f.get()) == c;
            }
        };
        try {
            MB_CUDA(cudaStreamSynchronize(ix->stream));
            dump(ix->corpus, (size_t)ix->n_rows * ix->dim * sizeof(__half));
            dump(ix->doc_of_row, (size_t)ix->n_rows * sizeof(int32_t));
            // trailer: score-modifier attribute columns
            const int64_t trailer[3] = {ix->max_doc, ix->attr_cap, (int64_t)ix->attr_cols.size()};
            ok = ok && fwrite(trailer, sizeof(trailer), 1, f.get()) == 1;
            for (double* col : ix->attr_cols) {
                const int32_t present = col ? 1 : 0;
                ok = ok && fwrite(&present, sizeof(present), 1, f.get()) == 1;
                if (col) dump(col, (size_t)ix->attr_cap * sizeof(double));
            }
            ok = ok && fflush(f.get()) == 0;
        } catch (...) {
            f.reset();
            remove(tmp_path.c_str());
            throw;
        }
        FILE* raw = f.release();
        ok = (fclose(raw) == 0) && ok;
        if (!ok) {
            remove(tmp_path.c_str());
            fail(B200_ERR_INTERNAL, "short write to %s", tmp_path.c_str());
        }
        if (rename(tmp_path.c_str(), path) != 0) {
            remove(tmp_path.c_str());
            fail(B200_ERR_INTERNAL, "cannot rename %s to %s", tmp_path.c_str(), path);
        }
    });
}

int b200_index_load(int device, const char* path, b200_index** out) {
    return guarded([&] {
        MB_CHECK_ARG(path && out, "NULL argument");
        *out = nullptr;
        std::unique_ptr<FILE, FileCloser> f(fopen(path, "rb"));
        if (!f) fail(B200_ERR_INVALID_ARG, "cannot open %s", path);
        SnapshotHeader hdr;
        b200_index* ix = nullptr;
        try {
            if (fread(&hdr, sizeof(hdr), 1, f.get()) != 1 || memcmp(hdr.magic, "B200IDX\0", 8) != 0 ||
                (hdr.version != 1 && hdr.version != 2))
                fail(B200_ERR_INVALID_ARG, "%s is not a marqo_b200 index snapshot", path);
            ix = index_new(device, hdr.dim, hdr.metric, hdr.n_rows);
            DeviceGuard g(device);
            std::vector<uint8_t> buf((size_t)1 << 24);
            auto slurp = [&](void* dptr, size_t bytes) {
                for (size_t o = 0; o < bytes; o += buf.size()) {
                    const size_t c = std::min(buf.size(), bytes - o);
                    if (fread(buf.data(), 1, c, f.get()) != c) fail(B200_ERR_INVALID_ARG, "%s is truncated", path);
                    MB_CUDA(cudaMemcpy((uint8_t*)dptr + o, buf.data(), c, cudaMemcpyHostToDevice));
                }
            };
            slurp(ix->corpus, (size_t)hdr.n_rows * hdr.dim * sizeof(__half));
            slurp(ix->doc_of_row, (size_t)hdr.n_rows * sizeof(int32_t));
            ix->n_rows = hdr.n_rows;
            ix->has_docs = hdr.has_docs != 0;
            if (hdr.version >= 2) {
                int64_t trailer[3];
                if (fread(trailer, sizeof(trailer), 1, f.get()) != 1) fail(B200_ERR_INVALID_ARG, "%s is truncated", path);
                MB_CHECK_ARG(trailer[2] >= 0 && trailer[2] <= B200_MAX_ATTRIBUTE_COLUMNS && trailer[1] >= 0,
                             "%s has a corrupt attribute trailer", path);
                ix->max_doc = trailer[0];
                ix->attr_cap = trailer[1];
                ix->attr_cols.assign((size_t)trailer[2], nullptr);
                for (size_t c = 0; c < ix->attr_cols.size(); ++c) {
                    int32_t present = 0;
                    if (fread(&present, sizeof(present), 1, f.get()) != 1) fail(B200_ERR_INVALID_ARG, "%s is truncated", path);
                    if (!present) continue;
                    cuda_alloc((void**)&ix->attr_cols[c], (size_t)ix->attr_cap * sizeof(double));
                    slurp(ix->attr_cols[c], (size_t)ix->attr_cap * sizeof(double));
                }
                ix->cols_table_dirty = true;
            } else if (ix->has_docs && hdr.n_rows > 0) {
                track_max_doc(ix, ix->doc_of_row, hdr.n_rows);
            }
            if (hdr.n_rows > 0) {   // per-row norms (euclidean) and the largest norm (error bound of the scan)
                row_norms_kernel<<<(unsigned)((hdr.n_rows + 7) / 8), 256, 0, ix->stream>>>(
                    ix->corpus, hdr.n_rows, hdr.dim, ix->metric == B200_METRIC_EUCLIDEAN ? ix->row_n2 : nullptr, ix->max_n2);
                MB_CUDA(cudaGetLastError());
                MB_CUDA(cudaStreamSynchronize(ix->stream));
            }
        } catch (...) {
            index_free(ix);
            throw;
        }
        *out = ix;
    });
}

}  // extern "C"
