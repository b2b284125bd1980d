// Score + top-k over a GPU-resident fp16 embedding matrix.
//
// Reference semantics (executed inside Vespa today, specified by
// src/marqo/core/unstructured_vespa_index/unstructured_vespa_index.py:59-133 and the rank profile in
// src/marqo/core/unstructured_vespa_index/unstructured_vespa_schema.py:225-230,292-294):
//   score(doc) = max over the doc's chunk rows of closeness(q, row);  top-`hits` documents.
//
// Pipeline per group of <= 64 queries:
//   1. scan_kernel   persistent, one CTA per SM.  The corpus is streamed once from HBM by TMA (128-byte
//                    swizzle, 16 KB stages) and multiplied against the smem-resident query block with
//                    tcgen05.mma (M = 64 queries, N = 128 rows, fp16 x fp16 -> fp32 in TMEM).  Four epilogue
//                    warps read the accumulators back (TMEM lane == query) and each thread keeps a sorted
//                    register list of the KP best (score, row, doc) for its query, de-duplicated by document.
//   2. merge_kernel  one CTA per query: bitonic-sorts the per-CTA lists, de-duplicates documents, then
//                    RE-SCORES the KP survivors exactly (fp64, fixed summation order — the order
//                    oracle/score_oracle.c restates) and emits the top-k under (score desc, doc asc).
// The approximate tensor-core score only selects candidates; ids and scores returned are from the exact pass,
// which is what makes the ids bit-exact against the oracle.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
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
constexpr int KP = 16;          // candidates kept per list
constexpr int SLACK = 6;        // candidates beyond k that absorb approximate-vs-exact reordering at the boundary
constexpr int K_SINGLE = KP - SLACK;  // largest k answered by one pass over the corpus
constexpr int ACC_STAGES = 4;   // TMEM accumulators (4 x 128 fp32 columns = all 512 columns)
constexpr int THREADS = 192;    // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr uint32_t STAGE_BYTES = TILE_N * BLOCK_K * 2;
constexpr uint32_t QCHUNK_BYTES = MQ * BLOCK_K * 2;
constexpr int MAX_DIM = 1024;
constexpr int SMEM_LIMIT = 232448;  // 227 KB

struct ScanParams {
    int n_rows;
    int dim;
    int num_tiles;
    int num_stages;
    int nq;
    const int32_t* doc_of_row;  // used when HAS_DOCS
    const float* row_bias;      // used when HAS_BIAS: score = 2 * dot - row_bias[row]  (euclidean: |row|^2)
    const float* bound_score;   // optional [MQ]: only rows strictly after (bound_score, bound_row) qualify
    const int32_t* bound_row;
    const float2* mod;          // used when HAS_MOD: per-document (mult, add); key = mult * closeness + add
    const float* q_n2;          // used when HAS_MOD and euclidean: |q|^2 per query
    int metric;
    float* out_score;           // [grid][MQ][KP]
    int32_t* out_row;
    int32_t* out_doc;
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

// Sorted (score desc, arrival order) list insert with per-document de-duplication.
template <bool HAS_DOCS>
__device__ __forceinline__ void list_insert(float (&ls)[KP], int (&lr)[KP], int (&ld)[KP], float s, int row, int doc) {
    if (HAS_DOCS) {
        int pos = -1;
#pragma unroll
        for (int i = 0; i < KP; ++i)
            if (lr[i] >= 0 && ld[i] == doc) pos = i;
        if (pos >= 0) {
            float old = 0.f;
#pragma unroll
            for (int i = 0; i < KP; ++i)
                if (i == pos) old = ls[i];
            if (old >= s) return;  // the document is already listed with a better (or equal, earlier) chunk
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

template <bool HAS_DOCS, bool HAS_BIAS, bool HAS_MOD>
__global__ void __launch_bounds__(THREADS, 1)
scan_kernel(const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_q, ScanParams p) {
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
        // ------------------------------------------------------------ epilogue: per-query running top-KP
        const int sp = warp & 3;  // TMEM sub-partition this warp may read
        const int q = sp * 16 + lane;  // M = 64 accumulator: row m lives in lane (m % 16) of sub-partition m / 16
        const bool active = lane < 16 && q < p.nq;
        int32_t* my_docs = smem_docs + (warp - 2) * TILE_N;
        float* my_bias = smem_bias + (warp - 2) * TILE_N;
        float2* my_mod = smem_mod + (HAS_MOD ? (warp - 2) * TILE_N : 0);
        const float my_qn2 = (HAS_MOD && p.q_n2 != nullptr && active) ? p.q_n2[q] : 0.f;
        float ls[KP];
        int lr[KP], ld[KP];
#pragma unroll
        for (int i = 0; i < KP; ++i) {
            ls[i] = -INFINITY;
            lr[i] = -1;
            ld[i] = -1;
        }
        float thr = -INFINITY;
        const bool bounded = p.bound_score != nullptr;
        float bs = INFINITY;
        int br = -1;
        if (bounded && active) {
            bs = p.bound_score[q];
            br = p.bound_row[q];
        }
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
                    for (int j = 0; j < 32; ++j) mask |= (__uint_as_float(v[j]) > thr) ? (1u << j) : 0u;
                    const int nvalid = valid - c * 32;
                    if (nvalid < 32) mask &= nvalid <= 0 ? 0u : ((1u << nvalid) - 1u);
                    while (mask) {
                        const int j = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const float s = pick32(v, j);
                        const int row = row0 + c * 32 + j;
                        if (!(s > thr)) continue;
                        if (bounded && !(s < bs || (s == bs && row > br))) continue;
                        int doc = row;
                        if (HAS_DOCS) {
                            doc = my_docs[c * 32 + j];
                            if (doc < 0) continue;  // tombstoned row
                        }
                        list_insert<HAS_DOCS>(ls, lr, ld, s, row, doc);
                        thr = ls[KP - 1];
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
        if (lane < 16) {
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
struct MergeParams {
    int num_lists;  // scan grid size
    int nq;
    int k;
    int dim;
    int metric;
    int doc_offset;  // added to every returned document number (global numbering of a row-sharded corpus)
    int sort_n;  // power of two >= num_lists * KP
    const float* in_score;
    const int32_t* in_row;
    const int32_t* in_doc;
    const __half* qh;      // [MQ, dim] fp16 queries as scanned
    const __half* corpus;  // [n_rows, dim]
    int32_t* out_doc;      // [nq, k]
    int32_t* out_row;
    double* out_score;
    // optional candidate dump for multi-round search: [nq][KP] sorted by (approx desc, row asc)
    float* cand_score;
    int32_t* cand_row;
    // optional score modifiers: per-document (mult, add) in fp64; final key = mult * closeness + add
    const double2* mod64;
};

__device__ __forceinline__ bool approx_before(float sa, int ra, float sb, int rb) {
    return sa > sb || (sa == sb && ra < rb);
}

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

constexpr int MERGE_THREADS = 256;

__global__ void __launch_bounds__(MERGE_THREADS) merge_kernel(MergeParams p) {
    extern __shared__ uint8_t msmem[];
    float* s_score = reinterpret_cast<float*>(msmem);
    int32_t* s_row = reinterpret_cast<int32_t*>(s_score + p.sort_n);
    int32_t* s_doc = s_row + p.sort_n;
    __shared__ int sel_row[KP], sel_doc[KP], sel_n;
    __shared__ double sel_dot[KP];
    __shared__ float s_head[256];
    __shared__ float s_thr;
    __shared__ int s_count, s_valid;

    const int q = blockIdx.x;
    const int total = p.num_lists * KP;
    // Every list is sorted, so the KP-th largest list HEAD is a lower bound of the global KP-th best score:
    // only candidates >= that bound can be in the global top-KP.  This shrinks the sort from ~2.4k to ~KP..100 keys.
    for (int l = threadIdx.x; l < p.num_lists; l += blockDim.x) {
        const size_t src = ((size_t)l * MQ + q) * KP;
        s_head[l] = p.in_row[src] >= 0 ? p.in_score[src] : -INFINITY;
    }
    if (threadIdx.x == 0) s_thr = -INFINITY;
    __syncthreads();
    if (p.num_lists >= KP) {
        for (int l = threadIdx.x; l < p.num_lists; l += blockDim.x) {
            const float h = s_head[l];
            int rank = 0;
            for (int j = 0; j < p.num_lists; ++j) rank += (s_head[j] > h) || (s_head[j] == h && j < l);
            if (rank == KP - 1) s_thr = h;
        }
    }
    __syncthreads();
    for (int attempt = 0; attempt < 2; ++attempt) {
        const float thr = attempt == 0 ? s_thr : -INFINITY;
        if (threadIdx.x == 0) {
            s_count = 0;
            s_valid = 0;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int list = i / KP, e = i % KP;
            const size_t src = ((size_t)list * MQ + q) * KP + e;
            const int rr = p.in_row[src];
            if (rr < 0) continue;
            atomicAdd(&s_valid, 1);
            const float sc = p.in_score[src];
            if (sc >= thr) {
                const int slot = atomicAdd(&s_count, 1);
                s_score[slot] = sc;
                s_row[slot] = rr;
                s_doc[slot] = p.in_doc[src];
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
        // bitonic sort, order: (score desc, row asc) — a total order, so the result does not depend on slot order
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
        if (threadIdx.x == 0) {
            int n = 0;
            for (int i = 0; i < count && n < KP; ++i) {
                const int d = s_doc[i];
                bool dup = false;
                for (int j = 0; j < n; ++j) dup |= (sel_doc[j] == d);
                if (dup) continue;
                sel_row[n] = s_row[i];
                sel_doc[n] = d;
                if (p.cand_score) {
                    p.cand_score[(size_t)q * KP + n] = s_score[i];
                    p.cand_row[(size_t)q * KP + n] = s_row[i];
                }
                ++n;
            }
            if (p.cand_score)
                for (int i = n; i < KP; ++i) {
                    p.cand_score[(size_t)q * KP + i] = -INFINITY;
                    p.cand_row[(size_t)q * KP + i] = -1;
                }
            sel_n = n;
        }
        __syncthreads();
        // duplicates of one document across lists can leave fewer than KP distinct documents above the bound:
        // fall back to the unfiltered merge
        if (sel_n >= KP || s_count >= s_valid) break;
        __syncthreads();
    }
    // exact re-score: lane l accumulates elements i = 32 j + l in ascending j (fp64), partials are then added in
    // ascending lane order.  oracle/score_oracle.c restates exactly this order.
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = warp; c < sel_n; c += MERGE_THREADS / 32) {
        const __half* qv = p.qh + (size_t)q * p.dim;
        const __half* cv = p.corpus + (size_t)sel_row[c] * p.dim;
        double part = 0.0;
        if (p.metric == B200_METRIC_EUCLIDEAN) {
            for (int i = lane; i < p.dim; i += 32) {
                const double d = (double)__half2float(qv[i]) - (double)__half2float(cv[i]);   // exact in fp64
                part -= d * d;
            }
        } else {
            for (int i = lane; i < p.dim; i += 32) part += (double)__half2float(qv[i]) * (double)__half2float(cv[i]);
        }
        double tot = 0.0;
        for (int l = 0; l < 32; ++l) tot += __shfl_sync(0xffffffffu, part, l);
        if (lane == 0) {
            if (p.mod64) {   // the ordering key becomes the modified score (separate multiply and add, no fma)
                const double2 ma = p.mod64[sel_doc[c]];
                tot = __dadd_rn(__dmul_rn(ma.x, closeness_from_dot(tot, p.metric)), ma.y);
            }
            sel_dot[c] = tot;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int n = sel_n;
        // insertion sort by (key desc, doc asc); key = exact dot product, or the modified score
        for (int i = 1; i < n; ++i) {
            const double d = sel_dot[i];
            const int r = sel_row[i], dc = sel_doc[i];
            int j = i - 1;
            while (j >= 0 && (sel_dot[j] < d || (sel_dot[j] == d && sel_doc[j] > dc))) {
                sel_dot[j + 1] = sel_dot[j];
                sel_row[j + 1] = sel_row[j];
                sel_doc[j + 1] = sel_doc[j];
                --j;
            }
            sel_dot[j + 1] = d;
            sel_row[j + 1] = r;
            sel_doc[j + 1] = dc;
        }
        for (int i = 0; i < p.k; ++i) {
            const size_t o = (size_t)q * p.k + i;
            if (i < n) {
                p.out_doc[o] = sel_doc[i] + p.doc_offset;
                p.out_row[o] = sel_row[i];
                p.out_score[o] = p.mod64 ? sel_dot[i] : closeness_from_dot(sel_dot[i], p.metric);
            } else {
                p.out_doc[o] = -1;
                p.out_row[o] = -1;
                p.out_score[o] = -INFINITY;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// fp32 -> fp16 row conversion (optionally L2-normalising first, for the angular metric).
__global__ void convert_rows_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t rows, int dim,
                                    int64_t dst_rows_total, int normalize, float* __restrict__ n2_out = nullptr) {
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
    for (int i = lane; i < dim; i += 32) {
        const __half hv = __float2half_rn(s[i] * scale);
        d[i] = hv;
        const float f = __half2float(hv);
        n2 = fmaf(f, f, n2);
    }
    if (n2_out) {   // |row|^2 of the STORED (fp16-rounded) values: the euclidean scan's per-row term
        for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
        if (lane == 0) n2_out[row] = n2;
    }
}

__global__ void row_norms_kernel(const __half* __restrict__ rows, int64_t n, int dim, float* __restrict__ n2_out) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    float n2 = 0.f;
    for (int i = lane; i < dim; i += 32) {
        const float f = __half2float(rows[row * dim + i]);
        n2 = fmaf(f, f, n2);
    }
    for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    if (lane == 0) n2_out[row] = n2;
}

__global__ void iota_kernel(int32_t* dst, int64_t n, int32_t start) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = start + (int32_t)i;
}

__global__ void tombstone_kernel(int32_t* doc_of_row, int64_t n, int32_t doc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && doc_of_row[i] == doc) doc_of_row[i] = -1;
}

// bound for the next round = the last kept candidate of this round (approximate order); a short list means the
// corpus is exhausted for that query: -inf makes every later comparison fail.
__global__ void next_bound_kernel(const float* cand_score, const int32_t* cand_row, float* bound_score,
                                  int32_t* bound_row) {
    const int q = threadIdx.x;
    const int r = cand_row[q * KP + KP - 1];
    bound_score[q] = r >= 0 ? cand_score[q * KP + KP - 1] : -INFINITY;
    bound_row[q] = r >= 0 ? r : INT_MAX;
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
}

__global__ void fill_nan_kernel(double* dst, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = NAN;
}

__global__ void scatter_attr_kernel(double* col, const int32_t* docs, const double* vals, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) col[docs[i]] = vals ? vals[i] : NAN;
}

__global__ void query_norms_kernel(const __half* __restrict__ qh, int dim, float* __restrict__ out) {
    // one warp per query row of the [MQ, dim] fp16 block
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= MQ) return;
    float n2 = 0.f;
    for (int i = lane; i < dim; i += 32) {
        const float f = __half2float(qh[(size_t)q * dim + i]);
        n2 = fmaf(f, f, n2);
    }
    for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    if (lane == 0) out[q] = n2;
}

// Merge of all-gathered per-shard lists on the device: one warp per query, candidates strided over the lanes,
// k rounds of warp arg-max under (score desc, doc asc).  Shard s's block: doc int32 [nq,k] | row int32 [nq,k] |
// score f64 [nq,k] packed back to back (the layout b200_index_search_device writes when given one buffer).
__global__ void merge_shards_kernel(const uint8_t* __restrict__ gathered, int nshards, int nq, int k, int32_t* out_doc,
                                    int32_t* out_row, double* out_score) {
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    const size_t blk = (size_t)nq * k * 16;
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
            const uint8_t* base = gathered + (size_t)s * blk;
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

struct b200_index {
    int device = 0;
    int dim = 0;
    int metric = 0;
    int sms = 0;
    int64_t capacity = 0;
    int64_t n_rows = 0;
    bool has_docs = false;  // false while doc_of_row[i] == i for every row (identity fast path)
    int32_t doc_offset = 0; // added to returned document numbers (shard -> global numbering)
    __half* corpus = nullptr;
    int32_t* doc_of_row = nullptr;
    float* row_n2 = nullptr;       // [capacity] squared norms (euclidean metric only)
    // per-search workspaces
    __half* qh = nullptr;          // [MQ, dim]
    float* q_stage = nullptr;      // [MQ, dim] fp32 staging for host queries
    float* list_score = nullptr;   // [sms][MQ][KP]
    int32_t* list_row = nullptr;
    int32_t* list_doc = nullptr;
    int32_t* o_doc = nullptr;      // [MQ, KP] device outputs for the host API
    int32_t* o_row = nullptr;
    double* o_score = nullptr;
    float *cand_score = nullptr, *bound_score = nullptr;   // [MQ, KP] / [MQ]: multi-round search (k > K_SINGLE)
    int32_t *cand_row = nullptr, *bound_row = nullptr;
    // score modifiers: per-document numeric attributes (one device column per attribute name, NaN = missing)
    std::vector<double*> attr_cols;
    int64_t attr_cap = 0;          // documents each column can hold
    int64_t max_doc = -1;          // largest explicit document number seen by add()
    double2* mod64 = nullptr;      // [mod_cap] (mult, add) of the current modified search
    float2* mod32 = nullptr;
    int64_t mod_cap = 0;
    float* q_n2 = nullptr;         // [MQ] |q|^2 (euclidean + modifiers)
    int* d_flag = nullptr;
    bool mod_active = false;
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
    cudaFree(ix->qh);
    cudaFree(ix->q_stage);
    cudaFree(ix->list_score);
    cudaFree(ix->list_row);
    cudaFree(ix->list_doc);
    cudaFree(ix->o_doc);
    cudaFree(ix->o_row);
    cudaFree(ix->o_score);
    cudaFree(ix->cand_score);
    cudaFree(ix->cand_row);
    cudaFree(ix->bound_score);
    cudaFree(ix->bound_row);
    for (double* c : ix->attr_cols) cudaFree(c);
    cudaFree(ix->mod64);
    cudaFree(ix->mod32);
    cudaFree(ix->q_n2);
    cudaFree(ix->d_flag);
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

void ensure_capacity(b200_index* ix, int64_t need_rows) {
    if (need_rows <= ix->capacity) return;
    int64_t cap = std::max<int64_t>(need_rows, ix->capacity + ix->capacity / 2);
    cap = (int64_t)round_up((size_t)cap, TILE_N);
    __half* nc = nullptr;
    int32_t* nd = nullptr;
    cuda_alloc((void**)&nc, (size_t)cap * ix->dim * sizeof(__half));
    cuda_alloc((void**)&nd, (size_t)cap * sizeof(int32_t));
    if (ix->n_rows > 0) {
        MB_CUDA(cudaMemcpyAsync(nc, ix->corpus, (size_t)ix->n_rows * ix->dim * sizeof(__half), cudaMemcpyDeviceToDevice,
                                ix->stream));
        MB_CUDA(cudaMemcpyAsync(nd, ix->doc_of_row, (size_t)ix->n_rows * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                                ix->stream));
    }
    float* nn = nullptr;
    if (ix->metric == B200_METRIC_EUCLIDEAN) {
        cuda_alloc((void**)&nn, (size_t)cap * sizeof(float));
        if (ix->n_rows > 0)
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
        ix->sms = sm_count(device);
        MB_CUDA(cudaStreamCreateWithFlags(&ix->own_stream, cudaStreamNonBlocking));
        ix->stream = ix->own_stream;
        for (auto& e : ix->ev) MB_CUDA(cudaEventCreate(&e));
        cuda_alloc((void**)&ix->qh, (size_t)MQ * dim * sizeof(__half));
        cuda_alloc((void**)&ix->q_stage, (size_t)MQ * dim * sizeof(float));
        const size_t nl = (size_t)ix->sms * MQ * KP;
        cuda_alloc((void**)&ix->list_score, nl * sizeof(float));
        cuda_alloc((void**)&ix->list_row, nl * sizeof(int32_t));
        cuda_alloc((void**)&ix->list_doc, nl * sizeof(int32_t));
        cuda_alloc((void**)&ix->o_doc, (size_t)MQ * KP * sizeof(int32_t));
        cuda_alloc((void**)&ix->o_row, (size_t)MQ * KP * sizeof(int32_t));
        cuda_alloc((void**)&ix->o_score, (size_t)MQ * KP * sizeof(double));
        cuda_alloc((void**)&ix->cand_score, (size_t)MQ * KP * sizeof(float));
        cuda_alloc((void**)&ix->cand_row, (size_t)MQ * KP * sizeof(int32_t));
        cuda_alloc((void**)&ix->bound_score, (size_t)MQ * sizeof(float));
        cuda_alloc((void**)&ix->bound_row, (size_t)MQ * sizeof(int32_t));
        cuda_alloc((void**)&ix->q_n2, (size_t)MQ * sizeof(float));
        cuda_alloc((void**)&ix->d_flag, sizeof(int));
        ensure_capacity(ix, std::max<int64_t>(capacity_rows, TILE_N));
        const auto smem_attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
        MB_CUDA(cudaFuncSetAttribute(scan_kernel<false, false, false>, smem_attr, SMEM_LIMIT));
        MB_CUDA(cudaFuncSetAttribute(scan_kernel<true, false, false>, smem_attr, SMEM_LIMIT));
        MB_CUDA(cudaFuncSetAttribute(scan_kernel<false, true, false>, smem_attr, SMEM_LIMIT));
        MB_CUDA(cudaFuncSetAttribute(scan_kernel<true, true, false>, smem_attr, SMEM_LIMIT));
        MB_CUDA(cudaFuncSetAttribute(scan_kernel<false, false, true>, smem_attr, SMEM_LIMIT));
        MB_CUDA(cudaFuncSetAttribute(scan_kernel<true, false, true>, smem_attr, SMEM_LIMIT));
        MB_CUDA(cudaFuncSetAttribute(scan_kernel<false, true, true>, smem_attr, SMEM_LIMIT));
        MB_CUDA(cudaFuncSetAttribute(scan_kernel<true, true, true>, smem_attr, SMEM_LIMIT));
        MB_CUDA(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    } catch (...) {
        index_free(ix);
        throw;
    }
    return ix;
}

void add_rows_device(b200_index* ix, const float* d_vecs, const int32_t* d_doc_ids, int64_t m) {
    ensure_capacity(ix, ix->n_rows + m);
    const int wpb = 8;
    const int64_t blocks = (m + wpb - 1) / wpb;
    convert_rows_kernel<<<(unsigned)blocks, wpb * 32, 0, ix->stream>>>(
        d_vecs, ix->corpus + (size_t)ix->n_rows * ix->dim, m, ix->dim, m, ix->metric == B200_METRIC_ANGULAR,
        ix->metric == B200_METRIC_EUCLIDEAN ? ix->row_n2 + ix->n_rows : nullptr);
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

// One pass over the corpus for <= MQ queries already converted into ix->qh.
void search_group(b200_index* ix, int nq, int k, int32_t* d_out_doc, int32_t* d_out_row, double* d_out_score,
                  bool record_timing, const float* bound_score = nullptr, const int32_t* bound_row = nullptr,
                  float* cand_score = nullptr, int32_t* cand_row = nullptr) {
    const int total = nq * k;
    if (ix->n_rows == 0) {
        fill_empty_kernel<<<(total + 255) / 256, 256, 0, ix->stream>>>(d_out_doc, d_out_row, d_out_score, total);
        MB_CUDA(cudaGetLastError());
        return;
    }
    const int num_tiles = (int)((ix->n_rows + TILE_N - 1) / TILE_N);
    const int grid = std::min(num_tiles, ix->sms);
    const bool mod = ix->mod_active;
    int stages = 16;
    while (stages > 2 && scan_smem_bytes(ix->dim, stages, mod) > (size_t)SMEM_LIMIT) --stages;
    const size_t smem = scan_smem_bytes(ix->dim, stages, mod);
    if (smem > (size_t)SMEM_LIMIT) fail(B200_ERR_INTERNAL, "scan kernel shared memory budget exceeded");

    CUtensorMap tmap_c = make_tmap_2d(ix->corpus, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)ix->dim,
                                      (uint64_t)ix->n_rows, (uint64_t)ix->dim * 2, BLOCK_K, TILE_N,
                                      CU_TENSOR_MAP_SWIZZLE_128B);
    CUtensorMap tmap_q = make_tmap_2d(ix->qh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)ix->dim, (uint64_t)MQ,
                                      (uint64_t)ix->dim * 2, BLOCK_K, MQ, CU_TENSOR_MAP_SWIZZLE_128B);
    ScanParams sp{};
    sp.n_rows = (int)ix->n_rows;
    sp.dim = ix->dim;
    sp.num_tiles = num_tiles;
    sp.num_stages = stages;
    sp.nq = nq;
    sp.doc_of_row = ix->doc_of_row;
    sp.row_bias = ix->row_n2;
    sp.bound_score = bound_score;
    sp.bound_row = bound_row;
    sp.out_score = ix->list_score;
    sp.out_row = ix->list_row;
    sp.out_doc = ix->list_doc;
    sp.mod = mod ? ix->mod32 : nullptr;
    sp.q_n2 = (mod && ix->metric == B200_METRIC_EUCLIDEAN) ? ix->q_n2 : nullptr;
    sp.metric = ix->metric;

    if (record_timing) MB_CUDA(cudaEventRecord(ix->ev[0], ix->stream));
    const bool bias = ix->metric == B200_METRIC_EUCLIDEAN;
    using ScanFn = void (*)(const CUtensorMap, const CUtensorMap, ScanParams);
    static const ScanFn table[8] = {scan_kernel<false, false, false>, scan_kernel<true, false, false>,
                                    scan_kernel<false, true, false>,  scan_kernel<true, true, false>,
                                    scan_kernel<false, false, true>,  scan_kernel<true, false, true>,
                                    scan_kernel<false, true, true>,   scan_kernel<true, true, true>};
    table[(ix->has_docs ? 1 : 0) | (bias ? 2 : 0) | (mod ? 4 : 0)]<<<grid, THREADS, smem, ix->stream>>>(tmap_c, tmap_q, sp);
    MB_CUDA(cudaGetLastError());
    if (record_timing) MB_CUDA(cudaEventRecord(ix->ev[1], ix->stream));

    MergeParams mp{};
    mp.num_lists = grid;
    mp.nq = nq;
    mp.k = k;
    mp.dim = ix->dim;
    mp.metric = ix->metric;
    mp.doc_offset = ix->doc_offset;
    int sort_n = 32;
    while (sort_n < grid * KP) sort_n <<= 1;
    mp.sort_n = sort_n;
    mp.in_score = ix->list_score;
    mp.in_row = ix->list_row;
    mp.in_doc = ix->list_doc;
    mp.qh = ix->qh;
    mp.corpus = ix->corpus;
    mp.out_doc = d_out_doc;
    mp.out_row = d_out_row;
    mp.out_score = d_out_score;
    mp.cand_score = cand_score;
    mp.cand_row = cand_row;
    mp.mod64 = mod ? ix->mod64 : nullptr;
    merge_kernel<<<nq, MERGE_THREADS, (size_t)sort_n * 12, ix->stream>>>(mp);
    MB_CUDA(cudaGetLastError());
    if (record_timing) {
        MB_CUDA(cudaEventRecord(ix->ev[2], ix->stream));
        ix->timing_valid = true;
    }
}

// k beyond the single-pass limit: repeated scans, each restricted to rows strictly after the previous round's last
// candidate in the (approximate score desc, row asc) order, until k + SLACK distinct documents are collected.
// Every round re-scores its candidates exactly, so the final order is again (exact score desc, doc asc).
void search_group_rounds(b200_index* ix, int nq, int k, int32_t* d_out_doc, int32_t* d_out_row, double* d_out_score) {
    struct Hit {
        double s;
        int32_t doc, row;
    };
    std::vector<std::vector<Hit>> acc(nq);
    std::vector<char> done(nq, 0);
    std::vector<int32_t> h_doc((size_t)MQ * KP), h_row((size_t)MQ * KP);
    std::vector<double> h_score((size_t)MQ * KP);
    std::vector<int32_t> seen;
    const int max_rounds = 20000;
    for (int round = 0;; ++round) {
        if (round >= max_rounds) fail(B200_ERR_INTERNAL, "search did not converge in %d rounds", max_rounds);
        search_group(ix, nq, KP, ix->o_doc, ix->o_row, ix->o_score, round == 0, round == 0 ? nullptr : ix->bound_score,
                     round == 0 ? nullptr : ix->bound_row, ix->cand_score, ix->cand_row);
        if (ix->n_rows == 0) break;
        next_bound_kernel<<<1, MQ, 0, ix->stream>>>(ix->cand_score, ix->cand_row, ix->bound_score, ix->bound_row);
        MB_CUDA(cudaGetLastError());
        MB_CUDA(cudaMemcpyAsync(h_doc.data(), ix->o_doc, (size_t)nq * KP * 4, cudaMemcpyDeviceToHost, ix->stream));
        MB_CUDA(cudaMemcpyAsync(h_row.data(), ix->o_row, (size_t)nq * KP * 4, cudaMemcpyDeviceToHost, ix->stream));
        MB_CUDA(cudaMemcpyAsync(h_score.data(), ix->o_score, (size_t)nq * KP * 8, cudaMemcpyDeviceToHost, ix->stream));
        MB_CUDA(cudaStreamSynchronize(ix->stream));
        bool all_done = true;
        for (int q = 0; q < nq; ++q) {
            if (done[q]) continue;
            int valid = 0;
            for (int i = 0; i < KP; ++i) {
                const size_t o = (size_t)q * KP + i;
                if (h_doc[o] < 0) continue;
                ++valid;
                acc[q].push_back({h_score[o], h_doc[o], h_row[o]});
            }
            seen.clear();
            for (const Hit& h : acc[q]) seen.push_back(h.doc);
            std::sort(seen.begin(), seen.end());
            const int distinct = (int)(std::unique(seen.begin(), seen.end()) - seen.begin());
            if (valid < KP || distinct >= k + SLACK) done[q] = 1;
            all_done = all_done && done[q];
        }
        if (all_done) break;
    }
    std::vector<int32_t> o_doc((size_t)nq * k, -1), o_row((size_t)nq * k, -1);
    std::vector<double> o_score((size_t)nq * k, -std::numeric_limits<double>::infinity());
    for (int q = 0; q < nq; ++q) {
        std::vector<Hit>& v = acc[q];
        // best chunk per document: (score desc, row asc), then documents by (score desc, doc asc)
        std::sort(v.begin(), v.end(), [](const Hit& a, const Hit& b) {
            return a.doc < b.doc || (a.doc == b.doc && (a.s > b.s || (a.s == b.s && a.row < b.row)));
        });
        size_t w = 0;
        for (size_t i = 0; i < v.size(); ++i)
            if (i == 0 || v[i].doc != v[i - 1].doc) v[w++] = v[i];
        v.resize(w);
        std::sort(v.begin(), v.end(), [](const Hit& a, const Hit& b) { return a.s > b.s || (a.s == b.s && a.doc < b.doc); });
        for (int i = 0; i < k && i < (int)v.size(); ++i) {
            o_doc[(size_t)q * k + i] = v[i].doc;
            o_row[(size_t)q * k + i] = v[i].row;
            o_score[(size_t)q * k + i] = v[i].s;
        }
    }
    MB_CUDA(cudaMemcpyAsync(d_out_doc, o_doc.data(), o_doc.size() * 4, cudaMemcpyHostToDevice, ix->stream));
    MB_CUDA(cudaMemcpyAsync(d_out_row, o_row.data(), o_row.size() * 4, cudaMemcpyHostToDevice, ix->stream));
    MB_CUDA(cudaMemcpyAsync(d_out_score, o_score.data(), o_score.size() * 8, cudaMemcpyHostToDevice, ix->stream));
    MB_CUDA(cudaStreamSynchronize(ix->stream));
}

void search_device(b200_index* ix, const float* d_q, int nq, int k, int32_t* d_out_doc, int32_t* d_out_row,
                   double* d_out_score) {
    for (int q0 = 0; q0 < nq; q0 += MQ) {
        const int g = std::min(MQ, nq - q0);
        convert_rows_kernel<<<MQ / 8, 256, 0, ix->stream>>>(d_q + (size_t)q0 * ix->dim, ix->qh, g, ix->dim, MQ,
                                                            ix->metric == B200_METRIC_ANGULAR);
        MB_CUDA(cudaGetLastError());
        if (ix->mod_active && ix->metric == B200_METRIC_EUCLIDEAN) {
            query_norms_kernel<<<MQ / 8, 256, 0, ix->stream>>>(ix->qh, ix->dim, ix->q_n2);
            MB_CUDA(cudaGetLastError());
        }
        if (k <= K_SINGLE)
            search_group(ix, g, k, d_out_doc + (size_t)q0 * k, d_out_row + (size_t)q0 * k, d_out_score + (size_t)q0 * k,
                         q0 + MQ >= nq);
        else
            search_group_rounds(ix, g, k, d_out_doc + (size_t)q0 * k, d_out_row + (size_t)q0 * k,
                                d_out_score + (size_t)q0 * k);
    }
}

void check_search_args(b200_index* ix, const void* q, int nq, int k, const void* a, const void* b, const void* c) {
    MB_CHECK_ARG(ix != nullptr, "index is NULL");
    MB_CHECK_ARG(q && a && b && c, "NULL buffer");
    MB_CHECK_ARG(nq > 0, "nq must be positive (got %d)", nq);
    MB_CHECK_ARG(k > 0, "k must be positive (got %d)", k);
    MB_CHECK_ARG(k <= 10000, "k = %d exceeds 10000 (Marqo's own limit + offset cap, tensor_search.py:1568-1588)", k);
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
}

double* attr_column(b200_index* ix, int column) {
    if ((int)ix->attr_cols.size() <= column) ix->attr_cols.resize(column + 1, nullptr);
    if (!ix->attr_cols[column]) {
        cuda_alloc((void**)&ix->attr_cols[column], (size_t)ix->attr_cap * sizeof(double));
        fill_nan(ix, ix->attr_cols[column], ix->attr_cap);
    }
    return ix->attr_cols[column];
}

struct ModScope {  // marks the index as "searching with modifiers" for the duration of one call
    b200_index* ix;
    explicit ModScope(b200_index* i) : ix(i) { ix->mod_active = true; }
    ~ModScope() { ix->mod_active = false; }
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
    mp.negative_flag = ix->d_flag;
    MB_CUDA(cudaMemsetAsync(ix->d_flag, 0, sizeof(int), ix->stream));
    modifier_kernel<<<(unsigned)((nd + 255) / 256), 256, 0, ix->stream>>>(mp);
    MB_CUDA(cudaGetLastError());
    int flag = 0;
    MB_CUDA(cudaMemcpyAsync(&flag, ix->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ix->stream));
    MB_CUDA(cudaStreamSynchronize(ix->stream));
    // closeness(field, embeddings) is the best chunk's closeness; the scan keeps, per document, the chunk with the best
    // MODIFIED key, which is the same chunk only while the multiplier is >= 0.
    if (flag && ix->has_docs)
        fail(B200_ERR_UNSUPPORTED,
             "a negative multiplicative score modifier on a corpus with explicit document ids (multi-chunk documents) "
             "is not supported");
}

void search_host(b200_index* ix, const float* q, int nq, int k, int32_t* out_doc, int32_t* out_row, double* out_score) {
    int32_t *dd = ix->o_doc, *dr = ix->o_row;
    double* ds = ix->o_score;
    void* tmp[3] = {nullptr, nullptr, nullptr};
    if (k > KP) {  // the resident [MQ, KP] output block is too small for a large k
        cuda_alloc(&tmp[0], (size_t)MQ * k * sizeof(int32_t));
        cuda_alloc(&tmp[1], (size_t)MQ * k * sizeof(int32_t));
        cuda_alloc(&tmp[2], (size_t)MQ * k * sizeof(double));
        dd = (int32_t*)tmp[0];
        dr = (int32_t*)tmp[1];
        ds = (double*)tmp[2];
    }
    try {
        for (int q0 = 0; q0 < nq; q0 += MQ) {
            const int gq = std::min(MQ, nq - q0);
            MB_CUDA(cudaMemcpyAsync(ix->q_stage, q + (size_t)q0 * ix->dim, (size_t)gq * ix->dim * sizeof(float),
                                    cudaMemcpyHostToDevice, ix->stream));
            search_device(ix, ix->q_stage, gq, k, dd, dr, ds);
            MB_CUDA(cudaMemcpyAsync(out_doc + (size_t)q0 * k, dd, (size_t)gq * k * sizeof(int32_t),
                                    cudaMemcpyDeviceToHost, ix->stream));
            MB_CUDA(cudaMemcpyAsync(out_row + (size_t)q0 * k, dr, (size_t)gq * k * sizeof(int32_t),
                                    cudaMemcpyDeviceToHost, ix->stream));
            MB_CUDA(cudaMemcpyAsync(out_score + (size_t)q0 * k, ds, (size_t)gq * k * sizeof(double),
                                    cudaMemcpyDeviceToHost, ix->stream));
            MB_CUDA(cudaStreamSynchronize(ix->stream));
        }
    } catch (...) {
        for (void* t : tmp) cudaFree(t);
        throw;
    }
    for (void* t : tmp) cudaFree(t);
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
        if (doc_ids)
            for (int64_t i = 0; i < m; ++i) {
                MB_CHECK_ARG(doc_ids[i] >= 0, "doc_ids[%lld] is negative", (long long)i);
                ix->max_doc = std::max<int64_t>(ix->max_doc, doc_ids[i]);
            }
        // When explicit ids are given but the index has been identity-mapped so far, the ids must be honoured.
        float* d_v = nullptr;
        int32_t* d_d = nullptr;
        const int64_t chunk = 1 << 16;
        cuda_alloc((void**)&d_v, (size_t)std::min(m, chunk) * ix->dim * sizeof(float));
        if (doc_ids) cuda_alloc((void**)&d_d, (size_t)std::min(m, chunk) * sizeof(int32_t));
        try {
            for (int64_t o = 0; o < m; o += chunk) {
                const int64_t c = std::min(chunk, m - o);
                MB_CUDA(cudaMemcpyAsync(d_v, vecs + (size_t)o * ix->dim, (size_t)c * ix->dim * sizeof(float),
                                        cudaMemcpyHostToDevice, ix->stream));
                if (doc_ids)
                    MB_CUDA(cudaMemcpyAsync(d_d, doc_ids + o, (size_t)c * sizeof(int32_t), cudaMemcpyHostToDevice,
                                            ix->stream));
                add_rows_device(ix, d_v, doc_ids ? d_d : nullptr, c);
                MB_CUDA(cudaStreamSynchronize(ix->stream));
            }
        } catch (...) {
            cudaFree(d_v);
            cudaFree(d_d);
            throw;
        }
        cudaFree(d_v);
        cudaFree(d_d);
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
        add_rows_device(ix, d_vecs, d_doc_ids, m);
        MB_CUDA(cudaStreamSynchronize(ix->stream));
        if (d_doc_ids) track_max_doc(ix, d_doc_ids, m);
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
    return guarded([&] {
        MB_CHECK_ARG(ix && out_vec, "NULL argument");
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        MB_CHECK_ARG(row >= 0 && row < ix->n_rows, "row %lld out of range", (long long)row);
        std::vector<__half> tmp(ix->dim);
        MB_CUDA(cudaMemcpyAsync(tmp.data(), ix->corpus + (size_t)row * ix->dim, (size_t)ix->dim * sizeof(__half),
                                cudaMemcpyDeviceToHost, ix->stream));
        MB_CUDA(cudaStreamSynchronize(ix->stream));
        for (int i = 0; i < ix->dim; ++i) out_vec[i] = __half2float(tmp[i]);
    });
}

int b200_index_search(b200_index* ix, const float* q, int nq, int k, int32_t* out_doc, int32_t* out_row,
                      double* out_score) {
    return guarded([&] {
        check_search_args(ix, q, nq, k, out_doc, out_row, out_score);
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        search_host(ix, q, nq, k, out_doc, out_row, out_score);
    });
}

int b200_index_search_modified(b200_index* ix, const float* q, int nq, int k, const int32_t* mult_cols,
                               const double* mult_w, int n_mult, const int32_t* add_cols, const double* add_w, int n_add,
                               int32_t* out_doc, int32_t* out_row, double* out_score) {
    return guarded([&] {
        check_search_args(ix, q, nq, k, out_doc, out_row, out_score);
        MB_CHECK_ARG(n_mult >= 0 && n_mult <= MAX_MOD_TERMS && n_add >= 0 && n_add <= MAX_MOD_TERMS,
                     "at most %d multiplicative and %d additive modifiers per search", MAX_MOD_TERMS, MAX_MOD_TERMS);
        MB_CHECK_ARG((n_mult == 0 || (mult_cols && mult_w)) && (n_add == 0 || (add_cols && add_w)), "NULL modifier list");
        for (int i = 0; i < n_mult; ++i) MB_CHECK_ARG(std::isfinite(mult_w[i]), "mult_w[%d] is not finite", i);
        for (int i = 0; i < n_add; ++i) MB_CHECK_ARG(std::isfinite(add_w[i]), "add_w[%d] is not finite", i);
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        prepare_modifiers(ix, mult_cols, mult_w, n_mult, add_cols, add_w, n_add);
        ModScope scope(ix);
        search_host(ix, q, nq, k, out_doc, out_row, out_score);
    });
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
        ensure_attr_capacity(ix, hi + 1);
        int32_t* d_ids = nullptr;
        double* d_vals = nullptr;
        cuda_alloc((void**)&d_ids, (size_t)n * sizeof(int32_t));
        try {
            if (values) cuda_alloc((void**)&d_vals, (size_t)n * sizeof(double));
            MB_CUDA(cudaMemcpyAsync(d_ids, doc_ids, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ix->stream));
            if (values)
                MB_CUDA(cudaMemcpyAsync(d_vals, values, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ix->stream));
            const unsigned blocks = (unsigned)((n + 255) / 256);
            if (column >= 0) {
                scatter_attr_kernel<<<blocks, 256, 0, ix->stream>>>(attr_column(ix, column), d_ids, d_vals, n);
                MB_CUDA(cudaGetLastError());
            } else {
                for (double* col : ix->attr_cols)
                    if (col) {
                        scatter_attr_kernel<<<blocks, 256, 0, ix->stream>>>(col, d_ids, nullptr, n);
                        MB_CUDA(cudaGetLastError());
                    }
            }
            MB_CUDA(cudaStreamSynchronize(ix->stream));
        } catch (...) {
            cudaFree(d_ids);
            cudaFree(d_vals);
            throw;
        }
        cudaFree(d_ids);
        cudaFree(d_vals);
    });
}

int b200_index_search_device(b200_index* ix, const float* d_q, int nq, int k, int32_t* d_out_doc, int32_t* d_out_row,
                             double* d_out_score, int sync) {
    return guarded([&] {
        check_search_args(ix, d_q, nq, k, d_out_doc, d_out_row, d_out_score);
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        search_device(ix, d_q, nq, k, d_out_doc, d_out_row, d_out_score);
        if (sync) MB_CUDA(cudaStreamSynchronize(ix->stream));
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

int b200_index_save(b200_index* ix, const char* path) {
    return guarded([&] {
        MB_CHECK_ARG(ix && path, "NULL argument");
        std::lock_guard<std::mutex> lk(ix->mu);
        DeviceGuard g(ix->device);
        FILE* f = fopen(path, "wb");
        if (!f) fail(B200_ERR_INVALID_ARG, "cannot open %s for writing", path);
        struct {
            char magic[8];
            int32_t version, dim, metric, has_docs;
            int64_t n_rows;
        } hdr;
        memcpy(hdr.magic, "B200IDX\0", 8);
        hdr.version = 2;   // 2 = version 1 + the attribute-column trailer
        hdr.dim = ix->dim;
        hdr.metric = ix->metric;
        hdr.has_docs = ix->has_docs ? 1 : 0;
        hdr.n_rows = ix->n_rows;
        bool ok = fwrite(&hdr, sizeof(hdr), 1, f) == 1;
        std::vector<uint8_t> buf((size_t)1 << 24);
        auto dump = [&](const void* dptr, size_t bytes) {
            for (size_t o = 0; o < bytes && ok; o += buf.size()) {
                const size_t c = std::min(buf.size(), bytes - o);
                MB_CUDA(cudaMemcpy(buf.data(), (const uint8_t*)dptr + o, c, cudaMemcpyDeviceToHost));
                ok = fwrite(buf.data(), 1, c, f) == c;
            }
        };
        MB_CUDA(cudaStreamSynchronize(ix->stream));
        dump(ix->corpus, (size_t)ix->n_rows * ix->dim * sizeof(__half));
        dump(ix->doc_of_row, (size_t)ix->n_rows * sizeof(int32_t));
        // trailer: score-modifier attribute columns
        const int64_t trailer[3] = {ix->max_doc, ix->attr_cap, (int64_t)ix->attr_cols.size()};
        ok = ok && fwrite(trailer, sizeof(trailer), 1, f) == 1;
        for (double* col : ix->attr_cols) {
            const int32_t present = col ? 1 : 0;
            ok = ok && fwrite(&present, sizeof(present), 1, f) == 1;
            if (col) dump(col, (size_t)ix->attr_cap * sizeof(double));
        }
        ok = (fclose(f) == 0) && ok;
        if (!ok) fail(B200_ERR_INTERNAL, "short write to %s", path);
    });
}

int b200_index_load(int device, const char* path, b200_index** out) {
    return guarded([&] {
        MB_CHECK_ARG(path && out, "NULL argument");
        *out = nullptr;
        FILE* f = fopen(path, "rb");
        if (!f) fail(B200_ERR_INVALID_ARG, "cannot open %s", path);
        struct {
            char magic[8];
            int32_t version, dim, metric, has_docs;
            int64_t n_rows;
        } hdr;
        b200_index* ix = nullptr;
        try {
            if (fread(&hdr, sizeof(hdr), 1, f) != 1 || memcmp(hdr.magic, "B200IDX\0", 8) != 0 ||
                (hdr.version != 1 && hdr.version != 2))
                fail(B200_ERR_INVALID_ARG, "%s is not a marqo_b200 index snapshot", path);
            ix = index_new(device, hdr.dim, hdr.metric, hdr.n_rows);
            DeviceGuard g(device);
            std::vector<uint8_t> buf((size_t)1 << 24);
            auto slurp = [&](void* dptr, size_t bytes) {
                for (size_t o = 0; o < bytes; o += buf.size()) {
                    const size_t c = std::min(buf.size(), bytes - o);
                    if (fread(buf.data(), 1, c, f) != c) fail(B200_ERR_INVALID_ARG, "%s is truncated", path);
                    MB_CUDA(cudaMemcpy((uint8_t*)dptr + o, buf.data(), c, cudaMemcpyHostToDevice));
                }
            };
            slurp(ix->corpus, (size_t)hdr.n_rows * hdr.dim * sizeof(__half));
            slurp(ix->doc_of_row, (size_t)hdr.n_rows * sizeof(int32_t));
            ix->n_rows = hdr.n_rows;
            ix->has_docs = hdr.has_docs != 0;
            if (hdr.version >= 2) {
                int64_t trailer[3];
                if (fread(trailer, sizeof(trailer), 1, f) != 1) fail(B200_ERR_INVALID_ARG, "%s is truncated", path);
                MB_CHECK_ARG(trailer[2] >= 0 && trailer[2] <= B200_MAX_ATTRIBUTE_COLUMNS && trailer[1] >= 0,
                             "%s has a corrupt attribute trailer", path);
                ix->max_doc = trailer[0];
                ix->attr_cap = trailer[1];
                ix->attr_cols.assign((size_t)trailer[2], nullptr);
                for (size_t c = 0; c < ix->attr_cols.size(); ++c) {
                    int32_t present = 0;
                    if (fread(&present, sizeof(present), 1, f) != 1) fail(B200_ERR_INVALID_ARG, "%s is truncated", path);
                    if (!present) continue;
                    cuda_alloc((void**)&ix->attr_cols[c], (size_t)ix->attr_cap * sizeof(double));
                    slurp(ix->attr_cols[c], (size_t)ix->attr_cap * sizeof(double));
                }
            } else if (ix->has_docs && hdr.n_rows > 0) {
                track_max_doc(ix, ix->doc_of_row, hdr.n_rows);
            }
            if (ix->metric == B200_METRIC_EUCLIDEAN && hdr.n_rows > 0) {
                row_norms_kernel<<<(unsigned)((hdr.n_rows + 7) / 8), 256, 0, ix->stream>>>(ix->corpus, hdr.n_rows, hdr.dim,
                                                                                         ix->row_n2);
                MB_CUDA(cudaGetLastError());
                MB_CUDA(cudaStreamSynchronize(ix->stream));
            }
        } catch (...) {
            fclose(f);
            index_free(ix);
            throw;
        }
        fclose(f);
        *out = ix;
    });
}

}  // extern "C"
