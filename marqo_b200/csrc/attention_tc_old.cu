// tcgen05 flash attention (head_dim 64, S <= a few thousand, bf16 in / fp32 softmax / bf16 out).
//
// One CTA = one (batch, head, 128-query block).  Warp 0: TMA producer (Q once, K/V blocks of 128 keys through a
// 2-stage ring, all 128B-swizzled straight from the packed qkv matrix).  Warp 1: single-thread tcgen05.mma issuer:
//   S = Q K_j^T   (M=128, N=128, K=64;  A, B K-major)          -> TMEM columns [0,128)
//   O += P_j V_j  (M=128, N=64,  K=128; A = P K-major from smem, B = V MN-major as loaded)  -> TMEM columns [128,192)
// Warps 2-5: softmax, ONE THREAD PER QUERY ROW (TMEM lane == row): two passes over the row's 128 scores in TMEM
// (max, then exp2 / sum / bf16 P written to smem in the UMMA K-major swizzled layout), running-max rescale of the O
// accumulator through tcgen05.ld/st, final 1/l scaling and the bf16 store.  112 KB smem + 256 TMEM columns per CTA ->
// two CTAs per SM, so one CTA's softmax overlaps the other's MMAs.
#include <mutex>

#include "attention.cuh"
#include "ptx.cuh"

namespace mb {
namespace attention {

namespace tc4 {

constexpr int HD = 64;
constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int THREADS = 192;
constexpr int KV_STAGES = 2;
constexpr uint32_t Q_BYTES = BQ * HD * 2;        // 16 KB
constexpr uint32_t KV_TILE_BYTES = BKV * HD * 2;  // 16 KB each for K and V
constexpr uint32_t P_BYTES = BQ * BKV * 2;        // 32 KB (two 64-key K-major chunks)
constexpr uint32_t SMEM_BYTES = Q_BYTES + KV_STAGES * 2 * KV_TILE_BYTES + P_BYTES + 128;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t S_COL = 0, O_COL = 128;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Query rows that do not fill a 128-row tile (e.g. row 256 of a 257-token ViT sequence): one warp per
// (batch, head, row).  Phase 1: lanes stride over the keys and compute the scores into shared memory; phase 2: lanes
// stride over the 64 output dims and accumulate p_j * V[j] with coalesced 128-byte row reads.
constexpr int TAIL_S_MAX = 1024;

template <int MASK>
__global__ void __launch_bounds__(128)
attention_tail_rows_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int S, int W, int H,
                           int row_lo, const int32_t* __restrict__ kv_len, float scale_log2e, int total) {
    __shared__ float s_sc[4][TAIL_S_MAX];
    const int wib = threadIdx.x >> 5;
    const int wid = blockIdx.x * 4 + wib;
    const int lane = threadIdx.x & 31;
    if (wid >= total) return;
    const int nrows = S - row_lo;
    const int qrow = row_lo + wid % nrows;
    const int h = (wid / nrows) % H;
    const int b = wid / (nrows * H);
    const size_t ld = (size_t)3 * W;
    const __nv_bfloat16* seq = qkv + (size_t)b * S * ld;
    int len = S;
    if (MASK == MASK_KEYLEN) len = min(S, max(kv_len[b], 0));
    if (MASK == MASK_CAUSAL) len = min(len, qrow + 1);
    float* sc = s_sc[wib];
    // ---- phase 1: scores (log2 domain)
    float q[HD];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(seq + (size_t)qrow * ld + h * HD);
#pragma unroll
        for (int u = 0; u < HD / 8; ++u) {
            const uint4 t4 = __ldg(qp + u);
            const uint32_t w4[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
                q[8 * u + 2 * e] = f2.x * scale_log2e;
                q[8 * u + 2 * e + 1] = f2.y * scale_log2e;
            }
        }
    }
    float m = -INFINITY;
#pragma unroll 2
    for (int key = lane; key < len; key += 32) {
        const uint4* kp = reinterpret_cast<const uint4*>(seq + (size_t)key * ld + W + h * HD);
        uint4 k4[HD / 8];
#pragma unroll
        for (int u = 0; u < HD / 8; ++u) k4[u] = __ldg(kp + u);
        float acc = 0.f;
#pragma unroll
        for (int u = 0; u < HD / 8; ++u) {
            const uint32_t w4[4] = {k4[u].x, k4[u].y, k4[u].z, k4[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
                acc = fmaf(q[8 * u + 2 * e], f2.x, acc);
                acc = fmaf(q[8 * u + 2 * e + 1], f2.y, acc);
            }
        }
        sc[key] = acc;
        m = fmaxf(m, acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    __syncwarp();
    // ---- softmax numerators (fp32 sum; bf16-rounded P for the PV product, like the tensor-core path)
    float l = 0.f;
    for (int key = lane; key < len; key += 32) {
        const float pe = ex2(sc[key] - m);
        l += pe;
        sc[key] = __bfloat162float(__float2bfloat16_rn(pe));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
    __syncwarp();
    // ---- phase 2: out[d] = sum_j p_j V[j][d]; lane owns dims 2*lane, 2*lane+1
    const __nv_bfloat162* vcol = reinterpret_cast<const __nv_bfloat162*>(seq + 2 * W + h * HD) + lane;
    float ox = 0.f, oy = 0.f;
#pragma unroll 32
    for (int key = 0; key < len; ++key) {   // 32 independent 128-byte row reads in flight per warp
        const float2 v2 = __bfloat1622float2(vcol[(size_t)key * (ld / 2)]);
        const float pj = sc[key];
        ox = fmaf(pj, v2.x, ox);
        oy = fmaf(pj, v2.y, oy);
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(out + ((size_t)b * S + qrow) * W + h * HD) + lane;
    *dst = __floats2bfloat162_rn(ox * inv, oy * inv);
}

template <int MASK>
__global__ void __launch_bounds__(THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __nv_bfloat16* __restrict__ qkv,
                    __nv_bfloat16* __restrict__ out, int S, int W, const int32_t* __restrict__ kv_len, float scale_log2e,
                    int s_main, int inline_tail_rows) {
    // Keys [0, s_main) go through the tensor cores in blocks of 128; the few keys [s_main, S) of a sequence length
    // such as 257 = 2 * 128 + 1 (ViT class token) are folded in on the CUDA cores in the epilogue instead of paying
    // for a whole extra 128-wide block.
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sKV = sQ + Q_BYTES;                       // stage s: K at sKV + s*32K, V at +16K
    uint8_t* sP = sKV + KV_STAGES * 2 * KV_TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;    // [2]
    uint64_t* kv_empty = bars + 3;   // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* s_free = bars + 6;
    uint64_t* p_full = bars + 7;
    uint64_t* pv_done = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
    if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();  // the swizzled tiles need 1024-byte alignment

    int len = S;
    if (MASK == MASK_KEYLEN) len = min(S, max(kv_len[b], 0));
    int kend = min(len, s_main);
    if (MASK == MASK_CAUSAL) kend = min(kend, q0 + BQ);
    const int nkb = (kend + BKV - 1) / BKV;
    const int row_base = b * S;  // first row of this sequence in the packed [B*S, 3W] matrix

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap);
        ptx::mbar_init(q_full, 1);
        for (int i = 0; i < KV_STAGES; ++i) {
            ptx::mbar_init(&kv_full[i], 1);
            ptx::mbar_init(&kv_empty[i], 1);
        }
        ptx::mbar_init(s_full, 1);
        ptx::mbar_init(s_free, 4);
        ptx::mbar_init(p_full, 4);
        ptx::mbar_init(pv_done, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc_n<TMEM_COLS>(tmem_slot);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            ptx::mbar_arrive_expect_tx(q_full, Q_BYTES);
            ptx::tma_load_2d(sQ, &tmap, q_full, h * HD, row_base + q0, ptx::kEvictNormal);
            for (int j = 0; j < nkb; ++j) {
                const int st = j & 1;
                ptx::mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
                ptx::mbar_arrive_expect_tx(&kv_full[st], 2 * KV_TILE_BYTES);
                uint8_t* dst = sKV + (size_t)st * 2 * KV_TILE_BYTES;
                ptx::tma_load_2d(dst, &tmap, &kv_full[st], W + h * HD, row_base + j * BKV, ptx::kEvictLast);
                ptx::tma_load_2d(dst + KV_TILE_BYTES, &tmap, &kv_full[st], 2 * W + h * HD, row_base + j * BKV,
                                 ptx::kEvictLast);
            }
        }
        __syncwarp();
        // ---------------------------------------------------------------- remainder query rows (S = k*128 + r)
        // The producer warp is idle once the loads are issued: in the CTA of the last full query block it computes
        // the r <= 8 remainder rows against the K / V tiles while they sit in shared memory (the host only enables
        // this when all key blocks fit the ring without reuse), instead of a second kernel re-reading K and V from HBM.
        if (inline_tail_rows > 0 && blockIdx.x == gridDim.x - 1) {
            const size_t ld = (size_t)3 * W;
            for (int tr = 0; tr < inline_tail_rows; ++tr) {
                const int trow = s_main + tr;
                int tlimit = len;                               // keys >= tlimit are masked for this row
                if (MASK == MASK_CAUSAL) tlimit = min(tlimit, trow + 1);
                const __nv_bfloat16* qrow_p = qkv + ((size_t)row_base + trow) * ld + h * HD;
                float qf[HD];
                {
                    const uint4* qp = reinterpret_cast<const uint4*>(qrow_p);
#pragma unroll
                    for (int u = 0; u < HD / 8; ++u) {
                        const uint4 t4 = __ldg(qp + u);
                        const uint32_t w4[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
                            qf[8 * u + 2 * e] = f2.x * scale_log2e;
                            qf[8 * u + 2 * e + 1] = f2.y * scale_log2e;
                        }
                    }
                }
                const float2 qpair = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(qrow_p)[lane]);
                float m_t = -INFINITY, l_t = 0.f, ox = 0.f, oy = 0.f;
                for (int j = 0; j < nkb; ++j) {
                    const int st = j & 1;
                    ptx::mbar_wait(&kv_full[st], (j >> 1) & 1);
                    const uint8_t* kt = sKV + (size_t)st * 2 * KV_TILE_BYTES;
                    const uint8_t* vt = kt + KV_TILE_BYTES;
                    float sc[4];
                    float bm = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int key = lane + 32 * i;   // row & 7 == lane & 7: the 8 swizzled units spread over the banks
                        float acc = 0.f;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const uint4 k4 = *reinterpret_cast<const uint4*>(kt + key * 128 + ((u ^ (key & 7)) << 4));
                            const uint32_t w4[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
                                acc = fmaf(qf[8 * u + 2 * e], f2.x, acc);
                                acc = fmaf(qf[8 * u + 2 * e + 1], f2.y, acc);
                            }
                        }
                        const int gkey = j * BKV + key;
                        sc[i] = (gkey < tlimit && gkey < s_main) ? acc : -INFINITY;
                        bm = fmaxf(bm, sc[i]);
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, off));
                    const float m_new = fmaxf(m_t, bm);
                    const float m_safe = m_new == -INFINITY ? 0.f : m_new;
                    const float alpha = ex2(m_t - m_safe);
                    float pb[4];
                    float ls = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float pe = sc[i] == -INFINITY ? 0.f : ex2(sc[i] - m_safe);
                        ls += pe;
                        pb[i] = __bfloat162float(__float2bfloat16_rn(pe));   // same P rounding as the MMA path
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, off);
                    l_t = l_t * alpha + ls;
                    ox *= alpha;
                    oy *= alpha;
                    m_t = m_new;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
#pragma unroll 8
                        for (int src = 0; src < 32; ++src) {
                            const int key = src + 32 * i;
                            const float pk = __shfl_sync(0xffffffffu, pb[i], src);
                            const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(
                                vt + key * 128 + (((lane >> 2) ^ (key & 7)) << 4) + (lane & 3) * 4);
                            const float2 vf = __bfloat1622float2(v2);
                            ox = fmaf(pk, vf.x, ox);
                            oy = fmaf(pk, vf.y, oy);
                        }
                    }
                }
                // remainder keys from global memory (lanes over the 64 dims, warp-reduced dot)
                for (int key = s_main; key < tlimit; ++key) {
                    const float2 kf = __bfloat1622float2(
                        reinterpret_cast<const __nv_bfloat162*>(qkv + ((size_t)row_base + key) * ld + W + h * HD)[lane]);
                    const float2 vf = __bfloat1622float2(
                        reinterpret_cast<const __nv_bfloat162*>(qkv + ((size_t)row_base + key) * ld + 2 * W + h * HD)[lane]);
                    float sd = qpair.x * kf.x + qpair.y * kf.y;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, off);
                    const float scv = sd * scale_log2e;
                    const float m_new = fmaxf(m_t, scv);
                    const float alpha = ex2(m_t - m_new);
                    const float pe = ex2(scv - m_new);
                    const float pbv = __bfloat162float(__float2bfloat16_rn(pe));
                    l_t = l_t * alpha + pe;
                    ox = fmaf(ox, alpha, pbv * vf.x);
                    oy = fmaf(oy, alpha, pbv * vf.y);
                    m_t = m_new;
                }
                const float inv_t = l_t > 0.f ? 1.f / l_t : 0.f;
                reinterpret_cast<__nv_bfloat162*>(out + ((size_t)row_base + trow) * W + h * HD)[lane] =
                    __floats2bfloat162_rn(ox * inv_t, oy * inv_t);
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc_s = ptx::make_idesc_f16_major(1, BQ, BKV, 0, 0);
        constexpr uint32_t idesc_o = ptx::make_idesc_f16_major(1, BQ, HD, 0, 1);  // B (= V) is MN-major
        ptx::mbar_wait(q_full, 0);
        for (int j = 0; j < nkb; ++j) {
            const int st = j & 1;
            const uint32_t par = j & 1;
            ptx::mbar_wait(&kv_full[st], (j >> 1) & 1);
            if (j > 0) ptx::mbar_wait(s_free, par ^ 1);  // softmax has finished reading S of block j-1
            ptx::tc_fence_after();
            const uint32_t k_base = ptx::smem_u32(sKV + (size_t)st * 2 * KV_TILE_BYTES);
            const uint32_t v_base = k_base + KV_TILE_BYTES;
            if (lane == 0) {
                const uint32_t q_base = ptx::smem_u32(sQ);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    ptx::umma_f16(tmem_base + S_COL, ptx::make_desc_k_sw128(q_base + k * 32),
                                  ptx::make_desc_k_sw128(k_base + k * 32), idesc_s, k != 0 ? 1u : 0u);
                ptx::umma_commit(s_full);
            }
            __syncwarp();
            ptx::mbar_wait(p_full, par);  // P_j is in smem and O has been rescaled
            ptx::tc_fence_after();
            if (lane == 0) {
                const uint32_t p_base = ptx::smem_u32(sP);
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k) {
                    const uint32_t a_addr = p_base + (k >> 2) * (BQ * 128) + (k & 3) * 32;
                    const uint32_t b_addr = v_base + k * 16 * 128;  // 16 keys = two 8-key swizzle atoms
                    ptx::umma_f16(tmem_base + O_COL, ptx::make_desc_k_sw128(a_addr),
                                  ptx::make_desc_mn_sw128(b_addr, 8192, 1024), idesc_o, (j | k) != 0 ? 1u : 0u);
                }
                ptx::umma_commit(&kv_empty[st]);
                ptx::umma_commit(pv_done);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ softmax: thread == query row
        const int sp = warp & 3;
        const int r = sp * 32 + lane;  // row within the tile == TMEM lane
        const int qrow = q0 + r;       // position in the sequence
        const uint32_t lane_addr = tmem_base + (uint32_t(sp * 32) << 16);
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < nkb; ++j) {
            const uint32_t par = j & 1;
            ptx::mbar_wait(s_full, par);
            ptx::tc_fence_after();
            int limit = kend - j * BKV;  // keys with block-local index >= limit are masked
            if (MASK == MASK_CAUSAL) limit = min(limit, qrow - j * BKV + 1);
            const bool full = limit >= BKV;
            // pass 1: row maximum of this block
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < BKV / 32; ++c) {
                uint32_t v[32];
                ptx::tmem_ld_32x32b_x32(lane_addr + S_COL + c * 32, v);
                ptx::tmem_ld_wait();
                if (full) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c * 32 + i < limit) mx = fmaxf(mx, __uint_as_float(v[i]));
                }
            }
            mx *= scale_log2e;  // scale > 0: max commutes with the scaling
            const float m_new = fmaxf(m_run, mx);
            const float m_safe = m_new == -INFINITY ? 0.f : m_new;
            const float alpha = ex2(m_run - m_safe);  // 0 on the first block
            if (j > 0) ptx::mbar_wait(pv_done, par ^ 1);  // P buffer and O accumulator are free again
            // pass 2: p = exp2(s - m), row sum, bf16 P into the K-major 128B-swizzled A-operand layout
            float lsum = 0.f;
#pragma unroll 1
            for (int c = 0; c < BKV / 32; ++c) {
                uint32_t v[32];
                ptx::tmem_ld_32x32b_x32(lane_addr + S_COL + c * 32, v);
                ptx::tmem_ld_wait();
                uint32_t pk[16];
                if (full) {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float p0 = ex2(fmaf(__uint_as_float(v[i]), scale_log2e, -m_safe));
                        const float p1 = ex2(fmaf(__uint_as_float(v[i + 1]), scale_log2e, -m_safe));
                        lsum += p0 + p1;
                        __nv_bfloat162 t2 = __floats2bfloat162_rn(p0, p1);
                        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&t2);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float p0 = c * 32 + i < limit ? ex2(fmaf(__uint_as_float(v[i]), scale_log2e, -m_safe)) : 0.f;
                        const float p1 =
                            c * 32 + i + 1 < limit ? ex2(fmaf(__uint_as_float(v[i + 1]), scale_log2e, -m_safe)) : 0.f;
                        lsum += p0 + p1;
                        __nv_bfloat162 t2 = __floats2bfloat162_rn(p0, p1);
                        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&t2);
                    }
                }
                // keys c*32 .. c*32+31 -> chunk (c >> 1), 16-byte units (c & 1) * 4 .. +3 of row r
                uint8_t* rowp = sP + (size_t)(c >> 1) * (BQ * 128) + (size_t)r * 128;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int unit = (c & 1) * 4 + u;
                    *reinterpret_cast<uint4*>(rowp + ((unit ^ (r & 7)) << 4)) =
                        make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                }
            }
            l_run = l_run * alpha + lsum;
            m_run = m_new;
            // S has been consumed: the MMA warp may overwrite it with the next block's scores
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(s_free);
            if (j > 0) {
                // rescale the running output by alpha (thread-local: lane == row)
#pragma unroll 1
                for (int c = 0; c < HD / 32; ++c) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32b_x32(lane_addr + O_COL + c * 32, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                    ptx::tmem_st_32x32b_x32(lane_addr + O_COL + c * 32, v);
                }
                ptx::tmem_st_wait();
            }
            ptx::fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core's async proxy
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(p_full);
        }
        // ------------------------------------------------------------------ epilogue: (+ tail keys) O / l -> bf16
        float o[HD];
        if (nkb > 0) {
            ptx::mbar_wait(pv_done, (nkb - 1) & 1);
            ptx::tc_fence_after();
            uint32_t v0[32], v1[32];
            ptx::tmem_ld_32x32b_x32(lane_addr + O_COL, v0);
            ptx::tmem_ld_32x32b_x32(lane_addr + O_COL + 32, v1);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                o[i] = __uint_as_float(v0[i]);
                o[32 + i] = __uint_as_float(v1[i]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < HD; ++i) o[i] = 0.f;
        }
        if (qrow < S) {
            const int tail_end = MASK == MASK_CAUSAL ? min(len, qrow + 1) : len;
            if (s_main < tail_end) {
                // tail keys on the CUDA cores: s = q . k, online-softmax update of (m, l, o) with p * v
                const size_t ld = (size_t)3 * W;
                uint32_t qreg[HD / 2];
                const uint4* qp = reinterpret_cast<const uint4*>(qkv + ((size_t)row_base + qrow) * ld + h * HD);
#pragma unroll
                for (int u = 0; u < HD / 8; ++u) {
                    const uint4 t4 = qp[u];
                    qreg[4 * u] = t4.x;
                    qreg[4 * u + 1] = t4.y;
                    qreg[4 * u + 2] = t4.z;
                    qreg[4 * u + 3] = t4.w;
                }
                for (int key = s_main; key < tail_end; ++key) {
                    const uint4* kp = reinterpret_cast<const uint4*>(qkv + ((size_t)row_base + key) * ld + W + h * HD);
                    const uint4* vp = reinterpret_cast<const uint4*>(qkv + ((size_t)row_base + key) * ld + 2 * W + h * HD);
                    float sdot = 0.f;
#pragma unroll
                    for (int u = 0; u < HD / 8; ++u) {
                        const uint4 k4 = __ldg(kp + u);
                        const uint32_t kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 qa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&qreg[4 * u + e]));
                            const float2 ka = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&kk[e]));
                            sdot = fmaf(qa.x, ka.x, sdot);
                            sdot = fmaf(qa.y, ka.y, sdot);
                        }
                    }
                    const float sc = sdot * scale_log2e;
                    const float m_new = fmaxf(m_run, sc);
                    const float alpha = ex2(m_run - m_new);
                    // the tensor-core path rounds P to bf16 before the PV product: do the same here
                    const float pexp = __bfloat162float(__float2bfloat16_rn(ex2(sc - m_new)));
                    l_run = l_run * alpha + ex2(sc - m_new);
                    m_run = m_new;
#pragma unroll
                    for (int u = 0; u < HD / 8; ++u) {
                        const uint4 v4 = __ldg(vp + u);
                        const uint32_t vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 va = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vv[e]));
                            o[8 * u + 2 * e] = fmaf(o[8 * u + 2 * e], alpha, pexp * va.x);
                            o[8 * u + 2 * e + 1] = fmaf(o[8 * u + 2 * e + 1], alpha, pexp * va.y);
                        }
                    }
                }
            }
            const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
            uint4* d4 = reinterpret_cast<uint4*>(out + ((size_t)row_base + qrow) * W + h * HD);
#pragma unroll
            for (int u = 0; u < HD / 8; ++u) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    __nv_bfloat162 t2 = __floats2bfloat162_rn(o[8 * u + 2 * e] * inv, o[8 * u + 2 * e + 1] * inv);
                    pk[e] = *reinterpret_cast<uint32_t*>(&t2);
                }
                d4[u] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace tc4

int launch_tc4(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int S, int W, int H, int mask, const int32_t* kv_len,
              cudaStream_t stream) {
    if (B <= 0 || S <= 0) return 0;
    if (W != H * tc4::HD) fail(B200_ERR_UNSUPPORTED, "attention: head_dim must be 64 (width %d, heads %d)", W, H);
    static std::once_flag once;
    std::call_once(once, [] {
        MB_CUDA(cudaFuncSetAttribute(tc4::attention_tc_kernel<MASK_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)tc4::SMEM_BYTES));
        MB_CUDA(cudaFuncSetAttribute(tc4::attention_tc_kernel<MASK_CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)tc4::SMEM_BYTES));
        MB_CUDA(cudaFuncSetAttribute(tc4::attention_tc_kernel<MASK_KEYLEN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)tc4::SMEM_BYTES));
    });
    // one tensor map over the packed [B*S, 3W] matrix serves Q, K and V tiles (64 columns x 128 rows, 128B swizzle);
    // rows past the end of the matrix are zero-filled, rows of the next sequence are masked by key index
    CUtensorMap tmap = make_tmap_2d(qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)3 * W, (uint64_t)B * S,
                                    (uint64_t)3 * W * 2, tc4::HD, tc4::BQ, CU_TENSOR_MAP_SWIZZLE_128B);
    // A short remainder (S = 257, 129, ...) is not worth a 128-wide tile in either dimension.
    constexpr int TAIL_MAX = 8;
    const int rem = S % tc4::BQ;
    const bool tail = S >= tc4::BQ && rem > 0 && rem <= TAIL_MAX && S <= tc4::TAIL_S_MAX;
    const int s_main = tail ? S - rem : S;                       // keys handled by the tensor cores
    const int q_blocks = tail ? S / tc4::BQ : (S + tc4::BQ - 1) / tc4::BQ;
    const dim3 grid(q_blocks, H, B);
    const float scale_log2e = 0.125f * 1.4426950408889634f;
    // remainder query rows inside the main kernel when every key block fits the 2-stage ring without reuse
    const bool inline_tail = tail && s_main / tc4::BKV <= tc4::KV_STAGES;
    const int inline_rows = inline_tail ? rem : 0;
    const int tail_total = (tail && !inline_tail) ? B * H * rem : 0;
    switch (mask) {
        case MASK_NONE:
            tc4::attention_tc_kernel<MASK_NONE><<<grid, tc4::THREADS, tc4::SMEM_BYTES, stream>>>(tmap, qkv, out, S, W, kv_len,
                                                                                             scale_log2e, s_main, inline_rows);
            if (tail_total > 0)
                tc4::attention_tail_rows_kernel<MASK_NONE><<<(tail_total + 3) / 4, 128, 0, stream>>>(
                    qkv, out, S, W, H, s_main, kv_len, scale_log2e, tail_total);
            break;
        case MASK_CAUSAL:
            tc4::attention_tc_kernel<MASK_CAUSAL><<<grid, tc4::THREADS, tc4::SMEM_BYTES, stream>>>(tmap, qkv, out, S, W, kv_len,
                                                                                               scale_log2e, s_main, inline_rows);
            if (tail_total > 0)
                tc4::attention_tail_rows_kernel<MASK_CAUSAL><<<(tail_total + 3) / 4, 128, 0, stream>>>(
                    qkv, out, S, W, H, s_main, kv_len, scale_log2e, tail_total);
            break;
        case MASK_KEYLEN:
            if (!kv_len) fail(B200_ERR_INTERNAL, "attention: kv_len required for key-length masking");
            tc4::attention_tc_kernel<MASK_KEYLEN><<<grid, tc4::THREADS, tc4::SMEM_BYTES, stream>>>(tmap, qkv, out, S, W, kv_len,
                                                                                               scale_log2e, s_main, inline_rows);
            if (tail_total > 0)
                tc4::attention_tail_rows_kernel<MASK_KEYLEN><<<(tail_total + 3) / 4, 128, 0, stream>>>(
                    qkv, out, S, W, H, s_main, kv_len, scale_log2e, tail_total);
            break;
        default:
            fail(B200_ERR_INTERNAL, "attention: unknown mask mode %d", mask);
    }
    MB_CUDA(cudaGetLastError());
    return tail_total > 0 ? 2 : 1;
}

}  // namespace attention
}  // namespace mb
