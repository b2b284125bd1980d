#include "gemm.cuh"

#include <cstdlib>
#include <mutex>

#include "ptx.cuh"

namespace mb {
namespace gemm {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
// Epilogue warps (warp 0 TMA, warp 1 MMA, warps 2.. epilogue):
//   EW = 8   general epilogue (fp32 / residual / token scatter): two warps per TMEM sub-partition, each draining one
//            half of the tile's columns through a 32 x 128 B transpose buffer.
//   EW = 16  bf16-out GEMMs without residual (QKV, fc1 + GELU): FOUR warps per sub-partition, a quarter of the columns
//            each.  The r01 profile of fc1 showed the 8-warp epilogue at 49 % issue-active with two warps per scheduler
//            (31 % of their stall samples fixed-latency dependencies, 16 % tcgen05.ld waits) while the tensor pipe idled
//            at 62 %: the erf-GELU epilogue took as long as the tile's MMAs.  Twice the warps = twice the independent
//            instruction streams per scheduler to hide those latencies; the staging buffer shrinks to 32 x 64 B per
//            warp so the smem ring keeps its depth.
constexpr int ACC_STAGES = 2;
constexpr uint32_t A_STAGE_BYTES = BM * BK * 2;
constexpr int SMEM_LIMIT = 232448;

template <int BN, int EW>
struct Cfg {
    static constexpr int THREADS = 64 + 32 * EW;
    static constexpr uint32_t B_STAGE_BYTES = (BN / 2) * BK * 2;   // each CTA of the pair holds half of the W tile
    static constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    // per epilogue warp: a 32-row transpose buffer (128-byte rows; 64-byte rows for EW = 16) + a 128-byte bias row
    static constexpr uint32_t EPI_ROW_BYTES = EW == 16 ? 64 : 128;
    // (EW = 16 keeps the bias row inside the transpose buffer, so the smem ring stays 6 stages deep at BN = 256)
    static constexpr uint32_t EPI_WARP_BYTES = 32 * EPI_ROW_BYTES + (EW == 16 ? 0 : 128);
    static constexpr uint32_t EPI_BYTES = EW * EPI_WARP_BYTES;
    static constexpr int STAGES_RAW = (SMEM_LIMIT - 2048 - (int)EPI_BYTES) / (int)STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 512 /*barriers*/;
    static constexpr uint32_t TMEM_COLS = ACC_STAGES * BN < 32 ? 32 : ACC_STAGES * BN;
};

// CTA pairs (cta_group::2): one tcgen05.mma spans both SMs of a pair — a 256 x BN tile, 128 rows of A and BN/2 rows
// of W in each SM's shared memory.  Per SM and k-block the tensor core reads 16 KB instead of 24 KB and TMA writes
// 32 KB instead of 48 KB, which takes the single-CTA kernel off its shared-memory-bandwidth ceiling (ncu: the MMA
// warp never waited for data, yet the tensor pipe stalled at ~70 %).
constexpr int CLUSTER = 2;

struct Params {
    int M, N, K;
    int tiles_m, tiles_n;   // tiles_m counts 128-row blocks
    int super_m;            // ceil(tiles_m / CLUSTER)
    Epilogue ep;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// erf(a) = sign(a) * (1 - exp(P(|a|))) with the single-branch minimax polynomial of the large-argument branch of the
// usual float erf (coefficients pre-multiplied by log2(e) so the exponential is one ex2.approx).  Max abs error of
// erf 1.9e-5, of GELU 1.9e-6 (5.9e-5 relative) — two orders below the bf16 rounding applied to the result
// (tests/test_kernels_gpu.py::test_gemm_epilogues compares against torch's exact erf GELU).
__device__ __forceinline__ float gelu_erf(float x) {
    const float a = x * 0.70710678118654752440f;
    const float t = fabsf(a);
    const float s = a * a;
    float r = fmaf(-2.49374837e-5f, t, 5.52836593e-4f);
    const float u = fmaf(-5.60337594e-3f, t, 3.49920232e-2f);
    r = fmaf(r, s, u);
    r = fmaf(r, t, -1.54047912e-1f);
    r = fmaf(r, t, -9.15890168e-1f);
    r = fmaf(r, t, -1.85700115e-1f);
    r = fmaf(r, t, -1.44269504f * t);
    const float e = copysignf(1.0f - ex2_approx(r), a);
    const float hx = 0.5f * x;
    return fmaf(hx, e, hx);
}

__device__ __forceinline__ float apply_act(float x, int act) {
    if (act == ACT_GELU) return gelu_erf(x);
    if (act == ACT_QUICKGELU) return x / (1.0f + ex2_approx(-1.702f * 1.44269504f * x));
    return x;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

template <int BN, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, Params p) {
    using C = Cfg<BN, EW>;
    constexpr int EPI_WARPS = EW;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)C::STAGES * A_STAGE_BYTES;
    uint8_t* smem_epi = smem + (size_t)C::STAGES * C::STAGE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_epi + C::EPI_BYTES);
    uint64_t* empty = full + C::STAGES;
    uint64_t* tfull = empty + C::STAGES;
    uint64_t* tempty = tfull + ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + ACC_STAGES);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int kblocks = p.K / BK;
    const uint32_t crank = ptx::cluster_ctarank();
    const int cluster_id = blockIdx.x / CLUSTER;
    const int num_clusters = gridDim.x / CLUSTER;
    const int num_super = p.super_m * p.tiles_n;   // (pair of m-blocks) x n-block units, one per cluster step

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
        for (int i = 0; i < C::STAGES; ++i) {
            ptx::mbar_init(&full[i], CLUSTER);    // leader's copy is used: its own arrive.expect_tx + the peer's arrive
            ptx::mbar_init(&empty[i], 1);         // one multicast tcgen05.commit per use, in each CTA
        }
        for (int i = 0; i < ACC_STAGES; ++i) {
            ptx::mbar_init(&tfull[i], 1);
            ptx::mbar_init(&tempty[i], CLUSTER * EPI_WARPS);   // leader's copy: epilogue warps of both CTAs
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {   // executed by both CTAs of the pair
        ptx::tmem_alloc_2sm<C::TMEM_COLS>(tmem_slot);
        ptx::tmem_relinquish_2sm();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();   // the peer's barriers are initialised before any multicast can land there
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = cluster_id; t < num_super; t += num_clusters) {
                const int m0 = ((t / p.tiles_n) * CLUSTER + (int)crank) * BM;
                const int n0 = (t % p.tiles_n) * BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    // completion of BOTH CTAs' loads is tracked by the leader's full barrier
                    const uint32_t leader_full = ptx::mapa_u32(ptx::smem_u32(&full[stage]), 0);
                    if (crank == 0)
                        ptx::mbar_arrive_expect_tx(&full[stage], CLUSTER * C::STAGE_BYTES);
                    else
                        ptx::mbar_arrive_cluster(leader_full);
                    ptx::tma_load_2d_2sm(smem_a + (size_t)stage * A_STAGE_BYTES, &tmap_a, leader_full, kb * BK, m0,
                                         ptx::kEvictNormal);
                    ptx::tma_load_2d_2sm(smem_b + (size_t)stage * C::B_STAGE_BYTES, &tmap_b, leader_full, kb * BK,
                                         n0 + (int)crank * (BN / CLUSTER), ptx::kEvictLast);
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1 && crank == 0) {
        constexpr uint32_t idesc = ptx::make_idesc_f16(1 /*bf16*/, CLUSTER * BM, BN);   // M = 256 across the pair
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int t = cluster_id; t < num_super; t += num_clusters) {
            ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
            ptx::tc_fence_after();
            for (int kb = 0; kb < kblocks; ++kb) {
                ptx::mbar_wait(&full[stage], phase);
                ptx::tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_base = ptx::smem_u32(smem_a + (size_t)stage * A_STAGE_BYTES);
                    const uint32_t b_base = ptx::smem_u32(smem_b + (size_t)stage * C::B_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        ptx::umma_f16_2sm(tmem_base + acc * BN, ptx::make_desc_k_sw128(a_base + k * UMMA_K * 2),
                                          ptx::make_desc_k_sw128(b_base + k * UMMA_K * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    // both CTAs' producers get their stage back; both CTAs' epilogues get the finished accumulator
                    ptx::umma_commit_2sm(&empty[stage], (uint16_t)((1u << CLUSTER) - 1));
                    if (kb == kblocks - 1) ptx::umma_commit_2sm(&tfull[acc], (uint16_t)((1u << CLUSTER) - 1));
                }
                __syncwarp();
                if (++stage == C::STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (++acc == ACC_STAGES) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    } else if (warp >= 2 && EW == 16) {
        // ---------------------------------------------------------------- fast bf16 epilogue (16 warps)
        // lane == accumulator row; each warp drains 32 rows x (BN / 4) columns in 32-column chunks:
        // bias -> activation -> bf16 -> 32 x 64 B swizzled staging buffer -> 64-byte coalesced row segments.
        const int sp = warp & 3;
        const int part = (warp - 2) >> 2;                  // column quarter of the tile
        constexpr int PART_COLS = BN / 4;
        constexpr int CHUNKS = PART_COLS / 32;
        static_assert(EW != 16 || (BN % 128 == 0), "the 16-warp epilogue needs whole 32-column chunks per quarter");
        const Epilogue& ep = p.ep;
        uint8_t* stage_buf = smem_epi + (size_t)(warp - 2) * C::EPI_WARP_BYTES;
        float* bias_row = reinterpret_cast<float*>(stage_buf);   // first 128 B of the buffer, consumed before it is filled
        const int frow = lane >> 2, funit = lane & 3;      // flush phase: 8 rows x 4 sixteen-byte units per instruction
        const int wswz = (lane >> 1) & 3;                  // write-phase swizzle of this lane's own row
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = cluster_id; t < num_super; t += num_clusters) {
            const int m0 = ((t / p.tiles_n) * CLUSTER + (int)crank) * BM;
            const int nt0 = (t % p.tiles_n) * BN + part * PART_COLS;
            const int wrow0 = m0 + sp * 32;
            ptx::mbar_wait(&tfull[acc], acc_phase);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < CHUNKS; ++c) {
                const int n0 = nt0 + c * 32;
                const bool cols_ok = n0 < p.N;
                float bias_v = 0.f;
                if (ep.bias && cols_ok) bias_v = __ldg(ep.bias + n0 + lane);
                uint32_t v[32];
                ptx::tmem_ld_32x32b_x32(tmem_base + (uint32_t(sp * 32) << 16) + acc * BN + part * PART_COLS + c * 32, v);
                ptx::tmem_ld_wait();
                if (c == CHUNKS - 1) {
                    // this warp's share of the accumulator is in registers: hand it back to the MMA warp
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(&tempty[acc]), 0));
                }
                if (cols_ok) {
                    bias_row[lane] = bias_v;
                    __syncwarp();
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b = *reinterpret_cast<const float4*>(bias_row + 4 * j);   // broadcast read
                        f[4 * j] = __uint_as_float(v[4 * j]) + b.x;
                        f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b.y;
                        f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b.z;
                        f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b.w;
                    }
                    __syncwarp();   // every lane has read the bias row: the buffer may be overwritten
                    if (ep.act != ACT_NONE) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], ep.act);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(stage_buf + lane * 64 + ((j ^ wswz) << 4)) =
                            make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                       pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
                    __syncwarp();
                    const int col = n0 + funit * 8;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr = i * 8 + frow;
                        const int grow = wrow0 + rr;
                        if (grow < p.M) {
                            const uint4 val = *reinterpret_cast<const uint4*>(stage_buf + rr * 64 + ((funit ^ ((rr >> 1) & 3)) << 4));
                            *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(ep.out) + ((size_t)grow * ep.ldo + col) * 2) = val;
                        }
                    }
                    __syncwarp();   // the buffer (and bias_row) are reused by the next chunk
                }
            }
            if (++acc == ACC_STAGES) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    } else if (warp >= 2) {
        // ---------------------------------------------------------------- epilogue (TMEM -> regs -> smem -> global)
        // A warp may only read the TMEM lanes of sub-partition (warp % 4); the two warps that share a sub-partition
        // split the tile's columns in halves, so 8 warps drain one 128 x BN accumulator.  TMEM hands every lane one ROW
        // (32 consecutive columns); global memory wants whole 128-byte lines per instruction, so each warp transposes
        // through a private, XOR-swizzled 32 x 128 B shared-memory buffer in both directions (residual in, result out).
        const int sp = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int HALF_COLS = BN / 2 >= 32 ? BN / 2 : 32;
        constexpr int CHUNKS = HALF_COLS / 32;
        const bool has_cols = half * HALF_COLS < BN;  // BN = 32 would leave the second half empty
        const Epilogue& ep = p.ep;
        uint8_t* stage_buf = smem_epi + (size_t)(warp - 2) * C::EPI_WARP_BYTES;
        float* bias_row = reinterpret_cast<float*>(stage_buf + 32 * 128);
        const int esz = ep.out_fp32 ? 4 : 2;                  // output element size
        const int cols_per_flush = 128 / esz;                 // 32 fp32 or 64 bf16 columns fill a 128-byte row
        // bf16 results are staged two chunks (64 columns) per flush; a residual block occupies the whole buffer, so
        // residual GEMMs (fp32 out in this engine) flush after every chunk
        const int chunks_per_flush = (ep.out_fp32 || ep.residual) ? 1 : cols_per_flush / 32;
        const int srow = lane >> 3, sunit = lane & 7;         // coalesced phase: 4 rows x 8 sixteen-byte units per instr
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = cluster_id; t < num_super; t += num_clusters) {
            const int m0 = ((t / p.tiles_n) * CLUSTER + (int)crank) * BM;
            const int nt0 = (t % p.tiles_n) * BN + half * HALF_COLS;
            const int wrow0 = m0 + sp * 32;                   // first row of this warp's 32-row band
            if (ep.residual && has_cols) {
                // pull the residual band this warp needs for its NEXT tile towards L2 (the very first tile: itself)
                for (int pass = (t == cluster_id ? 0 : 1); pass < 2; ++pass) {
                    const int tn = t + pass * num_clusters;
                    if (tn >= num_super) break;
                    const int prow = ((tn / p.tiles_n) * CLUSTER + (int)crank) * BM + sp * 32 + lane;
                    const int pn0 = (tn % p.tiles_n) * BN + half * HALF_COLS;
                    if (prow < p.M) {
                        const char* r = reinterpret_cast<const char*>(ep.residual + (size_t)prow * ep.ldr + pn0);
#pragma unroll
                        for (int l = 0; l < HALF_COLS * 4 / 128; ++l)
                            if (pn0 + l * 32 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(r + l * 128));
                    }
                }
            }
            ptx::mbar_wait(&tfull[acc], acc_phase);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < CHUNKS; ++c) {
                const int n0 = nt0 + c * 32;
                const bool cols_ok = has_cols && n0 < p.N;
                // (1) start the long-latency global reads first: bias (one float per lane) and, for residual GEMMs,
                //     the 32 x 32 fp32 residual block in coalesced order (4 rows x 128 B per instruction)
                float bias_v = 0.f;
                if (ep.bias && cols_ok) bias_v = __ldg(ep.bias + n0 + lane);
                float4 rres[8];
                if (ep.residual && cols_ok) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rr = wrow0 + i * 4 + srow;
                        rres[i] = rr < p.M ? *reinterpret_cast<const float4*>(ep.residual + (size_t)rr * ep.ldr + n0 + sunit * 4)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                // (2) accumulator chunk: lane == row
                uint32_t v[32];
                if (has_cols) {
                    ptx::tmem_ld_32x32b_x32(tmem_base + (uint32_t(sp * 32) << 16) + acc * BN + half * HALF_COLS + c * 32, v);
                    ptx::tmem_ld_wait();
                }
                if (c == CHUNKS - 1) {
                    // this warp's share of the accumulator is in registers: hand it back to the MMA warp
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(&tempty[acc]), 0));
                }
                if (cols_ok) {
                    bias_row[lane] = bias_v;
                    if (ep.residual) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rr = i * 4 + srow;
                            *reinterpret_cast<float4*>(stage_buf + rr * 128 + ((sunit ^ (rr & 7)) << 4)) = rres[i];
                        }
                    }
                    __syncwarp();
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b = *reinterpret_cast<const float4*>(bias_row + 4 * j);   // broadcast read
                        f[4 * j] = __uint_as_float(v[4 * j]) + b.x;
                        f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b.y;
                        f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b.z;
                        f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b.w;
                    }
                    if (ep.act != ACT_NONE) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], ep.act);
                    }
                    if (ep.rowbias) {   // ViT patch-embed: positional embedding of this lane's patch
                        const int row = wrow0 + lane;
                        const int brow = 1 + row % ep.remap_group;
                        const float4* r4 = reinterpret_cast<const float4*>(ep.rowbias + (size_t)brow * p.N + n0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = __ldg(r4 + j);
                            f[4 * j] += b.x;
                            f[4 * j + 1] += b.y;
                            f[4 * j + 2] += b.z;
                            f[4 * j + 3] += b.w;
                        }
                    }
                    if (ep.residual) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = *reinterpret_cast<const float4*>(stage_buf + lane * 128 + ((j ^ (lane & 7)) << 4));
                            f[4 * j] += b.x;
                            f[4 * j + 1] += b.y;
                            f[4 * j + 2] += b.z;
                            f[4 * j + 3] += b.w;
                        }
                        __syncwarp();   // everyone has read its residual row before the buffer is overwritten
                    }
                    // (3) own row -> staging buffer (swizzled 16-byte units)
                    if (ep.out_fp32) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<float4*>(stage_buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                                make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    } else {
                        const int ubase = (c % chunks_per_flush) * 4;   // this chunk fills units 0-3 or 4-7 of the row
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<uint4*>(stage_buf + lane * 128 + (((ubase + j) ^ (lane & 7)) << 4)) =
                                make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                           pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
                    }
                }
                // (4) flush full 128-byte rows: every instruction writes 4 rows x 128 contiguous bytes
                const bool flush = (c % chunks_per_flush) == chunks_per_flush - 1 || c == CHUNKS - 1;
                if (flush && has_cols) {
                    __syncwarp();
                    const int fc0 = nt0 + (c / chunks_per_flush) * chunks_per_flush * 32;   // first column held in the buffer
                    const int col = fc0 + sunit * (16 / esz);
                    const int filled_units = ((c % chunks_per_flush) + 1) * (32 * esz / 16);
                    if (col < p.N && sunit < filled_units) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rr = i * 4 + srow;
                            const int grow = wrow0 + rr;
                            if (grow < p.M) {
                                long long orow = grow;
                                if (ep.remap_group > 0) {
                                    const int b = grow / ep.remap_group;
                                    orow = (long long)b * (ep.remap_group + 1) + 1 + (grow - b * ep.remap_group);
                                }
                                const uint4 val = *reinterpret_cast<const uint4*>(stage_buf + rr * 128 + ((sunit ^ (rr & 7)) << 4));
                                *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(ep.out) +
                                                          ((size_t)orow * ep.ldo + col) * esz) = val;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            if (++acc == ACC_STAGES) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();   // nobody exits while the peer may still multicast into / arrive on this CTA's smem
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_2sm<C::TMEM_COLS>(tmem_base);
    }
}

void configure() {
    static std::once_flag once;
    std::call_once(once, [] {
        const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
        MB_CUDA(cudaFuncSetAttribute(gemm_kernel<256, 8>, attr, (int)Cfg<256, 8>::SMEM_BYTES));
        MB_CUDA(cudaFuncSetAttribute(gemm_kernel<128, 8>, attr, (int)Cfg<128, 8>::SMEM_BYTES));
        MB_CUDA(cudaFuncSetAttribute(gemm_kernel<64, 8>, attr, (int)Cfg<64, 8>::SMEM_BYTES));
        MB_CUDA(cudaFuncSetAttribute(gemm_kernel<256, 16>, attr, (int)Cfg<256, 16>::SMEM_BYTES));
        MB_CUDA(cudaFuncSetAttribute(gemm_kernel<128, 16>, attr, (int)Cfg<128, 16>::SMEM_BYTES));
    });
}

template <int BN, int EW>
static void launch_bn(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int M, int N, int K, const Epilogue& ep,
                      int sms, cudaStream_t stream) {
    Params p;
    p.M = M;
    p.N = N;
    p.K = K;
    p.tiles_m = (M + BM - 1) / BM;
    p.tiles_n = (N + BN - 1) / BN;
    p.super_m = (p.tiles_m + CLUSTER - 1) / CLUSTER;
    p.ep = ep;
    CUtensorMap ta = make_tmap_2d(A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, BK,
                                  BM, CU_TENSOR_MAP_SWIZZLE_128B);
    // each CTA of the pair fetches (and keeps) BN / 2 rows of the W tile
    CUtensorMap tb = make_tmap_2d(W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)K, (uint64_t)N, (uint64_t)K * 2, BK,
                                  BN / CLUSTER, CU_TENSOR_MAP_SWIZZLE_128B);
    const int max_clusters = sms / CLUSTER;
    const int clusters = std::min(p.super_m * p.tiles_n, max_clusters);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(clusters * CLUSTER);
    cfg.blockDim = dim3(Cfg<BN, EW>::THREADS);
    cfg.dynamicSmemBytes = Cfg<BN, EW>::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MB_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<BN, EW>, ta, tb, p));
}

void launch(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int M, int N, int K, const Epilogue& ep, int sms,
            cudaStream_t stream) {
    if (M <= 0 || N <= 0) return;
    if (K <= 0 || K % BK != 0) fail(B200_ERR_INTERNAL, "gemm: K = %d must be a positive multiple of %d", K, BK);
    if (N % 32 != 0) fail(B200_ERR_INTERNAL, "gemm: N = %d must be a multiple of 32", N);
    if (lda % 8 != 0 || ep.ldo % 8 != 0) fail(B200_ERR_INTERNAL, "gemm: leading dimensions must be multiples of 8");
    configure();
    // bf16 output, no residual / token scatter (QKV, fc1): the 16-warp epilogue.  MARQO_B200_GEMM_EPI8=1 keeps the
    // general 8-warp epilogue for A/B timing.
    static const bool force8 = [] {
        const char* e = getenv("MARQO_B200_GEMM_EPI8");
        return e != nullptr && e[0] == '1';
    }();
    const bool fast = !force8 && !ep.out_fp32 && ep.residual == nullptr && ep.rowbias == nullptr && ep.remap_group == 0;
    // Largest tile that wastes no columns, otherwise the widest one.
    if (N % 256 == 0 || N > 512) {
        if (fast) launch_bn<256, 16>(A, lda, W, M, N, K, ep, sms, stream);
        else launch_bn<256, 8>(A, lda, W, M, N, K, ep, sms, stream);
    } else if (N % 128 == 0 || N > 128) {
        if (fast) launch_bn<128, 16>(A, lda, W, M, N, K, ep, sms, stream);
        else launch_bn<128, 8>(A, lda, W, M, N, K, ep, sms, stream);
    } else {
        launch_bn<64, 8>(A, lda, W, M, N, K, ep, sms, stream);
    }
}

}  // namespace gemm
}  // namespace mb
