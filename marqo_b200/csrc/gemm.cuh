// Persistent warp-specialised tcgen05 GEMM:  out[M,N] = epilogue( A[M,K] (bf16, K-major) x W[N,K]^T (bf16, K-major) )
// CTA pairs (cta_group::2): one tcgen05.mma covers a 256 x BN tile across the two SMs of a cluster, each SM holding
// 128 rows of A and BN/2 rows of W; fp32 accumulation in TMEM; TMA-fed 128B-swizzled smem ring (7 stages of 32 KB);
// double-buffered accumulators so the 8-warp epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include "common.cuh"

namespace mb {
namespace gemm {

enum Act { ACT_NONE = 0, ACT_GELU = 1, ACT_QUICKGELU = 2 };

struct Epilogue {
    const float* bias = nullptr;      // [N]
    const float* residual = nullptr;  // fp32 [M, ldr], added after the activation
    int ldr = 0;
    int act = ACT_NONE;
    void* out = nullptr;  // bf16 or fp32 [*, ldo]
    int ldo = 0;
    int out_fp32 = 0;
    // ViT token assembly (patch-embed): GEMM row r = (image b, patch i) with G patches per image is written to token
    // row b * (G + 1) + 1 + i and gets rowbias[(1 + i), :] (the positional embedding) added.
    int remap_group = 0;
    const float* rowbias = nullptr;  // fp32 [G + 1, N]
};

// A: bf16 [M, K] row-major with leading dimension lda (elements); W: bf16 [N, K] row-major (nn.Linear layout).
// Requirements: K % 64 == 0, N % 32 == 0, lda % 8 == 0.
void launch(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int M, int N, int K, const Epilogue& ep,
            int sm_count, cudaStream_t stream);

void configure();  // one-time cudaFuncSetAttribute calls

}  // namespace gemm
}  // namespace mb
