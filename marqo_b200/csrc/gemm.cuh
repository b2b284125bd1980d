// Persistent warp-specialised tcgen05 GEMM:  out[M,N] = epilogue( A[M,K] (bf16, K-major) x W[N,K]^T (bf16, K-major) )
// CTA pairs (cta_group::2): one tcgen05.mma covers a 256 x BN tile across the two SMs of a cluster, each SM holding
// 128 rows of A and BN/2 rows of W; fp32 accumulation in TMEM; TMA-fed 128B-swizzled smem ring (5-6 stages of 32 KB);
// double-buffered accumulators so the 8-warp epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include "common.cuh"

namespace mb {
namespace gemm {

enum Act { ACT_NONE = 0, ACT_GELU = 1, ACT_QUICKGELU = 2 };
constexpr int LN_MAX_PARTS = 16;   // column parts per row in Epilogue::ln_stats (N <= 1024, >= 64 columns per part)

struct Epilogue {
    const float* bias = nullptr;      // [N]
    const float* residual = nullptr;  // fp32 [M, ldr], added after the activation
    int ldr = 0;
    int act = ACT_NONE;
    int act_fp32 = 0;     // 1: evaluate erf-GELU in fp32 even for a bf16 output (default: packed fp16, see gemm.cu gelu_erf_h2)
    void* out = nullptr;  // bf16 or fp32 [*, ldo]
    int ldo = 0;
    int out_fp32 = 0;
    // ViT token assembly (patch-embed): GEMM row r = (image b, patch i) with G patches per image is written to token
    // row b * (G + 1) + 1 + i and gets rowbias[(1 + i), :] (the positional embedding) added.
    int remap_group = 0;
    const float* rowbias = nullptr;  // fp32 [G + 1, N]
    // Fused LayerNorm of the fp32 output rows (the residual GEMMs out_proj / fc2; replaces a separate LayerNorm launch
    // and its HBM read of the residual stream).  Every epilogue warp owns a 32-row x HALF_COLS sub-tile of the output:
    // while it adds bias + residual it also reduces each row's (mean, M2) over its columns (lane == row in the TMEM
    // layout, so this is thread-local), publishes them in ln_stats and bumps the 32-row strip's counter.  One tile LATER —
    // by then the other N tiles of that row band have normally been written too, so the wait is a formality — it merges
    // the strip's partial statistics (Chan's formula), reads its OWN sub-tile back (L2 hits) and writes
    // LayerNorm(row) * gamma + beta as bf16 (the next GEMM's A operand) and / or fp32 (may alias `out`: BERT's post-LN
    // rewrites the residual stream in place).  The work is spread evenly over all epilogue warps.  (The first version
    // let the LAST writer of a strip normalise all of it: that concentrates the work on the CTA pair that finishes a row
    // band last, which then starts its next tile late and is last again — fc2 went from 0.375 to 1.175 ms.)
    // Requirements: out_fp32, ldo == N, N % 128 == 0, N <= 1024, no token remap.
    const float* ln_gamma = nullptr;   // [N]; NULL = no fused LayerNorm
    const float* ln_beta = nullptr;    // [N]
    float ln_eps = 1e-5f;
    __nv_bfloat16* ln_out_bf16 = nullptr;   // [M, N] or NULL
    float* ln_out_f32 = nullptr;            // [M, N] or NULL
    float2* ln_stats = nullptr;             // [M, LN_MAX_PARTS] (mean, M2) per row and column part
    int* ln_counters = nullptr;             // int32 [ceil(M / 32)], all zero on entry
    int* ln_zero = nullptr;                 // int32 [ceil(M / 32)] zeroed by this launch (the counters of the NEXT fused
                                            // GEMM: out_proj and fc2 alternate between two arrays); may be set alone
    int ln_debug_skip = 0;                  // timing experiments only: publish / count but skip the normalisation
};

// A: bf16 [M, K] row-major with leading dimension lda (elements); W: bf16 [N, K] row-major (nn.Linear layout).
// Requirements: K % 64 == 0, N % 32 == 0, lda % 8 == 0.
void launch(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int M, int N, int K, const Epilogue& ep,
            int sm_count, cudaStream_t stream);

// ViT patch embedding straight from uint8 pixels (add_docs.py:129-134 + clip_utils.py:48-67 fused into the
// conv1 GEMM's operand load): out = epilogue( patches(img) x Wg^T ), where row r of the virtual A matrix is patch r of the
// uint8 HWC batch [n, S, S, 3] (image r / g^2, row-major in the g x g grid), normalised as ToTensor + Normalize, and
// Wg [N, patch_gather_k(patch)] is conv1.weight re-laid by kernels::patch_weight_rows.  No patch matrix exists in HBM:
// the gather warps of the GEMM read the image rows (16-byte coalesced), convert and write the swizzled smem A stage.
struct PatchGather {
    const uint8_t* img = nullptr;
    int n = 0, S = 0, patch = 0;
    float mean[3] = {0.f, 0.f, 0.f}, std[3] = {1.f, 1.f, 1.f};
};
bool patch_gather_supported(int S, int patch);   // S % patch == 0, patch even, 3*S % 16 == 0, S <= 224, grid >= 7
inline int patch_gather_kbpd(int patch) { return (3 * patch + 63) / 64; }
inline int patch_gather_k(int patch) { return patch * patch_gather_kbpd(patch) * 64; }
void launch_patch_embed(const PatchGather& pg, const __nv_bfloat16* Wg, int N, const Epilogue& ep, int sm_count,
                        cudaStream_t stream);

void configure();  // one-time cudaFuncSetAttribute calls

}  // namespace gemm
}  // namespace mb
