// Multi-head attention over packed QKV (head_dim 64), flash-style online softmax in fp32.
//   S >= 128: tcgen05 / TMEM kernel (attention_tc.cu) — 128 x 128 tiles, TMA-fed, thread-per-row softmax.
//   S <  128: 64-row warp-level kernel (attention.cu, mma.sync m16n8k16) — short sequences would leave most of a
//             128-wide tcgen05 tile masked.
#pragma once
#include "common.cuh"

namespace mb {
namespace attention {

enum Mask { MASK_NONE = 0, MASK_CAUSAL = 1, MASK_KEYLEN = 2 };

// qkv: bf16 [B*S, 3*W] rows = tokens, columns = [q | k | v], head h occupies columns h*64..h*64+63 of each part.
// out: bf16 [B*S, W].  kv_len: int32 [B] valid key count per sequence (MASK_KEYLEN only).
// Returns the number of kernels launched.
int launch(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int S, int W, int H, int mask, const int32_t* kv_len,
           cudaStream_t stream);

// tcgen05 / TMEM implementation (attention_tc.cu)
int launch_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int S, int W, int H, int mask, const int32_t* kv_len,
              cudaStream_t stream);

// One-shot kernel for 129 <= S <= 257 (attention_os.cu): all keys in one N = 256 tcgen05.mma, exact two-pass softmax in
// TMEM, P fed to the P V product straight from TMEM.  mask: MASK_NONE or MASK_KEYLEN.
bool os_supported(int S, int mask);
int launch_os(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int S, int W, int H, int mask, const int32_t* kv_len,
              cudaStream_t stream);

}  // namespace attention
}  // namespace mb
