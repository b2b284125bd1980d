#!/bin/bash
# dev helper (gpurun, 1 GPU): A/B in ONE run — cp.async residual prefetch (current lib) vs the previous epilogue (old-gemm lib);
# one-shot attention with the softmax warps issuing their own MMAs (MARQO_B200_ATTN_ISSUE=1) vs the MMA warp
export MARQO_B200_USE_PREBUILT=1
OLD=$PWD/marqo_b200/build/ab/libmarqo_b200_oldgemm.so
MARQO_B200_ATTN_ISSUE=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -3
for i in 1 2 3; do
python tools/attn_probe.py 256 257 1024 16 0 30 2>&1 | tail -1
MARQO_B200_ATTN_ISSUE=1 python tools/attn_probe.py 256 257 1024 16 0 30 2>&1 | tail -1
done
MARQO_B200_ATTN_ISSUE=1 python tools/attn_probe.py 256 197 768 12 0 30 2>&1 | tail -1
L14=open_clip/ViT-L-14/laion2b_s32b_b82k
for i in 1 2 3; do
echo "== current (cp.async residual)"; python tools/encoder_probe.py $L14 256 image 0 8 2>&1 | tail -2
echo "== old gemm"; MARQO_B200_LIB=$OLD python tools/encoder_probe.py $L14 256 image 0 8 2>&1 | tail -2
done
