#!/bin/bash
# dev helper (gpurun, 1 GPU): cp.async residual prefetch — GEMM tests, encoder parity, timing, memcheck of the new kernels
export MARQO_B200_USE_PREBUILT=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gemm" 2>&1 | tail -3
timeout 400 python -m pytest tests/test_encoders_gpu.py -x -q -m gpu 2>&1 | tail -3
L14=open_clip/ViT-L-14/laion2b_s32b_b82k
python tools/encoder_probe.py $L14 256 image 0 8 2>&1 | tail -2
python tools/encoder_probe.py hf/e5-large-v2 64 text 512 6 2>&1 | tail -2
python tools/encoder_probe.py $L14 256 image 0 8 2>&1 | tail -2
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu \
   -k "patch_embed or fused_layernorm or (attention_matches_torch and (257 or 197 or 200 or 255 or 256)) or gemm_epilogues or (gemm_matches_torch and 333)" 2>&1 | tail -8
