#!/bin/bash
# dev helper (gpurun, 1 GPU): new kernels of this session — gather patch-embed, fused LayerNorm — tests + A/B timing
export MARQO_B200_USE_PREBUILT=1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "fused_layernorm or patch_embed or gemm" 2>&1 | tail -15
timeout 500 python -m pytest tests/test_encoders_gpu.py -x -q -m gpu --durations=6 2>&1 | tail -20
L14=open_clip/ViT-L-14/laion2b_s32b_b82k
echo "== ViT-L-14 b256 image: default (fused LN + gather)"
python tools/encoder_probe.py $L14 256 image 0 8 2>&1 | tail -2
echo "== no LN fusion"
MARQO_B200_NO_LN_FUSION=1 python tools/encoder_probe.py $L14 256 image 0 8 2>&1 | tail -2
echo "== no patch gather"
MARQO_B200_NO_PATCH_GATHER=1 python tools/encoder_probe.py $L14 256 image 0 8 2>&1 | tail -2
echo "== e5-large b64x512 fused / unfused"
python tools/encoder_probe.py hf/e5-large-v2 64 text 512 6 2>&1 | tail -2
MARQO_B200_NO_LN_FUSION=1 python tools/encoder_probe.py hf/e5-large-v2 64 text 512 6 2>&1 | tail -2
