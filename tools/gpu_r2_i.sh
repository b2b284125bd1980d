#!/bin/bash
# dev helper (gpurun, 1 GPU): fused LayerNorm modes — tests + A/B timing in one run (same box, same clocks)
export MARQO_B200_USE_PREBUILT=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "fused_layernorm" 2>&1 | tail -3
MARQO_B200_LN_FUSION=2 timeout 400 python -m pytest tests/test_encoders_gpu.py -x -q -m gpu -k "tiny or vit_b_32 or e5_base or graphs" 2>&1 | tail -3
L14=open_clip/ViT-L-14/laion2b_s32b_b82k
for i in 1 2; do
for mode in 0 2 1; do
echo "== LN fusion mode $mode"; MARQO_B200_LN_FUSION=$mode python tools/encoder_probe.py $L14 256 image 0 8 2>&1 | tail -2
done
done
echo "== e5-large b64x512 mode 0 / 2"
MARQO_B200_LN_FUSION=0 python tools/encoder_probe.py hf/e5-large-v2 64 text 512 6 2>&1 | tail -2
MARQO_B200_LN_FUSION=2 python tools/encoder_probe.py hf/e5-large-v2 64 text 512 6 2>&1 | tail -2
