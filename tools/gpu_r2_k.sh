#!/bin/bash
# dev helper (gpurun, 1 GPU): packed-half GELU — accuracy tests + A/B timing
export MARQO_B200_USE_PREBUILT=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gemm or attention" 2>&1 | tail -3
timeout 400 python -m pytest tests/test_encoders_gpu.py -x -q -m gpu 2>&1 | tail -3
L14=open_clip/ViT-L-14/laion2b_s32b_b82k
for i in 1 2; do
echo "== fp16x2 GELU"; python tools/encoder_probe.py $L14 256 image 0 8 2>&1 | tail -2
echo "== fp32 GELU"; MARQO_B200_GELU_FP32=1 python tools/encoder_probe.py $L14 256 image 0 8 2>&1 | tail -2
done
python tools/attn_probe.py 256 257 1024 16 0 30 2>&1 | tail -1
