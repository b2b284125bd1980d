#!/bin/bash
# dev helper (gpurun, 1 GPU): fused LayerNorm modes A/B + the whole GPU suite with durations
export MARQO_B200_USE_PREBUILT=1
mkdir -p gpurun_out
L14=open_clip/ViT-L-14/laion2b_s32b_b82k
timeout 200 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "fused_layernorm" 2>&1 | tail -2
for mode in 0 2 1 0 2; do
echo "== LN fusion mode $mode"; MARQO_B200_LN_FUSION=$mode python tools/encoder_probe.py $L14 256 image 0 8 2>&1 | tail -2
done
(time timeout 1200 python -m pytest tests -q -m gpu --durations=40 2>&1 | tail -60) > gpurun_out/r02_pytest_gpu_full.txt 2>&1
tail -50 gpurun_out/r02_pytest_gpu_full.txt
