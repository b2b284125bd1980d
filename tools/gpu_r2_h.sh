#!/bin/bash
# dev helper (gpurun, 1 GPU): one-shot attention iteration + GEMM epilogue check — tests + probes
export MARQO_B200_USE_PREBUILT=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention or gemm" 2>&1 | tail -5
python tools/attn_probe.py 256 257 1024 16 0 30 2>&1 | tail -1
python tools/attn_probe.py 256 197 768 12 0 30 2>&1 | tail -1
python tools/encoder_probe.py open_clip/ViT-L-14/laion2b_s32b_b82k 256 image 0 8 2>&1 | tail -2
