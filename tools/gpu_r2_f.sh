#!/bin/bash
# dev helper (gpurun, 1 GPU): one-shot attention kernel + fused LayerNorm (acq_rel counters) — tests + A/B timing
export MARQO_B200_USE_PREBUILT=1
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention or fused_layernorm" 2>&1 | tail -15
echo "== attention probe: one-shot (default) vs block kernel"
python tools/attn_probe.py 256 257 1024 16 0 30 2>&1 | tail -1
MARQO_B200_ATTN_NO_ONESHOT=1 python tools/attn_probe.py 256 257 1024 16 0 30 2>&1 | tail -1
python tools/attn_probe.py 256 197 768 12 0 30 2>&1 | tail -1
MARQO_B200_ATTN_NO_ONESHOT=1 python tools/attn_probe.py 256 197 768 12 0 30 2>&1 | tail -1
L14=open_clip/ViT-L-14/laion2b_s32b_b82k
echo "== ViT-L-14 b256 image: default (fused LN + one-shot attention)"
python tools/encoder_probe.py $L14 256 image 0 8 2>&1 | tail -2
echo "== no LN fusion"
MARQO_B200_NO_LN_FUSION=1 python tools/encoder_probe.py $L14 256 image 0 8 2>&1 | tail -2
echo "== no LN fusion, block attention"
MARQO_B200_NO_LN_FUSION=1 MARQO_B200_ATTN_NO_ONESHOT=1 python tools/encoder_probe.py $L14 256 image 0 8 2>&1 | tail -2
