#!/bin/bash
# dev helper (gpurun, 1 GPU): where does the fused LayerNorm's time go + ncu of the one-shot attention kernel
export MARQO_B200_USE_PREBUILT=1
mkdir -p gpurun_out
L14=open_clip/ViT-L-14/laion2b_s32b_b82k
echo "== fused LN, normalisation skipped (counters only)"
MARQO_B200_LN_DEBUG_SKIP=1 python tools/encoder_probe.py $L14 256 image 0 6 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:attention_os_kernel -s 3 -c 1 -o gpurun_out/r02_attn_os \
    python tools/attn_probe.py 256 257 1024 16 0 6 > /dev/null 2> gpurun_out/ncu_attn_os.err
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 12 -c 2 -o gpurun_out/r02_gemm_ln \
    python tools/encoder_probe.py $L14 256 image 0 1 > /dev/null 2> gpurun_out/ncu_gemm_ln.err
ls -la gpurun_out | tail -5
