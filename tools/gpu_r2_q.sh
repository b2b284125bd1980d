#!/bin/bash
# dev helper (gpurun, 1 GPU): TMA epilogue (residual in / fp32 out) — tests, memcheck, same-run A/B, ncu of out_proj
export MARQO_B200_USE_PREBUILT=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gemm or patch_embed or layernorm" 2>&1 | tail -3
timeout 400 python -m pytest tests/test_encoders_gpu.py tests/test_adapters_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu \
   -k "fused_layernorm or gemm_epilogues or (gemm_matches_torch and (333 or 200 or 300 or 50))" 2>&1 | tail -4
L14=open_clip/ViT-L-14/laion2b_s32b_b82k
for i in 1 2 3; do
echo "== TMA epilogue"; python tools/encoder_probe.py $L14 256 image 0 8 2>&1 | tail -2
echo "== manual epilogue"; MARQO_B200_GEMM_NO_TMA_EPILOGUE=1 python tools/encoder_probe.py $L14 256 image 0 8 2>&1 | tail -2
done
echo "== e5-large TMA / manual"
python tools/encoder_probe.py hf/e5-large-v2 64 text 512 6 2>&1 | tail -2
MARQO_B200_GEMM_NO_TMA_EPILOGUE=1 python tools/encoder_probe.py hf/e5-large-v2 64 text 512 6 2>&1 | tail -2
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 103 -c 1 -o gpurun_out/r02_gemm_outproj_tma \
    python tools/encoder_probe.py $L14 256 image 0 2 > /dev/null 2> gpurun_out/ncu_gemm.err
