#!/bin/bash
# dev helper (gpurun, 1 GPU): bench line + launch list + ncu captures of the kernels that changed in round 2
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -c 3000 gpurun_out/r02_bench_n1.err | tail -15
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2>> gpurun_out/r02_bench_n1.err
# launch list of a short headline run (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 3 --warmup 3 --quick --skip-topk --skip-cpu-baseline > /dev/null 2> gpurun_out/ncu_launch.err
# full captures: attention (S=257), fc1-shaped GEMM with the 16-warp epilogue, out_proj, score scan + merge
ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 30 -c 2 -o gpurun_out/r02_attn \
    python tools/encoder_probe.py open_clip/ViT-L-14/laion2b_s32b_b82k 256 image 0 2 > /dev/null 2> gpurun_out/ncu_attn.err
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 100 -c 8 -o gpurun_out/r02_gemm \
    python tools/encoder_probe.py open_clip/ViT-L-14/laion2b_s32b_b82k 256 image 0 2 > /dev/null 2> gpurun_out/ncu_gemm.err
ncu --set full --clock-control none --import-source on -k regex:"scan_kernel|merge_kernel|finalize_kernel" -s 12 -c 6 -o gpurun_out/r02_score \
    python tools/score_probe.py > /dev/null 2> gpurun_out/ncu_score.err
ls -la gpurun_out | tail -20
