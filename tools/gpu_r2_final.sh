#!/bin/bash
# dev helper (gpurun, 1 GPU): what the driver runs at round end (GPU suite, smoke, bench, reference arm) + launch list +
# ncu captures of the final build
export MARQO_B200_USE_PREBUILT=1
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4) 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -c 600 gpurun_out/r02_bench_n1.err | tail -3
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2>> gpurun_out/r02_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/r02_launches_final.csv \
    python bench.py --steps 3 --warmup 3 --quick --skip-topk --skip-cpu-baseline > /dev/null 2> gpurun_out/ncu_launch.err
L14=open_clip/ViT-L-14/laion2b_s32b_b82k
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 101 -c 4 -o gpurun_out/r02_gemm_final \
    python tools/encoder_probe.py $L14 256 image 0 2 > /dev/null 2> gpurun_out/ncu_gemm.err
cut -c1-400 gpurun_out/r02_bench_n1.json
