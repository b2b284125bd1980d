#!/bin/bash
# dev helper (gpurun --gpus 2): sharded search with the fused peer-store exchange + the 2-GPU bench line
export MARQO_B200_USE_PREBUILT=1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_score_exact_gpu.py -x -q -m gpu -k "two_gpus" 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus 2 --steps 10 --warmup 3 --quick > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -3 gpurun_out/r02_bench_n2.err | cut -c1-300
cut -c1-600 gpurun_out/r02_bench_n2.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2.json'))
print(d['value'], d['ms_per_step']); t=d['topk']; print({k:t[k] for k in ('value','ms_per_batch','scan_ms','merge_ms')}, t['config'], t['e2e'])
PY
