#!/bin/bash
for sm in 1 2; do
  echo "== softmax mode $sm"
  MARQO_B200_ATTN_SOFTMAX=$sm python tools/attn_probe.py 2>&1 | tail -4
  MARQO_B200_ATTN_SOFTMAX=$sm python tools/attn_probe.py 256 77 768 12 1 2>&1 | tail -1
done
MARQO_B200_ATTN_SOFTMAX=2 python -m pytest tests/test_kernels_gpu.py tests/test_encoders_gpu.py -x -q -m gpu --durations=12 2>&1 | tail -25
