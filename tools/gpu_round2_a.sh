#!/bin/bash
# dev helper (gpurun): full GPU suite + attention kernel probes
python -m pytest tests -x -q -m gpu 2>&1 | tail -40
python tools/attn_probe.py 2>&1 | tail -6
MARQO_B200_ATTN_SHORT=mma python tools/attn_probe.py 256 50 768 12 0 2>&1 | tail -2
MARQO_B200_ATTN_SHORT=mma python tools/attn_probe.py 256 77 768 12 1 2>&1 | tail -2
python tools/attn_probe.py 256 77 768 12 1 2>&1 | tail -2
