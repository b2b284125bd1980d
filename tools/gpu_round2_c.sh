#!/bin/bash
# dev helper (gpurun, 1 GPU): A/B of the attention softmax schedules and the GEMM epilogue variants
for sm in 0 1 2; do
  echo "== softmax mode $sm"
  MARQO_B200_ATTN_SOFTMAX=$sm python tools/attn_probe.py 2>&1 | tail -4
  MARQO_B200_ATTN_SOFTMAX=$sm python tools/attn_probe.py 256 77 768 12 1 2>&1 | tail -1
done
echo "== mma.sync short kernel"
MARQO_B200_ATTN_SHORT=mma python tools/attn_probe.py 256 50 768 12 0 2>&1 | tail -1
MARQO_B200_ATTN_SHORT=mma python tools/attn_probe.py 256 77 768 12 1 2>&1 | tail -1
echo "== ViT-L-14 b256 image: 16-warp epilogue (default)"
python tools/encoder_probe.py open_clip/ViT-L-14/laion2b_s32b_b82k 256 image 0 6 2>&1 | tail -2
echo "== ViT-L-14 b256 image: 8-warp epilogue"
MARQO_B200_GEMM_EPI8=1 python tools/encoder_probe.py open_clip/ViT-L-14/laion2b_s32b_b82k 256 image 0 6 2>&1 | tail -2
echo "== ViT-L-14 b256, softmax mode 0 / 2"
MARQO_B200_ATTN_SOFTMAX=0 python tools/encoder_probe.py open_clip/ViT-L-14/laion2b_s32b_b82k 256 image 0 6 2>&1 | tail -2
MARQO_B200_ATTN_SOFTMAX=2 python tools/encoder_probe.py open_clip/ViT-L-14/laion2b_s32b_b82k 256 image 0 6 2>&1 | tail -2
echo "== e5-large b64x512; ViT-B-32 b256 image / text"
python tools/encoder_probe.py hf/e5-large-v2 64 text 512 5 2>&1 | tail -2
python tools/encoder_probe.py open_clip/ViT-B-32/laion2b_s34b_b79k 256 image 0 6 2>&1 | tail -2
python tools/encoder_probe.py open_clip/ViT-B-32/laion2b_s34b_b79k 256 text 77 6 2>&1 | tail -2
python tools/score_probe.py 2>&1 | head -1 | cut -c1-400
