"""Dev probe: the drop-in call itself — B200OpenCLIP.encode(list of 256 decoded uint8 images) -> np.ndarray, wall clock,
i.e. what Marqo's vectorise() pays per batch on top of the kernels (staging, H2D, D2H).  Not a bench line."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from marqo_b200.loaders import B200OpenCLIP  # noqa: E402
from marqo_b200 import model_registry as R  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "open_clip/ViT-L-14/laion2b_s32b_b82k"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
props = dict(R.get_model_properties(name), random_init=1234, max_batch=n)
model = B200OpenCLIP(device="cuda:0", model_properties=props)
model.load()
g = torch.Generator().manual_seed(0)
imgs = [torch.randint(0, 256, (224, 224, 3), dtype=torch.uint8, generator=g) for _ in range(n)]
ts = []
for it in range(6):
    t0 = time.perf_counter()
    out = model.encode(imgs, default="image", normalize=True)
    ts.append((time.perf_counter() - t0) * 1e3)
dev_ms, launches = model.model.last_timing()
print(json.dumps({"model": name, "n": n, "call_ms": ts, "call_ms_med": float(np.median(ts[2:])),
                  "images_per_s": n / (float(np.median(ts[2:])) / 1e3), "device_ms_last": dev_ms,
                  "finite": bool(np.isfinite(out).all()), "shape": list(out.shape)}))
