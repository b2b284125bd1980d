"""Dev probe: wall-clock latency of small encode calls through the HOST entry points (the single-query path of a
search request): first call eager, second captured into a CUDA graph, later ones replayed.  Not a bench line."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from marqo_b200 import model_registry as R, weights as Wt  # noqa: E402
from marqo_b200.engine import Encoder  # noqa: E402

for name, kind, shapes in (("hf/e5-base-v2", "bert", [(1, 16), (1, 128), (8, 128)]),
                           ("open_clip/ViT-B-32/laion2b_s34b_b79k", "clip", [(1, 77), (8, 77)])):
    arch = R.get_model_properties(name)["arch"]
    if kind == "clip":
        sd = Wt.random_clip_weights(arch, 1234)
    else:
        sd = Wt.random_bert_weights(arch, 1234)
    enc = Encoder(kind, arch, sd, max_batch=16)
    rng = np.random.default_rng(0)
    for n, S in shapes:
        ids = rng.integers(1000, 20000, size=(n, S)).astype(np.int32)
        if kind == "clip":
            ids[:, 0], ids[:, -1] = 49406, 49407
        mask = np.ones_like(ids)
        ts = []
        for it in range(24):
            t0 = time.perf_counter()
            out = enc.encode_tokens(ids, mask if kind == "bert" else None)
            ts.append((time.perf_counter() - t0) * 1e3)
        print(json.dumps({"model": name, "n": n, "S": S, "first_ms": ts[0], "second_ms": ts[1],
                          "replay_ms_med": float(np.median(ts[4:])), "launches": enc.last_timing()[1],
                          "device_ms": enc.last_timing()[0], "finite": bool(np.isfinite(out).all())}))
    if kind == "clip":
        img = rng.integers(0, 256, size=(1, 224, 224, 3), dtype=np.uint8)
        ts = []
        for it in range(16):
            t0 = time.perf_counter()
            out = enc.encode_images_u8(img)
            ts.append((time.perf_counter() - t0) * 1e3)
        print(json.dumps({"model": name, "image_n": 1, "first_ms": ts[0], "second_ms": ts[1],
                          "replay_ms_med": float(np.median(ts[4:])), "device_ms": enc.last_timing()[0]}))
    enc.close()
