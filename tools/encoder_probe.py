"""Dev probe: device-resident encode timing for one model (not a bench line)."""
import sys, json, time
import numpy as np
import torch
sys.path.insert(0, ".")
from marqo_b200.engine import Encoder
from marqo_b200 import model_registry as R, weights as Wt

name = sys.argv[1] if len(sys.argv) > 1 else "open_clip/ViT-B-32/laion2b_s34b_b79k"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
mode = sys.argv[3] if len(sys.argv) > 3 else "image"
seq = int(sys.argv[4]) if len(sys.argv) > 4 else 0
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5
props = R.get_model_properties(name)
arch = props["arch"]
torch.cuda.set_device(0)
t0 = time.time()
if props["type"] == R.TYPE_OPEN_CLIP:
    if mode == "image":
        arch = dict(arch); arch["text"] = None
    else:
        arch = dict(arch); arch["vision"] = None
    sd = Wt.random_clip_weights(arch, 1234)
    enc = Encoder("clip", arch, sd, max_batch=B)
else:
    sd = Wt.random_bert_weights(arch, 1234)
    enc = Encoder("bert", arch, sd, max_batch=B)
print("load s", time.time() - t0, flush=True)
E = enc.embed_dim
out = torch.empty(B, E, dtype=torch.float32, device="cuda")
g = torch.Generator(device="cuda").manual_seed(0)
if mode == "image":
    img = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, device="cuda", generator=g)
    run = lambda: enc.encode_images_u8_device(img.data_ptr(), B, 224, 224, out.data_ptr())
    v = arch["vision"]; S = (224 // v["patch"]) ** 2 + 1
    w, L, mlp = v["width"], v["layers"], v["mlp"]
    flops = B * (L * (2 * S * (4 * w * w + 2 * w * mlp) + 4 * S * S * w) + 2 * (S - 1) * 3 * v["patch"] ** 2 * w)
else:
    t = arch["text"] if props["type"] == R.TYPE_OPEN_CLIP else arch
    S = seq or t.get("ctx", 128)
    ids = torch.randint(1, t["vocab"] - 2, (B, S), dtype=torch.int32, device="cuda", generator=g)
    ids[:, -1] = t["vocab"] - 1
    run = lambda: enc.encode_tokens_device(ids.data_ptr(), None, B, S, out.data_ptr())
    w, L, mlp = t["width"], t["layers"], t["mlp"]
    flops = B * L * (2 * S * (4 * w * w + 2 * w * mlp) + 4 * S * S * w)
torch.cuda.synchronize()
res = []
for i in range(iters):
    run()
    res.append(enc.last_timing())
ms = sorted(r[0] for r in res[1:])
med = ms[len(ms) // 2]
print(json.dumps({"model": name, "mode": mode, "B": B, "S": S, "ms_med": med, "ms_min": ms[0], "items_per_s": B / med * 1e3,
                  "TFLOPs": flops / med / 1e9, "launches": res[-1][1], "finite": bool(torch.isfinite(out).all())}))
enc.set_profiling(True)
for i in range(3):
    run()
torch.cuda.synchronize()
pr = enc.profile()
print(json.dumps({k: (v / 3 if k.endswith("_ms") else v // 3) for k, v in pr.items()}))
