"""Generates tests/golden/model_cache_golden.json by running THE REFERENCE's own model-cache management
(marqo.s2_inference.s2_inference: _update_available_models :286-337, _validate_model_into_device :419-457,
_check_memory_threshold_for_model :460-501, get_model_size :504-517) with `_load_model` replaced by a dummy, so only the
bookkeeping runs: which models stay loaded on a device after each request, and which requests are refused.

    python tests/golden/make_model_cache_golden.py
"""
import json
import os
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import _reference_import  # noqa: E402

_reference_import.install()
import torch  # noqa: E402
import torchaudio  # noqa: E402

if not hasattr(torchaudio, "set_audio_backend"):
    torchaudio.set_audio_backend = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None      # no GPU in the build container: the bookkeeping is what is recorded
torch.cuda.empty_cache = lambda *a, **k: None

import marqo.s2_inference.s2_inference as s2  # noqa: E402

s2._load_model = lambda *a, **k: object()

REQUESTS = [
    # (model name, properties, device)
    ("open_clip/ViT-B-32/laion2b_s34b_b79k", {"name": "ViT-B-32", "dimensions": 512, "type": "open_clip"}, "cuda:0"),
    ("open_clip/ViT-L-14/laion2b_s32b_b82k", {"name": "ViT-L-14", "dimensions": 768, "type": "open_clip"}, "cuda:0"),
    ("hf/e5-base-v2", {"name": "intfloat/e5-base-v2", "dimensions": 768, "type": "hf", "tokens": 512}, "cuda:0"),
    ("open_clip/ViT-B-32/laion2b_s34b_b79k", {"name": "ViT-B-32", "dimensions": 512, "type": "open_clip"}, "cuda:0"),  # renew
    ("custom-big", {"name": "x", "dimensions": 64, "type": "hf", "model_size": 2.5}, "cuda:0"),                       # evicts
    ("hf/e5-base-v2", {"name": "intfloat/e5-base-v2", "dimensions": 768, "type": "hf", "tokens": 512}, "cuda:1"),     # other device
    ("custom-huge", {"name": "y", "dimensions": 64, "type": "hf", "model_size": 9}, "cuda:0"),                        # > threshold
    ("open_clip/ViT-bigG-14/x", {"name": "ViT-bigG-14", "dimensions": 1280, "type": "open_clip"}, "cpu"),             # name mapping
    ("unknown-type", {"name": "z", "dimensions": 8, "type": "mystery"}, "cpu"),                                       # default size
    ("custom-mid", {"name": "w", "dimensions": 8, "type": "hf", "model_size": 3.5}, "cuda:0"),                        # evicts all
]
out = []
for cuda_thr, cpu_thr in ((4, 4), (2, 7)):
    os.environ["MARQO_MAX_CUDA_MODEL_MEMORY"] = str(cuda_thr)
    os.environ["MARQO_MAX_CPU_MODEL_MEMORY"] = str(cpu_thr)
    s2._available_models.clear()
    steps = []
    for name, props, device in REQUESTS:
        key = s2._create_model_cache_key(name, device, props)
        try:
            s2._update_available_models(key, name, props, device, True)
            err = None
        except Exception as e:  # noqa: BLE001
            err = type(e).__name__
        steps.append({"name": name, "props": props, "device": device, "error": err,
                      "size": s2.get_model_size(name, props),
                      "loaded": [[k, v["model_size"]] for k, v in s2._available_models.items()]})
    out.append({"cuda_threshold": cuda_thr, "cpu_threshold": cpu_thr, "steps": steps})
(HERE / "model_cache_golden.json").write_text(json.dumps(out, indent=1))
for st in out[0]["steps"]:
    print(st["name"], st["device"], st["size"], st["error"], len(st["loaded"]))
