"""Model-cache bookkeeping of the vectorise shell (SURVEY §8 a1 / b: "size accounting is by declared model_size",
s2_inference.py:286-337,419-517) against THE REFERENCE's own functions run in the build container
(tests/golden/make_model_cache_golden.py): which models stay loaded on each device after every request, which requests
are refused, and the declared size of every model.  Host logic: runs without a GPU."""
import json
from pathlib import Path

import pytest

GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "model_cache_golden.json").read_text())


class _Dummy:
    closed = False

    def close(self):
        self.closed = True


@pytest.mark.parametrize("case", GOLD, ids=lambda c: f"cuda{c['cuda_threshold']}-cpu{c['cpu_threshold']}")
def test_model_cache_management_matches_reference(case, monkeypatch):
    from marqo_b200 import s2_inference as s2
    monkeypatch.setenv("MARQO_MAX_CUDA_MODEL_MEMORY", str(case["cuda_threshold"]))
    monkeypatch.setenv("MARQO_MAX_CPU_MODEL_MEMORY", str(case["cpu_threshold"]))
    monkeypatch.setattr(s2, "_load_model", lambda *a, **k: _Dummy())
    s2._available_models.clear()
    evicted = []
    try:
        for step in case["steps"]:
            name, props, device = step["name"], step["props"], step["device"]
            assert s2.get_model_size(name, props) == step["size"]
            key = s2._create_model_cache_key(name, device, props)
            before = dict(s2._available_models)
            try:
                s2._update_available_models(key, name, props, device, True)
                err = None
            except Exception as e:  # noqa: BLE001
                err = type(e).__name__
            assert err == step["error"], (name, err)
            assert [[k, v["model_size"]] for k, v in s2._available_models.items()] == step["loaded"], name
            evicted += [v["model"] for k, v in before.items() if k not in s2._available_models]
        assert evicted and all(m.closed for m in evicted)      # ejected handles give their device memory back
    finally:
        s2._available_models.clear()


def test_preprocessors_seam(monkeypatch):
    """s2_inference.py:193-235: add_documents asks for (model, preprocessors) once per batch."""
    from marqo_b200 import s2_inference as s2
    from marqo_b200.errors import InternalError

    class _Clip(_Dummy):
        def preprocess(self, image):
            return image

    monkeypatch.setattr(s2, "_load_model", lambda name, props, device, model_auth=None: _Clip())
    s2._available_models.clear()
    try:
        arch = {"any": 1}
        model, pre = s2.load_multimodal_model_and_get_preprocessors(
            "tiny-clip", {"name": "tiny-clip", "dimensions": 8, "type": "open_clip", "arch": arch}, device="cuda:0")
        assert set(pre) == {"image", "video", "audio", "text"} and pre["image"] == model.preprocess
        assert pre["video"] is None and pre["audio"] is None and pre["text"] is None
        model2, pre2 = s2.load_multimodal_model_and_get_preprocessors(
            "tiny-bert", {"name": "tiny-bert", "dimensions": 8, "type": "hf", "arch": arch}, device="cuda:0")
        assert pre2["image"] is None                       # text models have no image preprocessor
        assert len(s2._available_models) == 2
        with pytest.raises(InternalError):
            s2.load_multimodal_model_and_get_preprocessors("tiny-clip", {"dimensions": 8, "type": "open_clip"}, device=None)
    finally:
        s2._available_models.clear()
