"""CPU tests: host logic of the boundary, the C-ABI library surface, the score oracle, the distributed host path."""
import ctypes
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


# ------------------------------------------------------------------------------------------------ C ABI surface
def test_library_exports_every_declared_symbol(native_lib):
    header = (ROOT / "include" / "marqo_b200.h").read_text()
    declared = set(re.findall(r"^\s*(?:int|const char\*)\s+(b200_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert len(declared) >= 35
    for name in sorted(declared):
        assert hasattr(native_lib, name), f"libmarqo_b200.so does not export {name}"
    from marqo_b200 import _native
    assert declared == set(_native.EXPORTED_SYMBOLS), declared ^ set(_native.EXPORTED_SYMBOLS)
    assert native_lib.b200_abi_version() == 1


def test_no_gpu_means_loud_failure_not_fallback(native_lib):
    """On a box without a B200 every compute entry point must fail with B200_ERR_NO_DEVICE."""
    from marqo_b200 import _native
    if _native.device_count() > 0:
        pytest.skip("a GPU is present")
    from marqo_b200.engine import RowStore, debug_gemm
    with pytest.raises(_native.NativeError) as ei:
        RowStore(64)
    assert ei.value.code == _native.ERR_NO_DEVICE and "no CPU fallback" in ei.value.message
    with pytest.raises(_native.NativeError) as ei:
        debug_gemm(np.zeros((4, 64), np.float32), np.zeros((32, 64), np.float32))
    assert ei.value.code == _native.ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    for p in (ROOT / "marqo_b200").rglob("*.py"):
        txt = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{p} imports the oracle"
    for p in (ROOT / "marqo_b200" / "csrc").glob("*.cu*"):
        txt = p.read_text()
        assert "libscore_oracle" not in txt and not re.search(r'#include\s*[<"][^>"]*oracle', txt), p


def test_topk_merge_host(native_lib):
    from marqo_b200.engine import topk_merge
    doc = np.array([[[5, 1, -1]], [[7, 2, 9]]], np.int32)          # 2 shards, 1 query, k = 3
    row = doc.copy()
    score = np.array([[[0.9, 0.5, -np.inf]], [[0.9, 0.8, 0.1]]])
    d, r, s = topk_merge(doc, row, score)
    assert list(d[0]) == [5, 7, 2] and list(s[0]) == [0.9, 0.9, 0.8]   # tie 0.9: lower doc id first


def test_fuse_vectors_c(native_lib):
    import ctypes as C
    v = np.array([[1.0, 2.0], [1.0, 2.0]])
    w = np.array([1.0, 2.0])
    out = np.zeros(2)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert native_lib.b200_fuse_vectors(p(v), p(w), 2, 2, 0, p(out)) == 0
    assert out.tolist() == [1.5, 3.0]


# ------------------------------------------------------------------------------------------------ score oracle
def test_score_oracle_known_answers(score_oracle):
    so = score_oracle
    e = np.eye(4, 64, dtype=np.float32)
    doc, row, score = so.search(e[:1], e, 4)
    assert list(doc[0]) == [0, 1, 2, 3]
    assert score[0, 0] == 1.0 and score[0, 1] == 0.5          # closeness 1/(1+(1-dot)); identical vector -> 1.0
    assert so.closeness(1.0, "angular") == 1.0 and abs(so.closeness(0.0, "angular") - 1 / (1 + np.pi / 2)) < 1e-15
    # max over chunks + deleted rows + tie order
    c = np.stack([e[0], e[1], e[0], e[2]])
    d, r, s = so.search(e[:1], c, 3, doc_of_row=np.array([3, 1, 0, -1], np.int32))
    assert list(d[0]) == [0, 3, 1] and list(r[0]) == [2, 0, 1]
    # empty corpus
    d, r, s = so.search(e[:1], np.zeros((0, 64), np.float32), 2)
    assert (d == -1).all() and np.isneginf(s).all()


def test_score_oracle_matches_numpy(score_oracle):
    rng = np.random.default_rng(1)
    c = rng.standard_normal((3000, 128)).astype(np.float32)
    q = rng.standard_normal((5, 128)).astype(np.float32)
    d, r, s = score_oracle.search(q, c, 7, "dotproduct")
    ch, qh = c.astype(np.float16).astype(np.float64), q.astype(np.float16).astype(np.float64)
    dots = qh @ ch.T
    order = np.argsort(-dots, axis=1, kind="stable")[:, :7]
    np.testing.assert_array_equal(d, order)
    np.testing.assert_allclose(s, np.take_along_axis(dots, order, 1), rtol=1e-12)


# ------------------------------------------------------------------------------------------------ vectorise shell
def test_vectorise_batching_order_and_types(monkeypatch):
    """Mirrors tests/s2_inference/test_vectorise.py:88-141 of the reference (model injected into _available_models)."""
    from marqo_b200 import s2_inference as s2

    class M:
        def __init__(self):
            self.batches = []

        def encode(self, content, normalize=True, **kw):
            items = [content] if isinstance(content, str) else list(content)
            self.batches.append(list(items))
            return np.asarray([[float(len(x)), 1.0] for x in items], np.float32)

    props = {"name": "m", "dimensions": 2, "type": "hf", "tokens": 128, "arch": {}}
    vp = s2.validate_model_properties("m", props)
    key = s2._create_model_cache_key("m", "cuda:0", vp)
    m = M()
    s2._available_models[key] = {"model": m, "most_recently_used_time": 0, "model_size": 1}
    try:
        monkeypatch.setenv("MARQO_MAX_VECTORISE_BATCH_SIZE", "4")
        content = ["a" * i for i in range(1, 11)]
        out = s2.vectorise("m", content, model_properties=props, device="cuda:0")
        assert [len(b) for b in m.batches] == [4, 4, 2]
        assert out == [[float(i), 1.0] for i in range(1, 11)]
        assert isinstance(out[0][0], float)
        assert s2.vectorise("m", "xyz", model_properties=props, device="cuda:0") == [[3.0, 1.0]]
        monkeypatch.setenv("MARQO_MAX_VECTORISE_BATCH_SIZE", "0")
        from marqo_b200.errors import ConfigurationError
        with pytest.raises(ConfigurationError):
            s2.vectorise("m", content, model_properties=props, device="cuda:0")
    finally:
        s2._available_models.pop(key, None)


def test_validate_model_properties_and_registry():
    from marqo_b200 import s2_inference as s2, model_registry as R
    from marqo_b200.errors import InvalidModelPropertiesError, UnknownModelError
    p = s2.validate_model_properties("hf/e5-base-v2", None)
    assert p["dimensions"] == 768 and p["tokens"] == 512 and p["text_query_prefix"] == "query: "
    p = s2.validate_model_properties("open_clip/ViT-B-32/laion2b_s34b_b79k", None)
    assert p["dimensions"] == 512 and p["arch"]["vision"]["patch"] == 32
    assert R.MODELS["open_clip/ViT-L-14/openai"]["arch"]["act"] == "quickgelu"
    with pytest.raises(UnknownModelError):
        s2.validate_model_properties("nope/model", None)
    with pytest.raises(InvalidModelPropertiesError):
        s2.validate_model_properties("x", {"type": "hf"})
    with pytest.raises(InvalidModelPropertiesError):
        s2.validate_model_properties("x", {"type": "sbert", "dimensions": 3})
    p = s2.validate_model_properties("custom", {"type": "hf", "dimensions": 768, "name": "intfloat/e5-base-v2"})
    assert p["type"] == R.TYPE_HF and p["arch"]["layers"] == 12 and p["tokens"] == 128


def test_device_strings_and_no_cpu_fallback():
    from marqo_b200 import s2_inference as s2
    from marqo_b200.errors import ModelLoadError
    assert s2._validate_device("cuda") == 0 and s2._validate_device("cuda:3") == 3
    with pytest.raises(ModelLoadError):
        s2._validate_device("cpu")


def test_convert_vectorized_output():
    import torch
    from marqo_b200.s2_inference import _convert_vectorized_output
    assert _convert_vectorized_output(np.array([1.0, 2.0], np.float32)) == [[1.0, 2.0]]
    assert _convert_vectorized_output(torch.tensor([[1.0, 2.0]])) == [[1.0, 2.0]]
    assert _convert_vectorized_output([np.array([1.0]), np.array([2.0])]) == [[1.0], [2.0]]
    assert _convert_vectorized_output([[1.0, 2.0]]) == [[1.0, 2.0]]
    with pytest.raises(TypeError):
        _convert_vectorized_output("nope")


def test_infer_modality():
    from marqo_b200.s2_inference import Modality, infer_modality
    assert infer_modality("a plain sentence") == Modality.TEXT
    assert infer_modality("https://example.com/cat.jpg") == Modality.IMAGE
    assert infer_modality("https://example.com/clip.mp4") == Modality.VIDEO
    assert infer_modality("https://example.com/a.wav") == Modality.AUDIO
    assert infer_modality(b"bytes") == Modality.TEXT


# ------------------------------------------------------------------------------------------------ B2 adapter (no GPU parts)
def test_gpu_tensor_index_query_discrimination():
    from marqo_b200.gpu_tensor_index import GpuTensorIndex
    from marqo_b200.errors import VespaError
    ix = GpuTensorIndex()
    yql = ("select * from s1 where (({targetHits:10, approximate:False, hnsw.exploreAdditionalHits:0}"
           "nearestNeighbor(marqo__embeddings_title, marqo__query_embedding)) OR "
           "({targetHits:10, approximate:False, hnsw.exploreAdditionalHits:0}"
           "nearestNeighbor(marqo__embeddings_body, marqo__query_embedding)))")
    qf = {"marqo__query_embedding": [0.0] * 64, "marqo__embeddings_title": 1}
    assert ix._is_tensor_query(yql, "embedding_similarity", qf)
    assert not ix._is_tensor_query(yql, "bm25", qf)
    assert not ix._is_tensor_query(yql + " AND (price > 3)", "embedding_similarity", qf)
    assert not ix._is_tensor_query(yql, "embedding_similarity", dict(qf, marqo__mult_weights_global={"a": 1}))
    with pytest.raises(VespaError):
        ix.query("select * from s1 where userQuery()", ranking="bm25", query_features={})

    class Delegate:
        def query(self, yql, **kw):
            return ("delegated", kw["ranking"])
    assert GpuTensorIndex(delegate=Delegate()).query("select * from s1 where userQuery()", ranking="bm25",
                                                      query_features={}) == ("delegated", "bm25")
    # empty schema answers with zero hits, full coverage
    res = ix.query(yql, hits=5, ranking="embedding_similarity", model_restrict="s1", query_features=qf)
    assert res.hits == [] and res.root.coverage.coverage == 100


# ------------------------------------------------------------------------------------------------ multi-process (gloo)
def test_shard_bounds():
    from marqo_b200.distributed import partition, shard_bounds
    assert [shard_bounds(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [shard_bounds(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    items = list(range(11))
    assert sum((list(partition(items, r, 3)) for r in range(3)), []) == items     # order preserved, nothing lost


_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["REPO_ROOT"])
from marqo_b200 import build
build.build_oracle()
from oracle import score_oracle as so
from marqo_b200.distributed import ShardedRowStore, allgather_topk, shard_bounds
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{os.environ['MASTER_PORT']}", rank=rank, world_size=world)
rng = np.random.default_rng(0)
corpus = rng.standard_normal((2001, 64)).astype(np.float32)
corpus[1500] = corpus[3]                    # a cross-shard tie
q = rng.standard_normal((5, 64)).astype(np.float32); q[0] = corpus[3]
lo, hi = shard_bounds(len(corpus), rank, world)
d, r, s = so.search(q, corpus[lo:hi], 6, "dotproduct")          # stand-in for the per-shard GPU scan
d = np.where(d >= 0, d + lo, -1); r = np.where(r >= 0, r + lo, -1)
md, mr, ms = allgather_topk(d, r, s)
ed, er, es = so.search(q, corpus, 6, "dotproduct")
assert (md == ed).all() and (mr == er).all() and (ms == es).all(), (rank, md, ed)
assert md[0, 0] == 3 and md[0, 1] == 1500
# the same through ShardedRowStore (host exchange mode: ONE packed all-gather per query block)
class _Store:
    dim = 64
    def __init__(self, rows): self.rows, self.off = rows, 0
    def set_doc_offset(self, o): self.off = o
    def add(self, v, ids): pass
    def search(self, q, k):
        d, r, s = so.search(q, self.rows, k, "dotproduct")
        return np.where(d >= 0, d + self.off, -1), r, s
sh = ShardedRowStore(_Store(corpus[lo:hi]), rank, world)
assert sh.mode == "host"
sh.add_local(None, None, lo)
sd, sr, ss = sh.search(q, 6)
assert (sd == ed).all() and (ss == es).all()
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_sharded_search_allgather_merge_gloo_world2(native_lib, score_oracle, tmp_path):
    """N>1 path on CPU: 2 processes, gloo, per-shard exact lists -> one all-gather -> merge == global exact top-k."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   REPO_ROOT=str(ROOT), OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out
        assert "ok" in out


# ------------------------------------------------------------------------------------------------ model-cache concurrency
def test_concurrent_model_load_is_rejected_like_the_reference(monkeypatch):
    """tests/s2_inference/test_automatic_model_ejection_and_concurrency.py:135-254 of the reference: a request that
    needs to load a model while another load holds the lock fails fast with ModelCacheManagementError."""
    import threading
    import time
    from marqo_b200 import loaders, s2_inference as s2
    from marqo_b200.errors import ModelCacheManagementError

    started, release = threading.Event(), threading.Event()

    class SlowModel:
        def __init__(self, device=None, model_properties=None, model_auth=None):
            pass

        def load(self):
            started.set()
            release.wait(timeout=20)

        def encode(self, content, normalize=True, **kw):
            items = [content] if isinstance(content, str) else list(content)
            return np.ones((len(items), 4), np.float32)

    monkeypatch.setattr(loaders, "get_model_loader", lambda name, props: SlowModel)
    s2.clear_loaded_models()
    props_a = {"name": "slow-a", "dimensions": 4, "type": "hf", "arch": {}}
    props_b = {"name": "slow-b", "dimensions": 4, "type": "hf", "arch": {}}
    result = {}

    def first():
        result["a"] = s2.vectorise("slow-a", ["x"], model_properties=props_a, device="cuda:0")

    t = threading.Thread(target=first)
    t.start()
    assert started.wait(timeout=10)
    try:
        with pytest.raises(ModelCacheManagementError):
            s2.vectorise("slow-b", ["y"], model_properties=props_b, device="cuda:0")
    finally:
        release.set()
        t.join(timeout=20)
    assert result["a"] == [[1.0, 1.0, 1.0, 1.0]]
    # once loaded, concurrent encode() calls need no lock (the reference only locks loading, s2_inference.py:293-298)
    assert s2.vectorise("slow-a", ["x", "y"], model_properties=props_a, device="cuda:0") == [[1.0] * 4, [1.0] * 4]
    s2.clear_loaded_models()


def test_clip_loader_keeps_device_images_on_the_device(monkeypatch):
    """add_docs.py:129-134 hands encode() tensors that are already on the GPU; the loader must stack them there and call
    the device entry point once per image size.  The plumbing is exercised here with host tensors standing in for device
    ones (the 'device' pointers are then readable through ctypes)."""
    import ctypes
    import numpy as np
    import torch
    from marqo_b200.loaders import B200OpenCLIP

    calls = []

    class FakeEncoder:
        device, embed_dim, image_size = 0, 4, 224

        def encode_images_u8_device(self, d_ptr, n, h, w, d_out_ptr, normalize=True, sync=True):
            img = np.ctypeslib.as_array((ctypes.c_uint8 * (n * h * w * 3)).from_address(d_ptr)).reshape(n, h, w, 3)
            out = np.ctypeslib.as_array((ctypes.c_float * (n * 4)).from_address(d_out_ptr)).reshape(n, 4)
            out[:] = np.stack([img.reshape(n, -1).sum(1), np.full(n, h), np.full(n, w), np.full(n, float(normalize))], 1)
            calls.append((n, h, w, sync))

    loader = B200OpenCLIP.__new__(B200OpenCLIP)
    loader.model = FakeEncoder()
    monkeypatch.setattr(B200OpenCLIP, "_on_model_device", lambda self, it: isinstance(it, torch.Tensor))
    monkeypatch.setattr(B200OpenCLIP, "_sync_device", staticmethod(lambda device: None))
    g = torch.Generator().manual_seed(0)
    sizes = [(8, 8), (8, 8), (6, 10), (8, 8), (6, 10)]
    imgs = [torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, generator=g) for h, w in sizes]
    out = loader.encode_image(imgs, normalize=True)
    assert out.shape == (5, 4) and sorted(calls) == [(2, 6, 10, True), (3, 8, 8, True)]
    for i, (h, w) in enumerate(sizes):                      # every row landed at its own position
        assert out[i].tolist() == [float(imgs[i].sum()), h, w, 1.0]
    calls.clear()
    out1 = loader.encode_image(imgs[:2], normalize=False)   # one size: written straight into the result block
    assert calls == [(2, 8, 8, True)] and out1[:, 3].tolist() == [0.0, 0.0]
