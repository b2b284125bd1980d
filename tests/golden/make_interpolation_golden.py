"""Generates tests/golden/interpolation_golden.npz by running THE REFERENCE's own
src/marqo/core/utils/vector_interpolation.py (Lerp / Nlerp / Slerp, :49-237) in the build container, where
/root/reference exists.  The GPU box has no reference: tests read only the committed .npz.

    python tests/golden/make_interpolation_golden.py
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import _reference_import  # noqa: E402

_reference_import.install()
import torchaudio  # noqa: E402

if not hasattr(torchaudio, "set_audio_backend"):   # the reference targets an older torchaudio
    torchaudio.set_audio_backend = lambda *a, **k: None
from marqo.core.models.interpolation_method import InterpolationMethod  # noqa: E402
from marqo.core.utils import vector_interpolation as vi  # noqa: E402

rng = np.random.default_rng(20240922)
out = {}
case = 0
for n, dim in [(1, 16), (2, 16), (3, 64), (5, 64), (8, 768), (13, 384)]:
    for wkind in ("positive", "mixed"):
        vecs = rng.standard_normal((n, dim))
        if case % 3 == 0:
            vecs /= np.linalg.norm(vecs, axis=1, keepdims=True)
        w = rng.uniform(0.1, 2.0, size=n)
        if wkind == "mixed" and n > 1:
            w[rng.integers(0, n)] *= -0.5
        out[f"c{case}_vectors"] = vecs
        out[f"c{case}_weights"] = w
        for method in (InterpolationMethod.LERP, InterpolationMethod.NLERP, InterpolationMethod.SLERP):
            res = vi.from_interpolation_method(method).interpolate(vecs.tolist(), w.tolist())
            out[f"c{case}_{method.value}"] = np.asarray(res, dtype=np.float64)
        case += 1
# co-linear pair: SLERP falls back to linear interpolation (vector_interpolation.py:187-189)
v = rng.standard_normal(32)
out["colinear_vectors"] = np.stack([v, 2.5 * v])
out["colinear_weights"] = np.asarray([1.0, 3.0])
out["colinear_slerp"] = np.asarray(vi.Slerp().interpolate([v.tolist(), (2.5 * v).tolist()], [1.0, 3.0]))
out["n_cases"] = np.asarray(case)
# error behaviour
errs = {}
for name, fn in {
    "lerp_zero_sum": lambda: vi.Lerp().interpolate([[1.0, 0.0], [0.0, 1.0]], [1.0, -1.0]),
    "nlerp_zero_magnitude": lambda: vi.Nlerp().interpolate([[1.0, 0.0], [-1.0, 0.0]], [1.0, 1.0]),
    "slerp_zero_sum": lambda: vi.Slerp().interpolate([[1.0, 0.0], [0.0, 1.0]], [1.0, -1.0]),
    "slerp_zero_length": lambda: vi.Slerp().interpolate([[0.0, 0.0], [0.0, 1.0]], [1.0, 1.0]),
    "empty": lambda: vi.Lerp().interpolate([], []),
    "length_mismatch": lambda: vi.Slerp().interpolate([[1.0, 0.0]], [1.0, 2.0]),
}.items():
    try:
        fn()
        errs[name] = "none"
    except Exception as e:  # noqa: BLE001
        errs[name] = type(e).__name__
out["error_names"] = np.asarray(list(errs.keys()))
out["error_types"] = np.asarray(list(errs.values()))
np.savez_compressed(HERE / "interpolation_golden.npz", **out)
print({k: v for k, v in errs.items()}, case)
