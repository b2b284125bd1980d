"""Host mirror of src/marqo/core/utils/vector_interpolation.py: the Recommender's LERP / NLERP / SLERP
(src/marqo/core/search/recommender.py:85-88,146-149), same class names, arguments and error classes; the arithmetic runs
behind the C ABI (b200_interpolate_vectors, fp64, the reference's operation order)."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np

from . import _native as N

LERP, NLERP, SLERP = 0, 1, 2
_KIND_ZERO_SUM, _KIND_ZERO_MAGNITUDE, _KIND_ZERO_LENGTH = 1, 2, 3


class ZeroSumWeightsError(ValueError):
    """vector_interpolation.py:12 (an InvalidArgumentError there)."""


class ZeroMagnitudeVectorError(ValueError):
    """vector_interpolation.py:16."""


def _run(method: int, vectors: Sequence[Sequence[float]], weights: Sequence[float]) -> List[float]:
    if len(vectors) < 1:
        raise ValueError("Cannot interpolate an empty list of vectors")                       # :66-67, :145-146
    if len(vectors) != len(weights):
        raise ValueError("Vectors and weights must have the same length")                     # :69-70, :148-149
    dim = len(vectors[0])
    if any(len(v) != dim for v in vectors):
        raise ValueError("Vectors must have the same length")                                 # :82-83, :161-162
    v = np.ascontiguousarray(vectors, dtype=np.float64)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    out = np.empty(dim, dtype=np.float64)
    kind = C.c_int(0)
    lib = N.load()
    st = lib.b200_interpolate_vectors(v.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), len(vectors), dim,
                                      method, out.ctypes.data_as(C.c_void_p), C.byref(kind))
    if st != N.OK:
        msg = (lib.b200_last_error() or b"").decode("utf-8", "replace")
        if kind.value == _KIND_ZERO_SUM:
            raise ZeroSumWeightsError(msg)
        if kind.value == _KIND_ZERO_MAGNITUDE:
            raise ZeroMagnitudeVectorError(msg)
        if kind.value == _KIND_ZERO_LENGTH:
            raise ValueError(msg)
        raise N.NativeError(st, msg)
    return out.tolist()


class VectorInterpolation:
    method = LERP

    def interpolate(self, vectors: List[List[float]], weights: List[float], prenormalized: bool = False) -> List[float]:
        """`prenormalized` is accepted and unused, as in the reference (LERP/NLERP ignore it by contract; the
        hierarchical SLERP never forwards it to _slerp, :232-234)."""
        return _run(self.method, vectors, weights)


class Lerp(VectorInterpolation):
    method = LERP


class Nlerp(Lerp):
    method = NLERP


class Slerp(VectorInterpolation):
    method = SLERP


def from_interpolation_method(method) -> VectorInterpolation:
    """vector_interpolation.py:38-46; accepts the InterpolationMethod enum or its value ('slerp' | 'nlerp' | 'lerp')."""
    name = str(getattr(method, "value", method)).lower()
    table = {"slerp": Slerp, "nlerp": Nlerp, "lerp": Lerp}
    if name not in table:
        raise ValueError(f"Unknown interpolation method: {method}")
    return table[name]()
