"""Recommender interpolation through the C ABI against vectors produced by THE REFERENCE's own
vector_interpolation.py (tests/golden/make_interpolation_golden.py).  Host-side fp64: runs without a GPU.

Bar: LERP / NLERP bit-exact is not promised (the reference sums Python floats element by element, the same order used
here, so in practice they agree to the last bit); asserted tolerance 1e-13 relative.  SLERP's angle comes from
numpy.dot (BLAS summation order unspecified): 1e-11."""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden" / "interpolation_golden.npz"


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_matches_reference_outputs(native_lib, gold):
    from marqo_b200 import vector_interpolation as vi
    for c in range(int(gold["n_cases"])):
        vecs, w = gold[f"c{c}_vectors"], gold[f"c{c}_weights"]
        for name, tol in (("lerp", 1e-13), ("nlerp", 1e-13), ("slerp", 1e-11)):
            got = np.asarray(vi.from_interpolation_method(name).interpolate(vecs.tolist(), w.tolist()))
            want = gold[f"c{c}_{name}"]
            assert got.shape == want.shape
            np.testing.assert_allclose(got, want, rtol=tol, atol=tol * np.abs(want).max())
    got = vi.Slerp().interpolate(gold["colinear_vectors"].tolist(), gold["colinear_weights"].tolist())
    np.testing.assert_allclose(got, gold["colinear_slerp"], rtol=1e-12, atol=1e-12)


def test_lerp_is_bitwise_the_reference_arithmetic(native_lib, gold):
    """Lerp.interpolate (:72-88) restated with a plain left-to-right weight sum.  (The golden file was produced on
    Python 3.12, whose built-in sum() of floats is compensated; on the reference's pinned interpreter (3.8/3.9,
    Dockerfile) it is the plain sum used here and by the C ABI — hence 1e-13, not bitwise, against the golden.)"""
    from marqo_b200 import vector_interpolation as vi
    for c in range(int(gold["n_cases"])):
        vecs, w = gold[f"c{c}_vectors"].tolist(), gold[f"c{c}_weights"].tolist()
        wsum = 0.0
        for x in w:
            wsum += x
        want = [0] * len(vecs[0])
        for vec, weight in zip(vecs, w):
            for i, value in enumerate(vec):
                want[i] += (weight / wsum) * value
        got = vi.Lerp().interpolate(vecs, w)
        assert got == want


def test_error_classes_follow_the_reference(native_lib, gold):
    from marqo_b200 import vector_interpolation as vi
    want = dict(zip(gold["error_names"].tolist(), gold["error_types"].tolist()))
    calls = {
        "lerp_zero_sum": lambda: vi.Lerp().interpolate([[1.0, 0.0], [0.0, 1.0]], [1.0, -1.0]),
        "nlerp_zero_magnitude": lambda: vi.Nlerp().interpolate([[1.0, 0.0], [-1.0, 0.0]], [1.0, 1.0]),
        "slerp_zero_sum": lambda: vi.Slerp().interpolate([[1.0, 0.0], [0.0, 1.0]], [1.0, -1.0]),
        "slerp_zero_length": lambda: vi.Slerp().interpolate([[0.0, 0.0], [0.0, 1.0]], [1.0, 1.0]),
        "empty": lambda: vi.Lerp().interpolate([], []),
        "length_mismatch": lambda: vi.Slerp().interpolate([[1.0, 0.0]], [1.0, 2.0]),
    }
    assert set(calls) == set(want)
    for name, fn in calls.items():
        with pytest.raises(Exception) as e:
            fn()
        assert type(e.value).__name__ == want[name], name
    with pytest.raises(ValueError):
        vi.from_interpolation_method("cubic")
    assert isinstance(vi.from_interpolation_method("SLERP"), vi.Slerp)
