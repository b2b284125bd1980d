import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def native_lib():
    from marqo_b200 import build, _native
    build.build_native()
    return _native.load()


@pytest.fixture(scope="session")
def score_oracle():
    from marqo_b200 import build
    build.build_oracle()
    from oracle import score_oracle as so
    return so


@pytest.fixture(scope="session")
def gpu_required(native_lib):
    from marqo_b200 import _native
    if _native.device_count() < 1:
        pytest.fail("no sm_100 device visible: GPU tests must run on the B200 box (there is no CPU fallback)")
    return True
