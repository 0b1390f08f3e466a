import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ctx():
    import ipfitting_jl_b200 as ipf
    from ipfitting_jl_b200 import _capi
    try:
        c = _capi.Context(0)
    except (_capi.IPFError, OSError) as e:        # no CUDA device / library not built: GPU tests are skipped, not errors
        pytest.skip(f"no CUDA device: {e}")
    yield c
    c.close()
