import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ob():
    import oracle_bind
    return oracle_bind


@pytest.fixture(scope="session")
def ref_lib(ob):
    r = ob.ref()
    if r is None:
        pytest.skip("oracle/_ref/libasciichat_ref.so not built (needs /root/reference)")
    return r
