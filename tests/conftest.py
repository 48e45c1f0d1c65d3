import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle, build
    build(ref=True)  # liboracle.so always; oracle/_ref only when /root/reference is present
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.pyoracle import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref/libcbs_ref.so not built (reference sources absent)")
    return Ref()


@pytest.fixture(scope="session")
def ctx():
    import genomic_b200
    c = genomic_b200.Context(0)  # raises if the CUDA library or the device is missing: no fallback
    yield c
    c.close()
