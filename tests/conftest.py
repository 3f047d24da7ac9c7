import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "yocto-gl_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    """The real reference CPU renderer behind oracle/ref_shim.cpp (oracle/_ref/libyocto_ref.so)."""
    import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref/libyocto_ref.so not built (needs /root/reference once)")
    return refbind.Ref()


@pytest.fixture(scope="session")
def ref_dlibm():
    import refbind
    if not refbind.available("_dlibm"):
        pytest.skip("oracle/_ref/libyocto_ref_dlibm.so not built")
    return refbind.Ref("_dlibm")


@pytest.fixture(scope="session")
def ref_count():
    """The instrumented oracle: the reference with traversal counters (oracle/ref_counters.h)."""
    import refbind
    if not refbind.available("_count"):
        pytest.skip("oracle/_ref/libyocto_ref_count.so not built")
    return refbind.Ref("_count")


@pytest.fixture(scope="session")
def ctx():
    from ygl_b200 import lib
    return lib.Context(0)
