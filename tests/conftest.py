import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-ann_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import build
    from oracle.oracle import Oracle
    build.build_oracle()
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled from /root/reference (prebuilt .so travels to the GPU box)."""
    from oracle import build
    from oracle.oracle import Ref, have_ref
    build.build_ref()
    if not have_ref():
        pytest.skip("oracle/_ref/libgpann_ref.so not built (needs /root/reference)")
    return Ref()


@pytest.fixture(scope="session")
def ref_mips():
    from oracle import build
    from oracle.oracle import MIPS, Ref, have_ref
    build.build_ref()
    if not have_ref(MIPS):
        pytest.skip("oracle/_ref/libgpann_ref_mips.so not built (needs /root/reference)")
    return Ref(MIPS)


@pytest.fixture(scope="session")
def gpann_mod():
    import gpann
    return gpann


@pytest.fixture(scope="session")
def ctx(gpann_mod):
    """CUDA context on device 0.  No fallback: a missing library or GPU is an error, not a skip."""
    c = gpann_mod.Context(0)
    yield c
    c.close()


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))
