import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    g.build()
    import p4_phylogenetics_b200 as P
    return P


@pytest.fixture(scope="session")
def pf(pkg):
    return pkg.pf


@pytest.fixture(scope="session")
def ref_pf():
    """The reference's own Pf engine (oracle/_ref), the parity truth."""
    import ref_loader
    if not ref_loader.have_ref_pf():
        pytest.skip("oracle/_ref/pf*.so not built (needs /root/reference at build time)")
    return ref_loader.load_ref_pf()
