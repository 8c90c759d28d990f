import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (only in the build container)")


def _have_reference():
    from oracle import ref_loader
    return ref_loader.package_available()


@pytest.fixture(scope="session")
def reference():
    """The real reference package (import from /root/reference); skip if not in this container."""
    from oracle import ref_loader
    if not ref_loader.package_available():
        try:
            from oracle import build_ref
            build_ref.build()
        except Exception:
            pass
    if not ref_loader.package_available():
        pytest.skip("reference package not available here")
    return ref_loader.load_package()
