import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    # Dev container (the reference tree is present): build everything BEFORE collection, so that the `skipif(not ...available())`
    # conditions of the tests pinned to oracle/_ref see the libraries a fresh checkout does not have yet.  On the GPU box the
    # prebuilt files travel with the snapshot and nothing is built here.
    if os.path.isdir("/root/reference/src"):
        try:
            import __graft_entry__ as g
            g.build()
        except Exception as e:  # the `built` fixture reports the failure properly
            print(f"conftest: build() failed before collection: {e}", file=sys.stderr)


@pytest.fixture(scope="session")
def built():
    """Build oracle + product libraries once (no-op when up to date)."""
    import __graft_entry__ as g
    g.build()
    return True
