import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the compiled reference under oracle/_ref (build container only)")


@pytest.fixture(scope="session", autouse=True)
def _assets():
    """scene files and stand-in assets are generated, not committed"""
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "pathed_b200", "libpathed_host.so")):
        import __graft_entry__
        __graft_entry__.build_host()
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_assets.py")])
