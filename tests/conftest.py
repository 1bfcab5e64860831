import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree artefacts exist (library, oracle restatement, host stepper).  On the GPU box the
    prebuilt files travel with the snapshot; building is a no-op when they are up to date."""
    from jpegdec_b200.build import build
    build()
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "all"], check=True)
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "hostsim"), "-s"], check=True)
    yield
