import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (VCL-pow variant when it was built in the authoring container)."""
    from oracle import orc

    return orc.Oracle("auto")


@pytest.fixture(scope="session")
def oracle_port():
    from oracle import orc

    return orc.Oracle("port")


def load_golden(name):
    import json

    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)
