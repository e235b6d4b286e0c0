import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# The oracle and the host-side generator are OpenMP parallel, and the test problems are small:
# on a many-core box that is shared with other jobs 64+ spinning threads per parallel region
# cost minutes.  Must be set before libgomp is loaded.
os.environ.setdefault("OMP_NUM_THREADS", str(min(8, os.cpu_count() or 1)))
if (os.cpu_count() or 1) > 16:
    # a big box that is shared with other tenants (the GPU boxes: 128 hardware threads, a quota of
    # 16): spinning threads steal from each other; on a small dedicated box spinning is 4x faster
    os.environ.setdefault("OMP_WAIT_POLICY", "passive")


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
