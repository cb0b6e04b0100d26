import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as entry
    return entry.load_package()


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def golden_philox():
    with open(os.path.join(GOLDEN, "philox_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_reference():
    with open(os.path.join(GOLDEN, "reference_cpu.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def engine(pkg):
    """The CUDA engine.  GPU tests FAIL (not skip) when the library or the device is missing:
    there is no CPU fallback to fall back to."""
    eng = pkg.Engine(0)
    yield eng
    eng.close()
