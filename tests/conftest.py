import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.build()
    binding.load()
    return binding


@pytest.fixture(scope="session")
def smcrt():
    import rsmcrt_b200
    rsmcrt_b200.load()
    return rsmcrt_b200


@pytest.fixture()
def engine(smcrt):
    e = smcrt.Engine(1)
    yield e
    e.close()


RES = ROOT / "res"
GOLDEN = ROOT / "tests" / "golden"
